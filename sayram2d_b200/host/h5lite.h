// Minimal HDF5 access for the files at the edge of the hot path (libhdf5 / HighFive /
// xtensor-io are not available in this image).
//  read : the subset D/AlbertYoung_chorus.h5 uses - superblock v0, v1 group B-tree +
//         local heap + symbol-table nodes (walked, not hard-coded offsets), v1 object
//         headers, contiguous (layout v3) little-endian f64 datasets without filters.
//  write: NumPy .npy (the /f/<k> snapshots of main.cc:74,83 become f_<k>.npy).
#ifndef SY2D_HOST_H5LITE_H_
#define SY2D_HOST_H5LITE_H_

#include <cstdint>
#include <cstring>
#include <fstream>
#include <iterator>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace h5lite {

struct Dataset {
  std::vector<std::size_t> shape;
  std::uint64_t addr = 0, bytes = 0;
};

class File {
 public:
  explicit File(const std::string& path) : path_(path) {
    std::ifstream in(path, std::ios::binary);
    if (!in) throw std::runtime_error("h5lite: cannot open " + path);
    buf_.assign(std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>());
    static const unsigned char magic[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    if (buf_.size() < 96 || std::memcmp(buf_.data(), magic, 8) != 0) fail("not an HDF5 file");
    if (at8(8) != 0) fail("only superblock version 0 is supported");
    if (at8(13) != 8 || at8(14) != 8) fail("only 8-byte offsets and lengths are supported");
    base_ = rd<std::uint64_t>(24);
    // root symbol-table entry at byte 56: name offset, header address, cache type, scratch
    if (rd<std::uint32_t>(56 + 16) != 1) fail("root group without cached B-tree/heap addresses");
    walk_group(rd<std::uint64_t>(56 + 24), rd<std::uint64_t>(56 + 32), "");
  }

  bool has(const std::string& name) const { return sets_.count(name) != 0; }
  const std::map<std::string, Dataset>& datasets() const { return sets_; }

  std::vector<double> read(const std::string& name, std::vector<std::size_t>* shape = nullptr) const {
    auto it = sets_.find(name);
    if (it == sets_.end()) fail("no dataset " + name);
    const Dataset& d = it->second;
    std::size_t n = 1;
    for (auto s : d.shape) n *= s;
    if (d.bytes != n * 8 || base_ + d.addr + d.bytes > buf_.size()) fail("bad extent of " + name);
    std::vector<double> out(n);
    std::memcpy(out.data(), buf_.data() + base_ + d.addr, n * 8);
    if (shape) *shape = d.shape;
    return out;
  }

 private:
  std::string path_;
  std::vector<unsigned char> buf_;
  std::uint64_t base_ = 0;
  std::map<std::string, Dataset> sets_;

  [[noreturn]] void fail(const std::string& why) const { throw std::runtime_error("h5lite: " + path_ + ": " + why); }
  unsigned at8(std::size_t o) const { need(o, 1); return buf_[o]; }
  void need(std::size_t o, std::size_t n) const { if (o + n > buf_.size()) fail("truncated file"); }
  template <class T>
  T rd(std::size_t o) const { need(o, sizeof(T)); T v; std::memcpy(&v, &buf_[o], sizeof(T)); return v; }
  bool tag(std::size_t o, const char* t) const { need(o, 4); return std::memcmp(&buf_[o], t, 4) == 0; }

  void walk_group(std::uint64_t btree, std::uint64_t heap, const std::string& prefix) {
    if (!tag(heap, "HEAP")) fail("bad local heap");
    walk_tree(btree, rd<std::uint64_t>(heap + 24), prefix);
  }
  void walk_tree(std::uint64_t a, std::uint64_t heap_data, const std::string& prefix) {
    if (!tag(a, "TREE") || at8(a + 4) != 0) fail("bad group B-tree node");
    const unsigned level = at8(a + 5), n = rd<std::uint16_t>(a + 6);
    for (unsigned k = 0; k < n; ++k) {
      const std::uint64_t child = rd<std::uint64_t>(a + 24 + 8 + 16 * k);  // key0, child0, key1, child1, ...
      if (level > 0) walk_tree(child, heap_data, prefix); else walk_snod(child, heap_data, prefix);
    }
  }
  void walk_snod(std::uint64_t a, std::uint64_t heap_data, const std::string& prefix) {
    if (!tag(a, "SNOD")) fail("bad symbol-table node");
    const unsigned n = rd<std::uint16_t>(a + 6);
    for (unsigned k = 0; k < n; ++k) {
      const std::size_t e = a + 8 + 40 * k;
      const std::size_t name_at = heap_data + rd<std::uint64_t>(e);
      need(name_at, 1);
      const std::string name = prefix + "/" + reinterpret_cast<const char*>(&buf_[name_at]);
      if (rd<std::uint32_t>(e + 16) == 1) {  // group with cached addresses
        walk_group(rd<std::uint64_t>(e + 24), rd<std::uint64_t>(e + 32), name);
        continue;
      }
      Dataset d;
      std::uint64_t stab[2] = {0, 0};
      if (parse_header(rd<std::uint64_t>(e + 8), &d, stab)) sets_[name] = d;
      else if (stab[0]) walk_group(stab[0], stab[1], name);
    }
  }
  bool parse_header(std::uint64_t a, Dataset* d, std::uint64_t* stab) const {
    if (at8(a) != 1) fail("only version-1 object headers are supported");
    const unsigned nmsg = rd<std::uint16_t>(a + 2);
    unsigned seen = 0;
    bool has_data = false;
    std::vector<std::pair<std::uint64_t, std::uint64_t>> blocks{{a + 16, rd<std::uint32_t>(a + 8)}};
    for (std::size_t b = 0; b < blocks.size() && seen < nmsg; ++b) {
      std::uint64_t p = blocks[b].first;
      const std::uint64_t end = p + blocks[b].second;
      while (p + 8 <= end && seen < nmsg) {
        const unsigned type = rd<std::uint16_t>(p), size = rd<std::uint16_t>(p + 2);
        const std::uint64_t body = p + 8;
        ++seen;
        switch (type) {
          case 0x0001: {  // dataspace
            const unsigned ver = at8(body), rank = at8(body + 1);
            if (ver != 1 && ver != 2) fail("dataspace version");
            const std::uint64_t dims = body + (ver == 1 ? 8 : 4);
            d->shape.clear();
            for (unsigned r = 0; r < rank; ++r) d->shape.push_back(rd<std::uint64_t>(dims + 8 * r));
            break;
          }
          case 0x0003:  // datatype: class 1 (floating point), little endian, 8 bytes
            if ((at8(body) & 0x0f) != 1 || (at8(body + 1) & 1) || rd<std::uint32_t>(body + 4) != 8) fail("only little-endian f64 data");
            break;
          case 0x0008:  // layout
            if (at8(body) != 3 || at8(body + 1) != 1) fail("only contiguous layout (message v3)");
            d->addr = rd<std::uint64_t>(body + 2);
            d->bytes = rd<std::uint64_t>(body + 10);
            has_data = true;
            break;
          case 0x000B: fail("filtered datasets are not supported");
          case 0x0010: blocks.emplace_back(rd<std::uint64_t>(body), rd<std::uint64_t>(body + 8)); break;
          case 0x0011: stab[0] = rd<std::uint64_t>(body); stab[1] = rd<std::uint64_t>(body + 8); break;
          default: break;
        }
        p = body + size;
      }
    }
    return has_data;
  }
};

inline void write_npy(const std::string& path, const double* data, const std::vector<std::size_t>& shape) {
  std::string dims = "(";
  std::size_t n = 1;
  for (auto s : shape) { dims += std::to_string(s) + ","; n *= s; }
  dims += ")";
  std::string hdr = "{'descr': '<f8', 'fortran_order': False, 'shape': " + dims + ", }";
  while ((10 + hdr.size() + 1) % 64 != 0) hdr += ' ';
  hdr += '\n';
  std::ofstream out(path, std::ios::binary);
  if (!out) throw std::runtime_error("h5lite: cannot write " + path);
  const unsigned char magic[8] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0};
  const std::uint16_t len = static_cast<std::uint16_t>(hdr.size());
  out.write(reinterpret_cast<const char*>(magic), 8);
  out.write(reinterpret_cast<const char*>(&len), 2);
  out.write(hdr.data(), static_cast<std::streamsize>(hdr.size()));
  out.write(reinterpret_cast<const char*>(data), static_cast<std::streamsize>(n * 8));
}

}  // namespace h5lite

#endif
