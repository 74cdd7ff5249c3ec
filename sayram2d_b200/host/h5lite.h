// Minimal HDF5 access for the files at the edge of the hot path (libhdf5 / HighFive /
// xtensor-io are not available in this image).
//  read : the subset D/AlbertYoung_chorus.h5 uses - superblock v0, v1 group B-tree +
//         local heap + symbol-table nodes (walked, not hard-coded offsets), v1 object
//         headers, contiguous (layout v3) little-endian f64 datasets without filters.
//  write: the same HDF5 subset (h5lite::Writer: /alpha0, /logEN, /f/<k>, /t of main.cc:58-89, readable
//         by libhdf5 / h5py / plot/cmp_ay.py) and NumPy .npy copies of the snapshots.
#ifndef SY2D_HOST_H5LITE_H_
#define SY2D_HOST_H5LITE_H_

#include <algorithm>
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

// ---------------------------------------------------------------------------------------------
// HDF5 WRITER for the output file of main.cc:58-67,74,83,87-89 (/alpha0, /logEN, /f/<k>, /t):
// the same on-disk flavour as the files the reference reads and writes through HighFive with the
// library defaults - superblock version 0, "old style" groups (version-1 object header with a
// symbol-table message, version-1 B-tree node + local heap + one symbol-table node per group),
// version-1 dataset headers with dataspace (v1), IEEE f64 little-endian datatype, fill-value (v2)
// and contiguous layout (v3) messages, byte for byte the messages of data/D/AlbertYoung_chorus.h5.
// Every group gets ONE symbol-table node, so the superblock's "group leaf node K" is raised to
// half the largest group when a group has more than 8 members (the K values are per-file
// parameters in the superblock; readers take them from there).
// ---------------------------------------------------------------------------------------------
class Writer {
 public:
  // path: "/name" or "/group/.../name"; data is copied
  void add(const std::string& path, const double* data, const std::vector<std::size_t>& shape) {
    if (path.empty() || path[0] != '/') throw std::runtime_error("h5lite::Writer: absolute path expected: " + path);
    Item it;
    it.shape = shape;
    std::size_t n = 1;
    for (auto d : shape) n *= d;
    it.data.assign(data, data + n);
    Group* g = &root_;
    std::size_t pos = 1;
    for (;;) {
      const std::size_t slash = path.find('/', pos);
      if (slash == std::string::npos) break;
      const std::string name = path.substr(pos, slash - pos);
      if (!g->groups.count(name)) g->order.push_back(name);
      g = &g->groups[name];
      pos = slash + 1;
    }
    const std::string leaf = path.substr(pos);
    if (leaf.empty() || g->groups.count(leaf)) throw std::runtime_error("h5lite::Writer: bad dataset name in " + path);
    items_.push_back(std::move(it));
    if (!g->datasets.count(leaf)) g->order.push_back(leaf);
    g->datasets[leaf] = items_.size() - 1;
  }

  void save(const std::string& filename) {
    leafK_ = 4;
    size_leafk(root_);
    buf_.assign(96, 0);          // superblock (56 bytes) + root symbol-table entry (40 bytes)
    const Placed r = place_group(root_);
    const std::uint64_t undef = ~0ull;
    static const unsigned char magic[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    std::memcpy(buf_.data(), magic, 8);
    buf_[13] = 8; buf_[14] = 8;                                   // sizes of offsets / lengths
    put<std::uint16_t>(16, static_cast<std::uint16_t>(leafK_));   // group leaf node K
    put<std::uint16_t>(18, 16);                                   // group internal node K
    put<std::uint64_t>(24, 0);                                    // base address
    put<std::uint64_t>(32, undef);                                // free-space info
    put<std::uint64_t>(40, buf_.size());                          // end-of-file address
    put<std::uint64_t>(48, undef);                                // driver info
    put<std::uint64_t>(56, 0);                                    // root entry: link name offset
    put<std::uint64_t>(64, r.header);
    put<std::uint32_t>(72, 1);                                    // cache type 1: scratch = B-tree, heap
    put<std::uint64_t>(80, r.btree);
    put<std::uint64_t>(88, r.heap);
    std::ofstream out(filename, std::ios::binary);
    if (!out) throw std::runtime_error("h5lite: cannot write " + filename);
    out.write(reinterpret_cast<const char*>(buf_.data()), static_cast<std::streamsize>(buf_.size()));
  }

 private:
  struct Item { std::vector<std::size_t> shape; std::vector<double> data; };
  struct Group { std::map<std::string, Group> groups; std::map<std::string, std::size_t> datasets; std::vector<std::string> order; };
  struct Placed { std::uint64_t header = 0, btree = 0, heap = 0; };

  void size_leafk(const Group& g) {
    const std::size_t n = g.groups.size() + g.datasets.size();
    if (n > 2 * leafK_) leafK_ = (n + 1) / 2;
    for (const auto& kv : g.groups) size_leafk(kv.second);
  }
  std::uint64_t alloc(std::size_t bytes) {
    const std::uint64_t at = (buf_.size() + 7) & ~std::uint64_t(7);
    buf_.resize(at + bytes, 0);
    return at;
  }
  template <class T>
  void put(std::uint64_t off, T v) { std::memcpy(buf_.data() + off, &v, sizeof v); }

  // version-1 object header with one symbol-table message; B-tree node; local heap; symbol-table node
  Placed place_group(const Group& g) {
    Placed p;
    p.header = alloc(16 + 24);
    p.btree = alloc(24 + (2 * 16 + 1) * 8 + 2 * 16 * 8);
    p.heap = alloc(32);
    // heap data: "" at offset 0, then the member names in creation order (as libhdf5 lays them out),
    // then one free block; the symbol-table node lists the members in strcmp order (std::map order)
    std::map<std::string, bool> members;
    for (const auto& kv : g.groups) members[kv.first] = true;
    for (const auto& kv : g.datasets) members[kv.first] = false;
    std::map<std::string, std::uint64_t> offset;
    std::size_t hsize = 8;
    for (const auto& name : g.order) { offset[name] = hsize; hsize += (name.size() + 1 + 7) & ~std::size_t(7); }
    std::vector<std::pair<std::string, std::uint64_t>> names;   // sorted name -> heap offset
    for (const auto& kv : members) names.emplace_back(kv.first, offset[kv.first]);
    const std::size_t free_off = hsize;
    hsize += 16;
    const std::uint64_t hdata = alloc(hsize);
    for (const auto& nm : names) std::memcpy(buf_.data() + hdata + nm.second, nm.first.data(), nm.first.size());
    put<std::uint64_t>(hdata + free_off, 1);       // free block: next = 1 (end of list), size = 16
    put<std::uint64_t>(hdata + free_off + 8, 16);
    std::memcpy(buf_.data() + p.heap, "HEAP", 4);
    put<std::uint64_t>(p.heap + 8, hsize);
    put<std::uint64_t>(p.heap + 16, free_off);
    put<std::uint64_t>(p.heap + 24, hdata);
    const std::uint64_t snod = alloc(8 + 2 * leafK_ * 40);
    // object header
    buf_[p.header] = 1;
    put<std::uint16_t>(p.header + 2, 1);            // one message
    put<std::uint32_t>(p.header + 4, 1);            // reference count
    put<std::uint32_t>(p.header + 8, 24);           // header data size
    put<std::uint16_t>(p.header + 16, 0x0011);      // symbol-table message
    put<std::uint16_t>(p.header + 18, 16);
    put<std::uint64_t>(p.header + 24, p.btree);
    put<std::uint64_t>(p.header + 32, p.heap);
    // B-tree: one leaf entry pointing at the symbol-table node
    std::memcpy(buf_.data() + p.btree, "TREE", 4);
    buf_[p.btree + 4] = 0;                          // node type 0: group
    buf_[p.btree + 5] = 0;                          // level 0
    put<std::uint16_t>(p.btree + 6, names.empty() ? 0 : 1);
    put<std::uint64_t>(p.btree + 8, ~0ull);
    put<std::uint64_t>(p.btree + 16, ~0ull);
    put<std::uint64_t>(p.btree + 24, 0);            // key 0: the empty string
    put<std::uint64_t>(p.btree + 32, snod);
    put<std::uint64_t>(p.btree + 40, names.empty() ? 0 : names.back().second);   // key 1: the largest name
    // symbol-table node
    std::memcpy(buf_.data() + snod, "SNOD", 4);
    buf_[snod + 4] = 1;
    put<std::uint16_t>(snod + 6, static_cast<std::uint16_t>(names.size()));
    std::size_t k = 0;
    for (const auto& nm : names) {
      const std::uint64_t e = snod + 8 + 40 * k++;
      put<std::uint64_t>(e, nm.second);
      if (members[nm.first]) {
        const Placed sub = place_group(g.groups.at(nm.first));
        put<std::uint64_t>(e + 8, sub.header);
        put<std::uint32_t>(e + 16, 1);
        put<std::uint64_t>(e + 24, sub.btree);
        put<std::uint64_t>(e + 32, sub.heap);
      } else {
        put<std::uint64_t>(e + 8, place_dataset(items_[g.datasets.at(nm.first)]));
      }
    }
    return p;
  }

  std::uint64_t place_dataset(const Item& it) {
    const std::size_t rank = it.shape.size();
    const std::size_t space = 8 + 16 * rank;                                  // dataspace message body
    const std::size_t used = (8 + space) + (8 + 24) + (8 + 8) + (8 + 24);
    const std::size_t hsize = std::max<std::size_t>(256, used + 8);           // the library's default header block; rest is a NIL message
    const std::uint64_t h = alloc(16 + hsize);
    const std::uint64_t raw = alloc(it.data.size() * 8);
    std::memcpy(buf_.data() + raw, it.data.data(), it.data.size() * 8);
    buf_[h] = 1;
    put<std::uint16_t>(h + 2, 5);
    put<std::uint32_t>(h + 4, 1);
    put<std::uint32_t>(h + 8, static_cast<std::uint32_t>(hsize));
    std::uint64_t o = h + 16;
    auto msg = [&](std::uint16_t type, std::size_t size, unsigned char flags) { put<std::uint16_t>(o, type); put<std::uint16_t>(o + 2, static_cast<std::uint16_t>(size)); buf_[o + 4] = flags; o += 8; return o; };
    std::uint64_t b = msg(0x0001, space, 0);          // dataspace v1 with maximum dimensions
    buf_[b] = 1; buf_[b + 1] = static_cast<unsigned char>(rank); buf_[b + 2] = 1;
    for (std::size_t d = 0; d < rank; ++d) { put<std::uint64_t>(b + 8 + 8 * d, it.shape[d]); put<std::uint64_t>(b + 8 + 8 * (rank + d), it.shape[d]); }
    o += space;
    b = msg(0x0003, 24, 1);                           // datatype: IEEE 754 binary64, little endian
    static const unsigned char f64[20] = {0x11, 0x20, 0x3f, 0x00, 0x08, 0x00, 0x00, 0x00, 0x00, 0x00, 0x40, 0x00, 0x34, 0x0b, 0x00, 0x34, 0xff, 0x03, 0x00, 0x00};
    std::memcpy(buf_.data() + b, f64, 20);
    o += 24;
    b = msg(0x0005, 8, 1);                            // fill value v2: allocate late, write if set, defined, size 0
    buf_[b] = 2; buf_[b + 1] = 2; buf_[b + 2] = 2; buf_[b + 3] = 1;
    o += 8;
    b = msg(0x0008, 24, 0);                           // layout v3, contiguous
    buf_[b] = 3; buf_[b + 1] = 1;
    put<std::uint64_t>(b + 2, raw);
    put<std::uint64_t>(b + 10, it.data.size() * 8);
    o += 24;
    msg(0x0000, hsize - used - 8, 0);                 // NIL message: the rest of the header block
    return h;
  }

  Group root_;
  std::vector<Item> items_;
  std::vector<unsigned char> buf_;
  std::size_t leafK_ = 4;
};

}  // namespace h5lite

#endif
