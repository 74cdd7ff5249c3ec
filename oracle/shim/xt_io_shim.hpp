// xtensor-io / HighFive look-alike (Albert_Young_IO.cc:21-35 load_hdf5,
// main.cc:58-89 HighFive::File + xt::dump).  TEST INFRASTRUCTURE.
//  * load_hdf5: minimal HDF5 parser (superblock v0, v1 B-tree/heap/SNOD, v1
//    object headers, contiguous little-endian f64) - libhdf5 is not in the image.
//  * HighFive::File(name, Overwrite) creates the directory "<name>.d";
//    xt::dump(file, "/f/3", data, ...) writes "<name>.d/f_3.npy" (NumPy format).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "eigen_shim.hpp"
#include "xt_shim.hpp"

namespace h5shim {

struct Dataset {
  std::vector<std::size_t> shape;
  std::uint64_t addr = 0, bytes = 0;
};

class Reader {
 public:
  explicit Reader(const std::string& path) {
    std::ifstream in(path, std::ios::binary);
    if (!in) throw std::runtime_error("h5shim: cannot open " + path);
    b_.assign(std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>());
    static const unsigned char sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    if (b_.size() < 96 || std::memcmp(b_.data(), sig, 8) != 0) throw std::runtime_error("h5shim: not HDF5: " + path);
    if (b_[8] != 0 || b_[13] != 8 || b_[14] != 8) throw std::runtime_error("h5shim: unsupported superblock");
    base_ = u64(24);
    if (u32(56 + 16) != 1) throw std::runtime_error("h5shim: root group not cached");
    group(u64(56 + 24), u64(56 + 32), "");
  }
  const Dataset& find(const std::string& name) const {
    auto it = ds_.find(name);
    if (it == ds_.end()) throw std::runtime_error("h5shim: no dataset " + name);
    return it->second;
  }
  std::vector<double> read(const std::string& name, std::vector<std::size_t>* shape) const {
    const Dataset& d = find(name);
    std::size_t n = 1;
    for (auto s : d.shape) n *= s;
    if (d.bytes != n * 8 || base_ + d.addr + d.bytes > b_.size()) throw std::runtime_error("h5shim: bad extent " + name);
    std::vector<double> v(n);
    std::memcpy(v.data(), b_.data() + base_ + d.addr, n * 8);
    if (shape) *shape = d.shape;
    return v;
  }

 private:
  std::vector<unsigned char> b_;
  std::uint64_t base_ = 0;
  std::map<std::string, Dataset> ds_;
  std::uint64_t u64(std::size_t o) const { std::uint64_t v; std::memcpy(&v, &b_[o], 8); return v; }
  std::uint32_t u32(std::size_t o) const { std::uint32_t v; std::memcpy(&v, &b_[o], 4); return v; }
  std::uint16_t u16(std::size_t o) const { std::uint16_t v; std::memcpy(&v, &b_[o], 2); return v; }
  bool sig(std::size_t o, const char* s) const { return std::memcmp(&b_[o], s, 4) == 0; }

  void group(std::uint64_t btree, std::uint64_t heap, const std::string& prefix) {
    if (!sig(heap, "HEAP")) throw std::runtime_error("h5shim: bad heap");
    tree(btree, u64(heap + 24), prefix);
  }
  void tree(std::uint64_t a, std::uint64_t heap_data, const std::string& prefix) {
    if (!sig(a, "TREE") || b_[a + 4] != 0) throw std::runtime_error("h5shim: bad group B-tree");
    const int level = b_[a + 5], n = u16(a + 6);
    for (int k = 0; k < n; ++k) {
      const std::uint64_t child = u64(a + 24 + 8 + 16 * k);
      if (level > 0) tree(child, heap_data, prefix); else snod(child, heap_data, prefix);
    }
  }
  void snod(std::uint64_t a, std::uint64_t heap_data, const std::string& prefix) {
    if (!sig(a, "SNOD")) throw std::runtime_error("h5shim: bad SNOD");
    const int n = u16(a + 6);
    for (int k = 0; k < n; ++k) {
      const std::size_t e = a + 8 + 40 * k;
      const std::string name = prefix + "/" + reinterpret_cast<const char*>(&b_[heap_data + u64(e)]);
      if (u32(e + 16) == 1) { group(u64(e + 24), u64(e + 32), name); continue; }
      Dataset d; std::uint64_t st[2] = {0, 0};
      if (header(u64(e + 8), &d, st)) ds_[name] = d; else if (st[0]) group(st[0], st[1], name);
    }
  }
  bool header(std::uint64_t a, Dataset* d, std::uint64_t* stab) {
    if (b_[a] != 1) throw std::runtime_error("h5shim: object header version");
    int nmsg = u16(a + 2), seen = 0; bool has_data = false;
    std::vector<std::pair<std::uint64_t, std::uint64_t>> blocks{{a + 16, u32(a + 8)}};
    for (std::size_t bi = 0; bi < blocks.size() && seen < nmsg; ++bi) {
      std::uint64_t p = blocks[bi].first, end = p + blocks[bi].second;
      while (p + 8 <= end && seen < nmsg) {
        const int type = u16(p), size = u16(p + 2); const std::uint64_t body = p + 8; ++seen;
        if (type == 0x0001) {
          const int ver = b_[body], rank = b_[body + 1]; const std::uint64_t o = body + (ver == 1 ? 8 : 4);
          d->shape.clear(); for (int r = 0; r < rank; ++r) d->shape.push_back(u64(o + 8 * r));
        } else if (type == 0x0003) {
          if ((b_[body] & 0x0f) != 1 || (b_[body + 1] & 1) || u32(body + 4) != 8) throw std::runtime_error("h5shim: only LE f64");
        } else if (type == 0x0008) {
          if (b_[body] != 3 || b_[body + 1] != 1) throw std::runtime_error("h5shim: only contiguous layout v3");
          d->addr = u64(body + 2); d->bytes = u64(body + 10); has_data = true;
        } else if (type == 0x000B) {
          throw std::runtime_error("h5shim: filters unsupported");
        } else if (type == 0x0010) {
          blocks.emplace_back(u64(body), u64(body + 8));
        } else if (type == 0x0011) {
          stab[0] = u64(body); stab[1] = u64(body + 8);
        }
        p = body + size;
      }
    }
    return has_data;
  }
};

inline void write_npy(const std::string& path, const double* data, const std::vector<std::size_t>& shape) {
  std::string sh = "(";
  std::size_t n = 1;
  for (std::size_t k = 0; k < shape.size(); ++k) { sh += std::to_string(shape[k]) + ","; n *= shape[k]; }
  sh += ")";
  std::string hdr = "{'descr': '<f8', 'fortran_order': False, 'shape': " + sh + ", }";
  while ((10 + hdr.size() + 1) % 64 != 0) hdr += ' ';
  hdr += '\n';
  std::ofstream out(path, std::ios::binary);
  if (!out) throw std::runtime_error("h5shim: cannot write " + path);
  const unsigned char magic[8] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0};
  out.write(reinterpret_cast<const char*>(magic), 8);
  const std::uint16_t hl = static_cast<std::uint16_t>(hdr.size());
  out.write(reinterpret_cast<const char*>(&hl), 2);
  out.write(hdr.data(), hdr.size());
  out.write(reinterpret_cast<const char*>(data), n * 8);
}

}  // namespace h5shim

namespace HighFive {
class File {
 public:
  enum AccessMode { ReadOnly = 0, Overwrite = 2 };
  File(const std::string& name, int /*mode*/) : dir_(name + ".d") { std::filesystem::create_directories(dir_); }
  const std::string& dir() const { return dir_; }
 private:
  std::string dir_;
};
}  // namespace HighFive

namespace xt {
enum class dump_mode { create, overwrite };

template <class T> struct load_impl;
template <> struct load_impl<xarray<double>> {
  static xarray<double> go(const std::string& f, const std::string& p) {
    return xarray<double>(h5shim::Reader(f).read(p, nullptr));
  }
};
template <> struct load_impl<xtensor<double, 2>> {
  static xtensor<double, 2> go(const std::string& f, const std::string& p) {
    std::vector<std::size_t> s;
    std::vector<double> v = h5shim::Reader(f).read(p, &s);
    if (s.size() != 2) throw std::runtime_error("h5shim: rank-2 dataset expected: " + p);
    xtensor<double, 2> t;
    t.resize({s[0], s[1]});
    std::memcpy(t.data(), v.data(), v.size() * 8);
    return t;
  }
};
template <class T>
inline T load_hdf5(const std::string& file, const std::string& path) { return load_impl<T>::go(file, path); }

inline std::string dump_name_(const HighFive::File& f, std::string path) {
  while (!path.empty() && path[0] == '/') path.erase(0, 1);
  for (auto& c : path) if (c == '/') c = '_';
  return f.dir() + "/" + path + ".npy";
}
inline void dump(HighFive::File& f, const std::string& path, const xtensor<double, 2>& t, dump_mode) {
  h5shim::write_npy(dump_name_(f, path), t.data(), {t.shape()[0], t.shape()[1]});
}
inline void dump(HighFive::File& f, const std::string& path, const xarray<double>& a, dump_mode) {
  h5shim::write_npy(dump_name_(f, path), a.data(), {a.size()});
}
inline void dump(HighFive::File& f, const std::string& path, const Eigen::VectorXd& v, dump_mode) {
  h5shim::write_npy(dump_name_(f, path), v.data(), {v.size()});
}
}  // namespace xt
