// Minimal HDF5 access for the files at the edge of the hot path (libhdf5 / HighFive /
// xtensor-io are not available in this image).
//  read : what D tables are written with in practice (h5py / libhdf5 / MATLAB defaults and `libver="latest"`): superblock
//         v0 - v3, old-style groups (v1 B-tree + local heap + symbol-table nodes, walked, not hard-coded offsets) and
//         new-style groups with compact links, v1 and v2 object headers, compact / contiguous / chunked (v1 B-tree)
//         layouts, deflate / shuffle / fletcher32 filters, little-endian f64, f32 and integer data (see `File`).
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

struct Filter {
  unsigned id = 0;                    // 1 deflate, 2 shuffle, 3 fletcher32
  std::vector<std::uint32_t> cd;      // client data
};

struct Dataset {
  std::vector<std::size_t> shape;
  std::uint64_t addr = 0, bytes = 0;  // contiguous: address (relative to the base address) and size; compact: absolute offset of the inline data
  int layout = 1;                     // 0 compact, 1 contiguous, 2 chunked (B-tree v1 index)
  int type_class = 1, type_size = 8;  // 0 fixed point, 1 IEEE floating point (little endian)
  bool type_signed = true;
  std::vector<std::size_t> chunk;     // chunked: chunk dimensions in elements
  std::uint64_t btree = 0;            // chunked: address of the chunk B-tree
  std::vector<Filter> filters;        // in the order they were applied when the chunk was written
};

// RFC 1950 / 1951 decoder for the deflate filter (zlib-wrapped chunks); small, bit-by-bit canonical Huffman decoding.
class Inflate {
 public:
  static std::vector<unsigned char> zlib(const unsigned char* in, std::size_t n) {
    if (n < 6 || (in[0] & 0x0f) != 8 || ((in[0] << 8) | in[1]) % 31 != 0 || (in[1] & 0x20)) throw std::runtime_error("h5lite: bad zlib header");
    Inflate z(in + 2, n - 2);
    z.run();
    return std::move(z.out_);
  }

 private:
  const unsigned char* in_;
  std::size_t n_, pos_ = 0;
  std::uint32_t bitbuf_ = 0;
  int bitcnt_ = 0;
  std::vector<unsigned char> out_;
  struct Huff { std::uint16_t count[16]; std::uint16_t symbol[288]; };

  Inflate(const unsigned char* in, std::size_t n) : in_(in), n_(n) {}
  [[noreturn]] static void bad() { throw std::runtime_error("h5lite: corrupt deflate stream"); }
  unsigned bits(int need) {
    std::uint32_t v = bitbuf_;
    while (bitcnt_ < need) {
      if (pos_ >= n_) bad();
      v |= (std::uint32_t)in_[pos_++] << bitcnt_;
      bitcnt_ += 8;
    }
    bitbuf_ = need < 32 ? v >> need : 0;
    bitcnt_ -= need;
    return need < 32 ? v & ((1u << need) - 1u) : v;
  }
  static void build(Huff& h, const std::uint16_t* len, int n) {
    std::fill(h.count, h.count + 16, 0);
    for (int k = 0; k < n; ++k) ++h.count[len[k]];
    std::uint16_t offs[16];
    offs[1] = 0;
    for (int l = 1; l < 15; ++l) offs[l + 1] = offs[l] + h.count[l];
    for (int k = 0; k < n; ++k) if (len[k]) h.symbol[offs[len[k]]++] = (std::uint16_t)k;
  }
  int decode(const Huff& h) {
    int code = 0, first = 0, index = 0;
    for (int l = 1; l <= 15; ++l) {
      code |= (int)bits(1);
      const int count = h.count[l];
      if (code - count < first) return h.symbol[index + (code - first)];
      index += count; first += count; first <<= 1; code <<= 1;
    }
    bad();
  }
  void codes(const Huff& lc, const Huff& dc) {
    static const std::uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const std::uint16_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const std::uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const std::uint16_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    for (;;) {
      int sym = decode(lc);
      if (sym < 256) { out_.push_back((unsigned char)sym); continue; }
      if (sym == 256) return;
      sym -= 257;
      if (sym >= 29) bad();
      const std::size_t len = lbase[sym] + bits(lext[sym]);
      const int ds = decode(dc);
      if (ds >= 30) bad();
      const std::size_t dist = dbase[ds] + bits(dext[ds]);
      if (dist > out_.size()) bad();
      for (std::size_t k = 0; k < len; ++k) out_.push_back(out_[out_.size() - dist]);
    }
  }
  void run() {
    for (bool last = false; !last;) {
      last = bits(1) != 0;
      const unsigned type = bits(2);
      if (type == 0) {   // stored
        bitbuf_ = 0; bitcnt_ = 0;
        if (pos_ + 4 > n_) bad();
        const unsigned len = in_[pos_] | (in_[pos_ + 1] << 8), nlen = in_[pos_ + 2] | (in_[pos_ + 3] << 8);
        pos_ += 4;
        if ((len ^ 0xffffu) != nlen || pos_ + len > n_) bad();
        out_.insert(out_.end(), in_ + pos_, in_ + pos_ + len);
        pos_ += len;
      } else if (type == 1) {   // fixed codes
        std::uint16_t len[320];
        int k = 0;
        for (; k < 144; ++k) len[k] = 8;
        for (; k < 256; ++k) len[k] = 9;
        for (; k < 280; ++k) len[k] = 7;
        for (; k < 288; ++k) len[k] = 8;
        Huff lc, dc;
        build(lc, len, 288);
        for (k = 0; k < 30; ++k) len[k] = 5;
        build(dc, len, 30);
        codes(lc, dc);
      } else if (type == 2) {   // dynamic codes
        static const int order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        const int nlen = (int)bits(5) + 257, ndist = (int)bits(5) + 1, ncode = (int)bits(4) + 4;
        if (nlen > 286 || ndist > 30) bad();
        std::uint16_t len[320];
        std::fill(len, len + 19, 0);
        for (int k = 0; k < ncode; ++k) len[order[k]] = (std::uint16_t)bits(3);
        Huff cc;
        build(cc, len, 19);
        for (int k = 0; k < nlen + ndist;) {
          int sym = decode(cc);
          if (sym < 16) { len[k++] = (std::uint16_t)sym; continue; }
          int rep, val = 0;
          if (sym == 16) { if (k == 0) bad(); val = len[k - 1]; rep = 3 + (int)bits(2); }
          else if (sym == 17) rep = 3 + (int)bits(3);
          else rep = 11 + (int)bits(7);
          if (k + rep > nlen + ndist) bad();
          while (rep--) len[k++] = (std::uint16_t)val;
        }
        Huff lc, dc;
        build(lc, len, nlen);
        build(dc, len + nlen, ndist);
        codes(lc, dc);
      } else {
        bad();
      }
    }
  }
};

// Reader.  Walks the whole file into a name -> Dataset map and converts on read().
//   superblock v0 / v1 (root symbol-table entry) and v2 / v3 (root object header);
//   groups: v1 B-tree + local heap + symbol-table nodes, or compact link messages of new-style groups (dense link
//   storage - fractal heap + v2 B-tree, used above 8 links - is not read);
//   object headers v1 and v2 (continuation blocks of both);
//   datatypes: little-endian IEEE f64 / f32 and fixed point of 1, 2, 4, 8 bytes, all returned as double;
//   layouts (message v3, and the contiguous / compact classes of v4): compact, contiguous, chunked with the v1 chunk
//   B-tree; filters: deflate, shuffle, fletcher32 (checksum skipped).
class File {
 public:
  explicit File(const std::string& path) : path_(path) {
    std::ifstream in(path, std::ios::binary);
    if (!in) throw std::runtime_error("h5lite: cannot open " + path);
    buf_.assign(std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>());
    static const unsigned char magic[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    if (buf_.size() < 48 || std::memcmp(buf_.data(), magic, 8) != 0) fail("not an HDF5 file");
    const unsigned ver = at8(8);
    if (ver <= 1) {
      if (at8(13) != 8 || at8(14) != 8) fail("only 8-byte offsets and lengths are supported");
      const std::size_t o = ver == 0 ? 24 : 28;     // v1 adds the indexed-storage K and two reserved bytes
      base_ = rd<std::uint64_t>(o);
      // root symbol-table entry: name offset, header address, cache type, reserved, scratch
      const std::size_t root = o + 32;
      if (rd<std::uint32_t>(root + 16) == 1) walk_group(rd<std::uint64_t>(root + 24), rd<std::uint64_t>(root + 32), "");
      else walk_object(rd<std::uint64_t>(root + 8), "");
    } else if (ver <= 3) {
      if (at8(9) != 8 || at8(10) != 8) fail("only 8-byte offsets and lengths are supported");
      base_ = rd<std::uint64_t>(12);
      walk_object(rd<std::uint64_t>(36), "");
    } else {
      fail("unknown superblock version");
    }
  }

  bool has(const std::string& name) const { return sets_.count(name) != 0; }
  const std::map<std::string, Dataset>& datasets() const { return sets_; }

  std::vector<double> read(const std::string& name, std::vector<std::size_t>* shape = nullptr) const {
    auto it = sets_.find(name);
    if (it == sets_.end()) fail("no dataset " + name);
    const Dataset& d = it->second;
    std::size_t n = 1;
    for (auto s : d.shape) n *= s;
    std::vector<double> out(n, 0.0);
    const std::size_t es = (std::size_t)d.type_size;
    if (d.layout == 0 || d.layout == 1) {
      const std::size_t at = d.layout == 0 ? (std::size_t)d.addr : (std::size_t)(base_ + d.addr);
      if (d.bytes != n * es || at + d.bytes > buf_.size()) fail("bad extent of " + name);
      for (std::size_t k = 0; k < n; ++k) out[k] = element(d, &buf_[at + k * es]);
    } else {
      read_chunks(d, d.btree, name, out);
    }
    if (shape) *shape = d.shape;
    return out;
  }

 private:
  std::string path_;
  std::vector<unsigned char> buf_;
  std::uint64_t base_ = 0;
  std::map<std::string, Dataset> sets_;
  int depth_ = 0;

  [[noreturn]] void fail(const std::string& why) const { throw std::runtime_error("h5lite: " + path_ + ": " + why); }
  unsigned at8(std::size_t o) const { need(o, 1); return buf_[o]; }
  void need(std::size_t o, std::size_t n) const { if (o + n > buf_.size() || o + n < o) fail("truncated file"); }
  template <class T>
  T rd(std::size_t o) const { need(o, sizeof(T)); T v; std::memcpy(&v, &buf_[o], sizeof(T)); return v; }
  bool tag(std::size_t o, const char* t) const { need(o, 4); return std::memcmp(&buf_[o], t, 4) == 0; }
  std::uint64_t rdn(std::size_t o, unsigned width) const {   // little-endian integer of 1, 2, 4 or 8 bytes
    need(o, width);
    std::uint64_t v = 0;
    for (unsigned k = 0; k < width; ++k) v |= (std::uint64_t)buf_[o + k] << (8 * k);
    return v;
  }

  static double element(const Dataset& d, const unsigned char* p) {
    if (d.type_class == 1) {
      if (d.type_size == 8) { double v; std::memcpy(&v, p, 8); return v; }
      float v; std::memcpy(&v, p, 4); return (double)v;
    }
    std::uint64_t u = 0;
    for (int k = 0; k < d.type_size; ++k) u |= (std::uint64_t)p[k] << (8 * k);
    if (d.type_signed && d.type_size < 8 && (u >> (8 * d.type_size - 1))) u |= ~std::uint64_t(0) << (8 * d.type_size);
    return d.type_signed ? (double)(std::int64_t)u : (double)u;
  }

  // ---- groups ----
  void walk_group(std::uint64_t btree, std::uint64_t heap, const std::string& prefix) {
    heap += base_; btree += base_;
    if (!tag(heap, "HEAP")) fail("bad local heap");
    walk_tree(btree, base_ + rd<std::uint64_t>(heap + 24), prefix);
  }
  void walk_tree(std::uint64_t a, std::uint64_t heap_data, const std::string& prefix) {
    if (!tag(a, "TREE") || at8(a + 4) != 0) fail("bad group B-tree node");
    const unsigned level = at8(a + 5), n = rd<std::uint16_t>(a + 6);
    for (unsigned k = 0; k < n; ++k) {
      const std::uint64_t child = base_ + rd<std::uint64_t>(a + 24 + 8 + 16 * k);  // key0, child0, key1, child1, ...
      if (level > 0) walk_tree(child, heap_data, prefix); else walk_snod(child, heap_data, prefix);
    }
  }
  void walk_snod(std::uint64_t a, std::uint64_t heap_data, const std::string& prefix) {
    if (!tag(a, "SNOD")) fail("bad symbol-table node");
    const unsigned n = rd<std::uint16_t>(a + 6);
    for (unsigned k = 0; k < n; ++k) {
      const std::size_t e = a + 8 + 40 * k;
      const std::size_t name_at = heap_data + rd<std::uint64_t>(e);
      const std::string name = prefix + "/" + cstring(name_at);
      if (rd<std::uint32_t>(e + 16) == 1) {  // group with cached addresses
        walk_group(rd<std::uint64_t>(e + 24), rd<std::uint64_t>(e + 32), name);
        continue;
      }
      walk_object(rd<std::uint64_t>(e + 8), name);
    }
  }
  std::string cstring(std::size_t at) const {
    std::size_t e = at;
    while (at8(e) != 0) ++e;
    return std::string(reinterpret_cast<const char*>(&buf_[at]), e - at);
  }

  // ---- objects ----
  struct Link { std::string name; std::uint64_t addr; };
  void walk_object(std::uint64_t addr, const std::string& name) {
    if (++depth_ > 64) fail("group nesting too deep (a cycle of hard links?)");
    Dataset d;
    std::uint64_t stab[2] = {0, 0};
    std::vector<Link> links;
    const bool is_data = parse_header(base_ + addr, &d, stab, &links);
    if (is_data) sets_[name] = d;
    else if (stab[0] || stab[1]) walk_group(stab[0], stab[1], name);
    for (const Link& l : links) walk_object(l.addr, name + "/" + l.name);
    --depth_;
  }
  // one header message; returns true for the layout message (= the object is a dataset)
  bool message(unsigned type, std::uint64_t body, std::uint64_t size, Dataset* d, std::uint64_t* stab, std::vector<Link>* links,
               std::vector<std::pair<std::uint64_t, std::uint64_t>>* blocks) const {
    need(body, size);
    switch (type) {
      case 0x0001: {  // dataspace
        const unsigned ver = at8(body), rank = at8(body + 1);
        if (ver != 1 && ver != 2) fail("dataspace version");
        const std::uint64_t dims = body + (ver == 1 ? 8 : 4);
        d->shape.clear();
        for (unsigned r = 0; r < rank; ++r) d->shape.push_back(rd<std::uint64_t>(dims + 8 * r));
        return false;
      }
      case 0x0002:  // link info: dense storage when the fractal-heap address is defined
        if (rd<std::uint64_t>(body + 2 + ((at8(body + 1) & 1) ? 8 : 0)) != ~std::uint64_t(0)) fail("dense link storage (more than 8 links in a new-style group) is not supported");
        return false;
      case 0x0003: {  // datatype
        d->type_class = at8(body) & 0x0f;
        d->type_size = (int)rd<std::uint32_t>(body + 4);
        const unsigned bits0 = at8(body + 1);
        if (bits0 & 1) fail("big-endian data");
        if (d->type_class == 1) { if (d->type_size != 8 && d->type_size != 4) fail("floating-point type of unsupported size"); }
        else if (d->type_class == 0) {
          d->type_signed = (bits0 & 0x08) != 0;
          if (d->type_size != 1 && d->type_size != 2 && d->type_size != 4 && d->type_size != 8) fail("fixed-point type of unsupported size");
        } else fail("only floating-point and fixed-point data");
        return false;
      }
      case 0x0006: {  // link (new-style group, compact storage)
        const unsigned flags = at8(body + 1);
        std::uint64_t p = body + 2;
        unsigned ltype = 0;
        if (flags & 0x08) ltype = at8(p++);
        if (flags & 0x04) p += 8;
        if (flags & 0x10) p += 1;
        const unsigned w = 1u << (flags & 3);
        const std::uint64_t len = rdn(p, w);
        p += w;
        need(p, len);
        const std::string lname(reinterpret_cast<const char*>(&buf_[p]), (std::size_t)len);
        p += len;
        if (ltype == 0) links->push_back(Link{lname, rd<std::uint64_t>(p)});   // soft / external links are ignored
        return false;
      }
      case 0x0008: {  // layout
        const unsigned ver = at8(body), cls = at8(body + 1);
        if (ver != 3 && ver != 4) fail("data layout message version " + std::to_string(ver));
        d->layout = (int)cls;
        if (cls == 0) { d->bytes = rd<std::uint16_t>(body + 2); d->addr = body + 4; }
        else if (cls == 1) { d->addr = rd<std::uint64_t>(body + 2); d->bytes = rd<std::uint64_t>(body + 10); }
        else if (cls == 2 && ver == 3) {
          const unsigned nd = at8(body + 2);
          if (nd < 2 || nd > 9) fail("chunk dimensionality");
          d->btree = rd<std::uint64_t>(body + 3);
          d->chunk.clear();
          for (unsigned r = 0; r + 1 < nd; ++r) d->chunk.push_back(rd<std::uint32_t>(body + 11 + 4 * r));
        } else fail("unsupported data layout (class " + std::to_string(cls) + ", message v" + std::to_string(ver) + ")");
        return true;
      }
      case 0x000B: {  // filter pipeline
        const unsigned ver = at8(body), nf = at8(body + 1);
        std::uint64_t p = body + (ver == 1 ? 8 : 2);
        if (ver != 1 && ver != 2) fail("filter pipeline version");
        d->filters.clear();
        for (unsigned k = 0; k < nf; ++k) {
          Filter f;
          f.id = rd<std::uint16_t>(p); p += 2;
          unsigned nlen = 0;
          if (ver == 1 || f.id >= 256) { nlen = rd<std::uint16_t>(p); p += 2; }
          p += 2;   // flags
          const unsigned ncd = rd<std::uint16_t>(p); p += 2;
          p += ver == 1 ? (nlen + 7) / 8 * 8 : nlen;
          for (unsigned c = 0; c < ncd; ++c) { f.cd.push_back(rd<std::uint32_t>(p)); p += 4; }
          if (ver == 1 && (ncd & 1)) p += 4;
          if (f.id != 1 && f.id != 2 && f.id != 3) fail("filter " + std::to_string(f.id) + " is not supported (deflate, shuffle, fletcher32 are)");
          d->filters.push_back(f);
        }
        return false;
      }
      case 0x0010: blocks->emplace_back(base_ + rd<std::uint64_t>(body), rd<std::uint64_t>(body + 8)); return false;
      case 0x0011: stab[0] = rd<std::uint64_t>(body); stab[1] = rd<std::uint64_t>(body + 8); return false;
      default: return false;
    }
  }
  bool parse_header(std::uint64_t a, Dataset* d, std::uint64_t* stab, std::vector<Link>* links) const {
    bool has_data = false;
    std::vector<std::pair<std::uint64_t, std::uint64_t>> blocks;
    if (tag(a, "OHDR")) {   // version 2
      if (at8(a + 4) != 2) fail("object header version");
      const unsigned flags = at8(a + 5);
      std::uint64_t p = a + 6;
      if (flags & 0x20) p += 16;
      if (flags & 0x10) p += 4;
      const unsigned w = 1u << (flags & 3);
      const std::uint64_t size0 = rdn(p, w);
      p += w;
      blocks.emplace_back(p, size0);
      const unsigned mh = 4 + ((flags & 0x04) ? 2 : 0);
      for (std::size_t b = 0; b < blocks.size(); ++b) {
        std::uint64_t q = blocks[b].first, end = q + blocks[b].second;
        if (b > 0) {   // continuation block: "OCHK" + messages + checksum
          if (!tag(q, "OCHK")) fail("bad object header continuation block");
          q += 4; end -= 4;
        }
        need(q, end - q);
        while (q + mh <= end) {
          const unsigned type = at8(q);
          const std::uint64_t size = rd<std::uint16_t>(q + 1);
          const std::uint64_t body = q + mh;
          if (body + size > end) break;   // gap at the end of a block
          if (message(type, body, size, d, stab, links, &blocks)) has_data = true;
          q = body + size;
        }
      }
      return has_data;
    }
    if (at8(a) != 1) fail("object header version");
    const unsigned nmsg = rd<std::uint16_t>(a + 2);
    unsigned seen = 0;
    blocks.emplace_back(a + 16, rd<std::uint32_t>(a + 8));
    for (std::size_t b = 0; b < blocks.size() && seen < nmsg; ++b) {
      std::uint64_t p = blocks[b].first;
      const std::uint64_t end = p + blocks[b].second;
      while (p + 8 <= end && seen < nmsg) {
        const unsigned type = rd<std::uint16_t>(p), size = rd<std::uint16_t>(p + 2);
        ++seen;
        if (message(type, p + 8, size, d, stab, links, &blocks)) has_data = true;
        p += 8 + size;
      }
    }
    return has_data;
  }

  // ---- chunked storage ----
  void read_chunks(const Dataset& d, std::uint64_t node, const std::string& name, std::vector<double>& out) const {
    const std::size_t rank = d.shape.size();
    if (d.chunk.size() != rank || rank == 0) fail("chunk rank of " + name);
    if (node == ~std::uint64_t(0)) return;   // no chunk was ever written: fill value 0
    const std::uint64_t a = base_ + node;
    if (!tag(a, "TREE") || at8(a + 4) != 1) fail("bad chunk B-tree node of " + name);
    const unsigned level = at8(a + 5), n = rd<std::uint16_t>(a + 6);
    const std::size_t key = 8 + 8 * (rank + 1);
    for (unsigned k = 0; k < n; ++k) {
      const std::size_t e = a + 24 + k * (key + 8);
      const std::uint64_t child = rd<std::uint64_t>(e + key);
      if (level > 0) { read_chunks(d, child, name, out); continue; }
      const std::uint32_t csize = rd<std::uint32_t>(e), mask = rd<std::uint32_t>(e + 4);
      std::vector<std::size_t> off(rank);
      for (std::size_t r = 0; r < rank; ++r) off[r] = (std::size_t)rd<std::uint64_t>(e + 8 + 8 * r);
      need(base_ + child, csize);
      std::vector<unsigned char> raw(buf_.begin() + (std::size_t)(base_ + child), buf_.begin() + (std::size_t)(base_ + child) + csize);
      for (std::size_t fk = d.filters.size(); fk-- > 0;) {   // undo the pipeline, last filter first
        if (mask & (1u << fk)) continue;
        const Filter& f = d.filters[fk];
        if (f.id == 3) { if (raw.size() < 4) fail("fletcher32 chunk of " + name); raw.resize(raw.size() - 4); }
        else if (f.id == 1) raw = Inflate::zlib(raw.data(), raw.size());
        else if (f.id == 2) {
          const std::size_t es = f.cd.empty() ? (std::size_t)d.type_size : f.cd[0], ne = es ? raw.size() / es : 0;
          std::vector<unsigned char> un(raw.size());
          for (std::size_t b = 0; b < es; ++b)
            for (std::size_t i = 0; i < ne; ++i) un[i * es + b] = raw[b * ne + i];
          for (std::size_t i = ne * es; i < raw.size(); ++i) un[i] = raw[i];
          raw.swap(un);
        }
      }
      std::size_t celems = 1;
      for (auto c : d.chunk) celems *= c;
      if (raw.size() != celems * (std::size_t)d.type_size) fail("chunk size of " + name);
      // scatter the chunk (row-major) into the array, clipped at the array's edges
      std::vector<std::size_t> idx(rank, 0);
      for (std::size_t c = 0; c < celems; ++c) {
        bool inside = true;
        std::size_t flat = 0;
        for (std::size_t r = 0; r < rank; ++r) {
          const std::size_t g = off[r] + idx[r];
          if (g >= d.shape[r]) { inside = false; break; }
          flat = flat * d.shape[r] + g;
        }
        if (inside) out[flat] = element(d, &raw[c * (std::size_t)d.type_size]);
        for (std::size_t r = rank; r-- > 0;) { if (++idx[r] < d.chunk[r]) break; idx[r] = 0; }
      }
    }
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
