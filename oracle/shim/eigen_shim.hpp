// Eigen look-alike: the exact API surface the Sayram-2D reference sources use
// (grep of /root/reference/source: Solver.h:35, Solver.cc:53-54,84-85,201,276-278,
// Edge.h:43-50, main.cc:60-64, common.h:22-25), so that the reference's own
// .cc files compile UNMODIFIED from /root/reference into oracle/_ref/.
//
// TEST INFRASTRUCTURE.  Eigen itself is absent from this image (and unpinned by
// the reference), so the arithmetic it would contribute is restated here:
//   * 2x2 / 2-vector products: plain IEEE double expressions, same association
//     order as Eigen's lazy products ((n^T * Lambda) * r).
//   * SparseLU<SpMat, COLAMDOrdering<int>>: a direct LU WITHOUT pivoting.  The
//     PPFV matrix is strictly column diagonally dominant, so
//     partial pivoting never moves off the diagonal and the factorisation is
//     backward stable; the result equals Eigen's to round-off (cond*eps ~ 3e-11).
//     Small problems use a dense-band kernel; large ones a left-looking sparse
//     LU on a nested-dissection ordering (see sparse_lu_shim.hpp).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <stdexcept>
#include <vector>

namespace Eigen {

enum { ColMajor = 0, RowMajor = 1 };

struct Vector2d;
struct Matrix2d;

struct RowVector2d {
  double v[2];
  double operator()(std::size_t i) const { return v[i]; }
};

struct Scalar1 {  // 1x1 product result: (a^T b)(0)
  double s;
  double operator()(std::size_t) const { return s; }
  operator double() const { return s; }
};

struct Vector2d {
  double v[2];
  Vector2d() : v{0.0, 0.0} {}
  Vector2d(double a, double b) : v{a, b} {}
  double& operator()(std::size_t i) { return v[i]; }
  double operator()(std::size_t i) const { return v[i]; }
  double& operator[](std::size_t i) { return v[i]; }
  double operator[](std::size_t i) const { return v[i]; }
  Vector2d operator-(const Vector2d& o) const { return Vector2d(v[0] - o.v[0], v[1] - o.v[1]); }
  Vector2d operator+(const Vector2d& o) const { return Vector2d(v[0] + o.v[0], v[1] + o.v[1]); }
  RowVector2d transpose() const { return RowVector2d{{v[0], v[1]}}; }
  double norm() const { return std::sqrt(v[0] * v[0] + v[1] * v[1]); }
};

struct Matrix2d {
  double m[2][2];  // m[row][col]
  Matrix2d() : m{{0.0, 0.0}, {0.0, 0.0}} {}
  struct CommaInit {
    Matrix2d& M;
    int k;
    CommaInit& operator,(double x) {
      M.m[k / 2][k % 2] = x;
      ++k;
      return *this;
    }
  };
  CommaInit operator<<(double x) {
    m[0][0] = x;
    return CommaInit{*this, 1};
  }
  double operator()(std::size_t r, std::size_t c) const { return m[r][c]; }
};

inline RowVector2d operator*(const RowVector2d& a, const Matrix2d& M) {
  return RowVector2d{{a.v[0] * M.m[0][0] + a.v[1] * M.m[1][0], a.v[0] * M.m[0][1] + a.v[1] * M.m[1][1]}};
}
inline Scalar1 operator*(const RowVector2d& a, const Vector2d& b) {
  return Scalar1{a.v[0] * b.v[0] + a.v[1] * b.v[1]};
}

// ---------------------------------------------------------------- VectorXd
struct ArrayXd;
struct VectorXd {
  std::vector<double> d;
  VectorXd() = default;
  explicit VectorXd(std::size_t n) : d(n, 0.0) {}
  void resize(std::size_t n) { d.resize(n); }
  void setZero() { std::fill(d.begin(), d.end(), 0.0); }
  std::size_t size() const { return d.size(); }
  double& operator()(std::size_t i) { return d[i]; }
  double operator()(std::size_t i) const { return d[i]; }
  double* data() { return d.data(); }
  const double* data() const { return d.data(); }
  VectorXd operator/(double s) const {
    VectorXd r(*this);
    for (auto& x : r.d) x = x / s;
    return r;
  }
  VectorXd operator*(double s) const {
    VectorXd r(*this);
    for (auto& x : r.d) x = x * s;
    return r;
  }
  inline VectorXd(const ArrayXd& a);
};
struct ArrayXd {
  std::vector<double> d;
  ArrayXd operator-(double s) const {
    ArrayXd r(*this);
    for (auto& x : r.d) x -= s;
    return r;
  }
};
inline VectorXd::VectorXd(const ArrayXd& a) : d(a.d) {}

template <class V>
struct Map;
template <>
struct Map<const VectorXd> {
  const double* p;
  std::size_t n;
  Map(const double* p_, std::size_t n_) : p(p_), n(n_) {}
  VectorXd operator/(double s) const {
    VectorXd r(n);
    for (std::size_t i = 0; i < n; ++i) r.d[i] = p[i] / s;
    return r;
  }
  ArrayXd array() const {
    ArrayXd a;
    a.d.assign(p, p + n);
    return a;
  }
};

// ---------------------------------------------------------------- sparse
template <class S>
struct Triplet {
  long r, c;
  S v;
  Triplet() : r(0), c(0), v(0) {}
  Triplet(long r_, long c_, S v_) : r(r_), c(c_), v(v_) {}
  long row() const { return r; }
  long col() const { return c; }
  S value() const { return v; }
};

template <class S, int Opt = ColMajor>
struct SparseMatrix {
  long nrows = 0, ncols = 0;
  std::vector<long> colptr;  // CSC
  std::vector<int> rowind;
  std::vector<S> val;
  void resize(long r, long c) {
    nrows = r;
    ncols = c;
    colptr.assign(c + 1, 0);
    rowind.clear();
    val.clear();
  }
  long rows() const { return nrows; }
  long cols() const { return ncols; }
  long nonZeros() const { return (long)val.size(); }
  void makeCompressed() {}
  template <class It>
  void setFromTriplets(It b, It e) {  // duplicates are summed (Eigen semantics)
    std::vector<long> cnt(ncols + 1, 0);
    for (It t = b; t != e; ++t) cnt[t->col() + 1]++;
    for (long c = 0; c < ncols; ++c) cnt[c + 1] += cnt[c];
    std::vector<int> ri(cnt[ncols]);
    std::vector<S> vv(cnt[ncols]);
    std::vector<long> pos(cnt.begin(), cnt.end() - 1);
    for (It t = b; t != e; ++t) {
      long p = pos[t->col()]++;
      ri[p] = (int)t->row();
      vv[p] = t->value();
    }
    colptr.assign(ncols + 1, 0);
    rowind.clear();
    val.clear();
    std::vector<std::pair<int, S>> tmp;
    for (long c = 0; c < ncols; ++c) {
      tmp.clear();
      for (long p = cnt[c]; p < cnt[c + 1]; ++p) tmp.emplace_back(ri[p], vv[p]);
      std::stable_sort(tmp.begin(), tmp.end(), [](const auto& a, const auto& b2) { return a.first < b2.first; });
      for (std::size_t q = 0; q < tmp.size(); ++q) {
        if (!rowind.empty() && (long)rowind.size() > colptr[c] && rowind.back() == tmp[q].first)
          val.back() += tmp[q].second;
        else {
          rowind.push_back(tmp[q].first);
          val.push_back(tmp[q].second);
        }
      }
      colptr[c + 1] = (long)rowind.size();
    }
  }
};

template <class I>
struct COLAMDOrdering {};

}  // namespace Eigen

#include "sparse_lu_shim.hpp"
