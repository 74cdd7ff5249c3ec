// Common types and constants of the host layer (mirrors the names of the reference's
// source/common.h:22-44 so that case code written against the reference compiles here),
// without Eigen/xtensor: Xtensor2d is a small row-major 2-D array with the subset of
// the xtensor API the reference's public interfaces expose.
#ifndef SY2D_HOST_COMMON_H_
#define SY2D_HOST_COMMON_H_

#include <array>
#include <cmath>
#include <cstddef>
#include <initializer_list>
#include <iostream>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

// (nx, ny) row-major, last index (j = log E) fastest: the layout of xt::xtensor<double,2>
// and of every field crossing the C ABI (include/sayram2d.h).
class Xtensor2d {
 public:
  using shape_type = std::array<std::size_t, 2>;
  Xtensor2d() : shape_{{0, 0}} {}
  Xtensor2d(std::size_t nx, std::size_t ny, double v = 0.0) : shape_{{nx, ny}}, d_(nx * ny, v) {}
  void resize(std::initializer_list<std::size_t> s) { auto it = s.begin(); const std::size_t a = *it++; resize(shape_type{{a, *it}}); }
  void resize(const shape_type& s) { shape_ = s; d_.resize(s[0] * s[1]); }
  void fill(double v) { for (auto& x : d_) x = v; }
  double& operator()(std::size_t i, std::size_t j) { return d_[i * shape_[1] + j]; }
  double operator()(std::size_t i, std::size_t j) const { return d_[i * shape_[1] + j]; }
  const shape_type& shape() const { return shape_; }
  std::size_t size() const { return d_.size(); }
  double* data() { return d_.data(); }
  const double* data() const { return d_.data(); }

 private:
  shape_type shape_;
  std::vector<double> d_;
};

class Xarray1d {
 public:
  Xarray1d() = default;
  explicit Xarray1d(std::vector<double> v) : d_(std::move(v)) {}
  std::size_t size() const { return d_.size(); }
  double& operator[](std::size_t i) { return d_[i]; }
  double operator[](std::size_t i) const { return d_[i]; }
  Xarray1d operator*(double s) const { Xarray1d r(*this); for (auto& x : r.d_) x = x * s; return r; }
  Xarray1d operator/(double s) const { Xarray1d r(*this); for (auto& x : r.d_) x = x / s; return r; }
  const double* data() const { return d_.data(); }

 private:
  std::vector<double> d_;
};

struct Loc {  // bilinear lookup position in the D table
  int i0, j0;
  double wi, wj;
};

// same values as source/common.h:38-44
const double gEPS = std::numeric_limits<double>::epsilon();
const double gPI = 3.141592653589793238462;
const double gD2R = gPI / 180.0;
const double gC = 1;
const double gE0 = 0.511875;  // MeV
const double gME = gE0 / (gC * gC);
const double gRE = 6371000;

#endif
