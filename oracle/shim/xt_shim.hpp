// xtensor look-alike: only what the Sayram-2D reference sources touch
// (common.h:27-28, Solver.cc:20-28, Equation.h:29-42, Albert_Young_IO.cc:21-35,
// main.cc:74-89).  TEST INFRASTRUCTURE: lets the reference's .cc files compile
// unmodified from /root/reference into oracle/_ref/.
#pragma once
#include <array>
#include <cstddef>
#include <initializer_list>
#include <vector>

namespace xt {

template <class T, std::size_t N>
class xtensor;

// row-major 2-D container, last index fastest (xtensor's default layout)
template <class T>
class xtensor<T, 2> {
 public:
  using shape_type = std::array<std::size_t, 2>;
  xtensor() : shape_{{0, 0}} {}
  void resize(std::initializer_list<std::size_t> s) {
    auto it = s.begin();
    shape_[0] = *it++;
    shape_[1] = *it;
    d_.resize(shape_[0] * shape_[1]);
  }
  void resize(const shape_type& s) {
    shape_ = s;
    d_.resize(shape_[0] * shape_[1]);
  }
  void fill(const T& v) {
    for (auto& x : d_) x = v;
  }
  T& operator()(std::size_t i, std::size_t j) { return d_[i * shape_[1] + j]; }
  const T& operator()(std::size_t i, std::size_t j) const { return d_[i * shape_[1] + j]; }
  const shape_type& shape() const { return shape_; }
  std::size_t size() const { return d_.size(); }
  T* data() { return d_.data(); }
  const T* data() const { return d_.data(); }

 private:
  shape_type shape_;
  std::vector<T> d_;
};

// dynamic-rank array; the reference only uses it as a 1-D vector of doubles
template <class T>
class xarray {
 public:
  xarray() = default;
  explicit xarray(std::vector<T> v) : d_(std::move(v)) {}
  std::size_t size() const { return d_.size(); }
  T& operator[](std::size_t i) { return d_[i]; }
  const T& operator[](std::size_t i) const { return d_[i]; }
  xarray operator*(T s) const {
    xarray r(*this);
    for (auto& x : r.d_) x = x * s;
    return r;
  }
  xarray operator/(T s) const {
    xarray r(*this);
    for (auto& x : r.d_) x = x / s;
    return r;
  }
  const T* data() const { return d_.data(); }
  std::vector<T>& storage() { return d_; }

 private:
  std::vector<T> d_;
};

template <class T>
inline xarray<T> linspace(T a, T b, std::size_t n) {
  std::vector<T> v(n);
  const T step = (n > 1) ? (b - a) / static_cast<T>(n - 1) : T(0);
  for (std::size_t i = 0; i < n; ++i) v[i] = a + step * static_cast<T>(i);
  return xarray<T>(std::move(v));
}

}  // namespace xt
