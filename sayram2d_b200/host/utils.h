// Scalar helpers with the reference's names and formulas (source/utils.h:7-51).
#ifndef SY2D_HOST_UTILS_H_
#define SY2D_HOST_UTILS_H_

#include <cmath>

#include "common.h"

inline double p2e(double p, double E0) { return std::sqrt(p * p * gC * gC + E0 * E0) - E0; }
inline double e2p(double E, double E0) { return std::sqrt(E * (E + 2 * E0)) / gC; }
inline double dlogE_dp(double logE, double E0) {
  const double E = std::exp(logE);
  return e2p(E, gE0) * gC * gC / (E * (E + E0));
}

// bilinear interpolation in a table at a located position
inline double interp2D(const Xtensor2d& raw, const Loc& loc) {
  const int i = loc.i0, j = loc.j0;
  const double wi = loc.wi, wj = loc.wj;
  return raw(i, j) * wi * wj + raw(i + 1, j) * (1 - wi) * wj + raw(i + 1, j + 1) * (1 - wi) * (1 - wj) + raw(i, j + 1) * wi * (1 - wj);
}

// Index/weight of a fractional table position, clamped to [0, n-1].  The reference's
// version takes an unsigned index, so its `i < 0` branch never fires: a position below the
// table wraps around and is treated like one above it (i = n-1, w = 0).  Reproduced here.
inline void calWeight(std::size_t& i, double& w, std::size_t n, double pos) {
  if (i < n) {
    w = 1.0 - (pos - static_cast<double>(i));
  } else {
    i = n - 1;
    w = 0.0;
  }
}

#endif
