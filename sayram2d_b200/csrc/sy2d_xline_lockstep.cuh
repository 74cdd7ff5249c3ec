// Engine 1 with a segmented x-line preconditioner (large grids and row slabs).
//
// The tridiagonal along i (alpha0) is cut into segments of kSeg rows: T~ = blockdiag of the
// tridiag(wW, 1, wE) restricted to each segment.  Block-Jacobi with these short lines keeps most of
// the benefit of full lines (512^2 synthetic case: 628 iterations unpreconditioned, 144 with full
// lines, 171/182 with segments of 64/32 rows) and needs no communication between threads: one thread
// owns one (segment, column) pair, runs the Thomas recurrence over its kSeg rows in registers, and
// neighbouring threads (adjacent columns) touch adjacent addresses, so every access is coalesced.
// Segments never cross a row slab, so the preconditioner works unchanged in slab mode.
//
// Right preconditioning:  phat = T~^-1 p,   v = A phat = p + wS phat_S + wN phat_N + (the W/E couplings
// that T~ dropped at segment boundaries) * phat_{W/E};   the residual and the stopping rule are those
// of the unpreconditioned iteration.
//
// Kernels per iteration (algorithmic bytes per cell):
//   k_xl_sweep_p  p = r + beta (p - omega v); hat = T~^-1 p        read r p v l dinv e, write p hat     64
//   k_xl_spmv_v   v = p + S/N(hat) + cut(hat); (rhat, v)            read p hat wS wN rhs, write v         48
//   k_xl_sweep_s  s = r - alpha v; x += alpha hat; hat = T~^-1 s    read r v x hat l dinv e, write s x hat 80
//   k_xl_spmv_t   t = s + S/N(hat) + cut(hat); (t,s), (t,t)         read s hat wS wN, write t             40
//   k_xl_xr       x += omega hat; r = s - omega t; (rhat,r), max|r| read x hat s t rhs, write x r         56
// = 288 B per cell and iteration, against 216 B unpreconditioned with ~6x more iterations at 1024^2.
#pragma once
#include "sy2d_kernels.cuh"

namespace sy2d {

constexpr int kSeg = 16;  // rows per segment (power of two)
constexpr int kSweepThreads = 128;

struct XlVecs {
  KrylovVecs k;
  double *l, *dinv, *e, *hat;   // LU factors of T~ and the preconditioned vector
  int ny;
  int row0;      // first owned local row (0, or 1 in a slab with halo rows)
  int nrows;     // owned rows
};

// LU of every segment: l_i = wW_i / d_{i-1} (0 on the first row of a segment), d_i = 1 - l_i wE_{i-1},
// e_i = wE_i / d_i (0 on the last row of a segment).
__global__ void __launch_bounds__(kBlock) k_xl_factor(XlVecs x, size_t N) {
  const int ny = x.ny;
  const int nseg = (x.nrows + kSeg - 1) / kSeg;
  const size_t base = (size_t)blockIdx.y * N;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nseg * ny; t += gridDim.x * blockDim.x) {
    const int s = t / ny, j = t - s * ny;
    const int r0 = s * kSeg, r1 = min(r0 + kSeg, x.nrows);
    double dprev = 1.0, wEprev = 0.0;
    for (int r = r0; r < r1; ++r) {
      const size_t n = base + (size_t)(x.row0 + r) * ny + j;
      const XlineFactor fc = xline_factor(x.k.wW[n], wEprev, dprev, r == r0);
      const double wE = x.k.wE[n];
      x.l[n] = fc.l;
      x.dinv[n] = fc.dinv;
      x.e[n] = r == r1 - 1 ? 0.0 : wE * fc.dinv;
      dprev = fc.d;
      wEprev = wE;
    }
  }
}

// MODE 0: p-update + solve;  MODE 1: s-update, x += alpha*phat, solve.
// Only (rows/kSeg)*ny threads exist (65 k at 1024^2), so a thread must keep many loads in flight.  The
// body is straight-line code for a full segment (CNT == kSeg, no per-row predicates, `first` resolved
// outside), which lets the compiler issue the loads of all rows ahead of the recurrences; ragged last
// segments take the generic path.
template <int MODE, bool FIRST, bool FULL>
__device__ __forceinline__ void xl_sweep_segment(const XlVecs& x, size_t n0, int ny, int cnt, double alpha, double beta, double omega) {
  double z[kSeg], lm[kSeg];
  const double* rsrc = FIRST ? x.k.rhs : x.k.r;
#pragma unroll
  for (int m = 0; m < kSeg; ++m) {
    z[m] = 0.0; lm[m] = 0.0;
    if (FULL || m < cnt) {
      const size_t n = n0 + (size_t)m * ny;
      const double rr = rsrc[n];
      if (MODE == 0) {
        z[m] = FIRST ? rr : rr + beta * (x.k.p[n] - omega * x.k.v[n]);
      } else {
        z[m] = rr - alpha * x.k.v[n];
        const double xn = (FIRST ? 0.0 : x.k.x[n]) + alpha * x.hat[n];
        x.k.x[n] = xn;
      }
      lm[m] = x.l[n];
    }
  }
#pragma unroll
  for (int m = 0; m < kSeg; ++m)
    if (FULL || m < cnt) { if (MODE == 0) x.k.p[n0 + (size_t)m * ny] = z[m]; else x.k.s[n0 + (size_t)m * ny] = z[m]; }
  double carry = 0.0;
#pragma unroll
  for (int m = 0; m < kSeg; ++m) { carry = z[m] - lm[m] * carry; z[m] = carry; }
  double em[kSeg];
#pragma unroll
  for (int m = 0; m < kSeg; ++m) {
    lm[m] = 1.0; em[m] = 0.0;
    if (FULL || m < cnt) { lm[m] = x.dinv[n0 + (size_t)m * ny]; em[m] = x.e[n0 + (size_t)m * ny]; }
  }
  carry = 0.0;
#pragma unroll
  for (int m = kSeg - 1; m >= 0; --m) { carry = (FULL || m < cnt) ? z[m] * lm[m] - em[m] * carry : 0.0; z[m] = carry; }
#pragma unroll
  for (int m = 0; m < kSeg; ++m)
    if (FULL || m < cnt) x.hat[n0 + (size_t)m * ny] = z[m];
}

template <int MODE>
__global__ void __launch_bounds__(kSweepThreads) k_xl_sweep(XlVecs x, size_t N) {
  const Scal* sc = x.k.scal + blockIdx.y;
  if (sc->state != 0) return;
  const int ny = x.ny;
  const int nseg = (x.nrows + kSeg - 1) / kSeg;
  const size_t base = (size_t)blockIdx.y * N;
  const bool first = sc->first;
  const double beta = sc->beta, omega = sc->omega, alpha = sc->alpha;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nseg * ny; t += gridDim.x * blockDim.x) {
    const int s = t / ny, j = t - s * ny;
    const int r0 = s * kSeg;
    const int cnt = min(kSeg, x.nrows - r0);
    const size_t n0 = base + (size_t)(x.row0 + r0) * ny + j;
    if (cnt == kSeg) {
      if (first) xl_sweep_segment<MODE, true, true>(x, n0, ny, cnt, alpha, beta, omega);
      else xl_sweep_segment<MODE, false, true>(x, n0, ny, cnt, alpha, beta, omega);
    } else {
      if (first) xl_sweep_segment<MODE, true, false>(x, n0, ny, cnt, alpha, beta, omega);
      else xl_sweep_segment<MODE, false, false>(x, n0, ny, cnt, alpha, beta, omega);
    }
  }
}

// y = c + wS hat_S + wN hat_N + (W/E couplings cut by the segmentation) for cell n of the local array
__device__ __forceinline__ double xl_apply(const XlVecs& x, const double* __restrict__ c, size_t g, size_t n, size_t N, int ny) {
  const double* hat = x.hat + (g - n);
  const int li = (int)(n / ny) - x.row0;   // owned-row index
  // out-of-range neighbours are clamped to the cell itself: their weights are exactly zero
  double y = c[g] + (x.k.wS[g] * hat[n > 0 ? n - 1 : n] + x.k.wN[g] * hat[n + 1 < N ? n + 1 : n]);
  if ((li & (kSeg - 1)) == 0) y += x.k.wW[g] * hat[n >= (size_t)ny ? n - ny : n];   // first row of a segment
  if ((li & (kSeg - 1)) == kSeg - 1 || li == x.nrows - 1) y += x.k.wE[g] * hat[n + ny < N ? n + ny : n];
  return y;
}

// two cells (n, n+1; ny even => same row) of y = c + wS hat_S + wN hat_N + cut W/E couplings
__device__ __forceinline__ void xl_apply2(const XlVecs& x, const double* __restrict__ c, size_t g, size_t n, size_t N, int ny,
                                          double& y0, double& y1) {
  const double* hat = x.hat + (g - n);
  const int li = (int)(n / ny) - x.row0;
  const double2 cc = ld2(c + g), h = ld2(hat + n), wS = ld2(x.k.wS + g), wN = ld2(x.k.wN + g);
  const double hS = hat[n > 0 ? n - 1 : n], hN = hat[n + 2 < N ? n + 2 : n + 1];
  y0 = cc.x + (wS.x * hS + wN.x * h.y);
  y1 = cc.y + (wS.y * h.x + wN.y * hN);
  if ((li & (kSeg - 1)) == 0) {
    const double2 wW = ld2(x.k.wW + g), hW = ld2(hat + (n >= (size_t)ny ? n - ny : n));
    y0 += wW.x * hW.x; y1 += wW.y * hW.y;
  }
  if ((li & (kSeg - 1)) == kSeg - 1 || li == x.nrows - 1) {
    const double2 wE = ld2(x.k.wE + g), hE = ld2(hat + (n + ny < N ? n + ny : n));
    y0 += wE.x * hE.x; y1 += wE.y * hE.y;
  }
}

template <int V>
__global__ void __launch_bounds__(kBlock, 6) k_xl_spmv_v(XlVecs x, size_t N) {
  __shared__ double red[32];
  Scal* sc = x.k.scal + blockIdx.y;
  if (sc->state != 0) return;
  const size_t base = (size_t)blockIdx.y * N;
  const size_t stride = (size_t)V * gridDim.x * blockDim.x;
  double dot = 0.0;
  for (size_t n = x.k.n_begin + (size_t)V * ((size_t)blockIdx.x * blockDim.x + threadIdx.x); n < x.k.n_end; n += stride) {
    const size_t g = base + n;
    if (V == 2) {
      double v0, v1;
      xl_apply2(x, x.k.p, g, n, N, x.ny, v0, v1);
      st2(x.k.v + g, v0, v1);
      const double2 rh = ld2(x.k.rhs + g);
      dot += rh.x * v0 + rh.y * v1;
    } else {
      const double v = xl_apply(x, x.k.p, g, n, N, x.ny);
      x.k.v[g] = v;
      dot += x.k.rhs[g] * v;
    }
  }
  double sums[1] = {dot};
  block_sums<1>(sums, red);
  {
    __shared__ int last_flag;
    double* const dst[1] = {&sc->acc_rv};
    if (cta_totals<1>(sc, x.k.part + (size_t)blockIdx.y * x.k.part_stride, sums, dst, nullptr, 0.0, &last_flag) && !x.k.defer) {
      const double rv = sc->acc_rv;
      sc->acc_rv = 0.0;
      sc->alpha = rv != 0.0 ? sc->rho / rv : 0.0;
    }
  }
}

template <int V>
__global__ void __launch_bounds__(kBlock, 6) k_xl_spmv_t(XlVecs x, size_t N) {
  __shared__ double red[2 * 32];
  Scal* sc = x.k.scal + blockIdx.y;
  if (sc->state != 0) return;
  const size_t base = (size_t)blockIdx.y * N;
  const size_t stride = (size_t)V * gridDim.x * blockDim.x;
  double ts = 0.0, tt = 0.0;
  for (size_t n = x.k.n_begin + (size_t)V * ((size_t)blockIdx.x * blockDim.x + threadIdx.x); n < x.k.n_end; n += stride) {
    const size_t g = base + n;
    if (V == 2) {
      double t0, t1;
      xl_apply2(x, x.k.s, g, n, N, x.ny, t0, t1);
      st2(x.k.t + g, t0, t1);
      const double2 sv = ld2(x.k.s + g);
      ts += t0 * sv.x + t1 * sv.y;
      tt += t0 * t0 + t1 * t1;
    } else {
      const double t = xl_apply(x, x.k.s, g, n, N, x.ny);
      x.k.t[g] = t;
      ts += t * x.k.s[g];
      tt += t * t;
    }
  }
  double sums[2] = {ts, tt};
  block_sums<2>(sums, red);
  {
    __shared__ int last_flag;
    double* const dst[2] = {&sc->acc_ts, &sc->acc_tt};
    if (cta_totals<2>(sc, x.k.part + (size_t)blockIdx.y * x.k.part_stride, sums, dst, nullptr, 0.0, &last_flag) && !x.k.defer) {
      const double a = sc->acc_ts, b = sc->acc_tt;
      sc->acc_ts = 0.0;
      sc->acc_tt = 0.0;
      sc->omega = b > 0.0 ? a / b : 0.0;
    }
  }
}

template <int V>
__global__ void __launch_bounds__(kBlock, 6) k_xl_xr(XlVecs x, size_t N) {
  __shared__ double red[32];
  Scal* sc = x.k.scal + blockIdx.y;
  if (sc->state != 0) return;
  const size_t base = (size_t)blockIdx.y * N;
  const size_t stride = (size_t)V * gridDim.x * blockDim.x;
  const double omega = sc->omega;
  double dot = 0.0, rabs = 0.0;
  for (size_t n = x.k.n_begin + (size_t)V * ((size_t)blockIdx.x * blockDim.x + threadIdx.x); n < x.k.n_end; n += stride) {
    const size_t g = base + n;
    if (V == 2) {
      const double2 s = ld2(x.k.s + g), h = ld2(x.hat + g), xx = ld2(x.k.x + g), t = ld2(x.k.t + g), rh = ld2(x.k.rhs + g);
      st2(x.k.x + g, xx.x + omega * h.x, xx.y + omega * h.y);
      const double r0 = s.x - omega * t.x, r1 = s.y - omega * t.y;
      st2(x.k.r + g, r0, r1);
      dot += rh.x * r0 + rh.y * r1;
      rabs = nmax(rabs, nmax(fabs(r0), fabs(r1)));
    } else {
      const double s = x.k.s[g];
      x.k.x[g] += omega * x.hat[g];
      const double r = s - omega * x.k.t[g];
      x.k.r[g] = r;
      dot += x.k.rhs[g] * r;
      rabs = nmax(rabs, fabs(r));
    }
  }
  double sums[1] = {dot};
  block_sums<1>(sums, red);
  const double bmax = block_max(rabs, red);
  {
    __shared__ int last_flag;
    double* const dst[1] = {&sc->acc_rho};
    if (cta_totals<1>(sc, x.k.part + (size_t)blockIdx.y * x.k.part_stride, sums, dst, &sc->acc_rmax, bmax, &last_flag) && !x.k.defer) xr_finish_iteration(sc, x.k);
  }
}

}  // namespace sy2d
