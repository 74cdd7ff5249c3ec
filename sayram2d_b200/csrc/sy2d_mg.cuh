// Engine 1 with a multigrid preconditioner: y-semi-coarsening + damped x-line Jacobi smoothing.
//
// Why: the scaled operator A = I + W/E/S/N couplings is strongly anisotropic along i (alpha0):
// relative to the mass term the x coupling is O(1e4) and the y coupling O(4e2) at 1024^2.  Exact
// x-line solves remove the x stiffness, which leaves a 1-D-like problem along
// j (log E) whose condition number grows like ny^2.  Coarsening j only, with whole x-lines as the
// smoother on every level, is the textbook cure for this kind of anisotropy: at 1024^2 BiCGSTAB needs
// 14-16 iterations per time step instead of 290 (full lines) / 390 (16-row segments); prototype and
// measurements in profiles/proto_mg.py.
//
// Levels l = 0..L-1, level l has nx x (ny >> l) cells, all stored like the fine grid ([i][j], j
// fastest).  Every level holds a UNIT-DIAGONAL 5-point operator A_l = (wW, wE, wS, wN) plus the row
// weight om_l with B_l = diag(om_l) A_l the "conservative" form (level 0: om = 1, B_0 = A).
// Coarsening by aggregating the pair (i,2J), (i,2J+1):
//   B_{l+1} = P^T B_l P   (piecewise-constant P, plain Galerkin)  for the W/E couplings and the diagonal,
//   the S/N couplings between aggregates are multiplied by theta = 1/2 and the same amount is taken
//   off the diagonal of the column they sit in (column sums kept): for a 1-D Laplacian this is exactly
//   the rediscretisation on cells twice as high; plain Galerkin with constant P would be 2x too stiff.
// V(1,1) cycle, damping kOmega:
//   z  = omega T^-1 r                                   k_mg_line<MODE 0>
//   rc = P^T om (r - A z) / om_c                        k_mg_resid<restrict>
//   zc = V(rc)
//   t  = r - A (z + P zc)                               k_mg_resid<prolong>
//   z  = z + P zc + omega T^-1 t                        k_mg_line<MODE 1>
// coarsest level (<= 64 columns by default): kMgCoarseSweeps damped line-Jacobi sweeps.
// T = tridiag(wW, 1, wE) along i over the WHOLE column, solved exactly by a partitioned Thomas
// algorithm: a thread owns kMgSeg consecutive rows of one column, runs both recurrences in registers
// with carry-in 0, the per-segment affine maps (carry -> last value) are composed by a warp-shuffle
// scan over the segments of the column, and a correction pass adds carry * (prefix product).  A CTA
// covers all segments of kMgCols adjacent columns, so neighbouring threads touch adjacent addresses.
#pragma once
#include <type_traits>

#include "sy2d_kernels.cuh"

namespace sy2d {

constexpr int kMgMaxLevels = 8;
constexpr double kMgOmega = 0.7;   // line-Jacobi damping
constexpr double kMgTheta = 0.5;   // rescaling of the inter-aggregate couplings
constexpr int kMgCoarseSweeps = 2;  // default number of smoothing sweeps on the coarsest level

struct MgLevel {
  const double *wW, *wE, *wS, *wN;  // unit-diagonal operator
  const double* om;                 // row weights (nullptr on level 0: all ones)
  double *l, *dinv, *e;             // LU of the x-lines: l_i = wW_i/d_{i-1}, 1/d_i, wE_i/d_i
  const double* r;                  // right-hand side of the level (level 0: the vector to precondition)
  double *z, *t;                    // solution of the level, work vector
  int ny;                           // columns of this level
  size_t N;                         // nx * ny
};

struct MgArgs {
  const Scal* scal;   // converged problems skip all work
  int nx;             // rows of the (local) grid the kernels work on
  int halo;           // row-slab mode: the level arrays carry one halo row below row 0 and one above row nx-1 (the
                      // pointers address the first OWNED row), filled by the host with the neighbour ranks' rows:
                      // the residual kernels read them instead of clamping.  The line solves never look at them -
                      // The line kernel solves the rank's OWN rows; the coupling to the neighbour ranks' rows is
                      // restored afterwards by the spike correction below (exact global lines).
  double* tips;       // row-slab mode: [2][ny] - first and last value of every local line solution (before damping)
  // Fused Krylov update (first line solve of a V-cycle, MODE 0 on level 0, single GPU): the right-hand side of the solve is FORMED
  // here instead of by a kernel of its own and stored for the kernels that follow -
  //   fuse = 1: p = first ? rhs : r + beta (p - omega v)     (k_p_update2)      fuse = 2: s = (first ? rhs : r) - alpha v     (k_s_update2)
  // with the scalars of the problem's Scal; f_dst = the array the solve otherwise reads (lv.r).
  int fuse = 0;
  const double* f_rhs = nullptr;
  const double* f_r = nullptr;
  const double* f_v = nullptr;
  double* f_dst = nullptr;
};

// ---------------------------------------------------------------------------------------------
// Setup, once per time step
// ---------------------------------------------------------------------------------------------

// Coarse operator of level l+1 from level l.  One thread per coarse cell (i, J).
__global__ void __launch_bounds__(kBlock) k_mg_coarsen(MgLevel f, double* __restrict__ cW, double* __restrict__ cE,
                                                       double* __restrict__ cS, double* __restrict__ cN,
                                                       double* __restrict__ cOm, int nx) {
  const int nyf = f.ny, nyc = nyf >> 1;
  const size_t Nc = (size_t)nx * nyc;
  const size_t bf = (size_t)blockIdx.y * f.N, bc = (size_t)blockIdx.y * Nc;
  for (size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x; n < Nc; n += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(n / nyc), J = (int)(n - (size_t)i * nyc);
    const size_t a = bf + (size_t)i * nyf + 2 * J;   // fine cells a = (i, 2J), b = a + 1
    const double2 wW = ld2(f.wW + a), wE = ld2(f.wE + a), wS = ld2(f.wS + a), wN = ld2(f.wN + a);
    double2 om = make_double2(1.0, 1.0);
    if (f.om) om = ld2(f.om + a);
    // entries of the neighbouring aggregates that sit in this aggregate's column
    double sUp = 0.0, nDn = 0.0;   // row (J+1) col J: om wS of cell (i, 2J+2);  row (J-1) col J: om wN of cell (i, 2J-1)
    if (J + 1 < nyc) sUp = (f.om ? f.om[a + 2] : 1.0) * f.wS[a + 2];
    if (J > 0) nDn = (f.om ? f.om[a - 1] : 1.0) * f.wN[a - 1];
    const double d = om.x * (1.0 + wN.x) + om.y * (1.0 + wS.y) + (1.0 - kMgTheta) * (sUp + nDn);
    const double inv = sy2d_div(1.0, d);
    cW[bc + n] = (om.x * wW.x + om.y * wW.y) * inv;
    cE[bc + n] = (om.x * wE.x + om.y * wE.y) * inv;
    cS[bc + n] = kMgTheta * om.x * wS.x * inv;
    cN[bc + n] = kMgTheta * om.y * wN.y * inv;
    cOm[bc + n] = d;
  }
}

struct MgLevels {
  MgLevel lv[kMgMaxLevels];
  int nlev;
};

// LU of every x-line of every level: one thread per (level, column), sequential along i.  The
// recurrence d_i = 1 - wW_i wE_{i-1} / d_{i-1} has one reciprocal on its critical path; the
// coefficients of kFactorChunk rows are loaded ahead of the recurrence (the loads do not depend on it).
constexpr int kFactorChunk = 16;
__global__ void __launch_bounds__(64) k_mg_factor(MgLevels L, int nx) {
  const MgLevel& lv = L.lv[blockIdx.z];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= lv.ny) return;
  const size_t base = (size_t)blockIdx.y * lv.N + j;
  const int ny = lv.ny;
  double dinv_prev = 0.0, wE_prev = 0.0;
  for (int i0 = 0; i0 < nx; i0 += kFactorChunk) {
    double w[kFactorChunk], e[kFactorChunk];
#pragma unroll
    for (int m = 0; m < kFactorChunk; ++m) {
      const int i = min(i0 + m, nx - 1);
      w[m] = lv.wW[base + (size_t)i * ny];
      e[m] = lv.wE[base + (size_t)i * ny];
    }
#pragma unroll
    for (int m = 0; m < kFactorChunk; ++m) {
      if (i0 + m < nx) {
        const size_t n = base + (size_t)(i0 + m) * ny;
        const double l = w[m] * dinv_prev;          // 0 on the first row (wW = 0 there as well)
        const double d = 1.0 - l * wE_prev;
        const double dinv = sy2d_div(1.0, d);
        lv.l[n] = l;
        lv.dinv[n] = dinv;
        lv.e[n] = e[m] * dinv;
        dinv_prev = dinv;
        wE_prev = e[m];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Whole-column Thomas solves, partitioned over the threads of a CTA
// ---------------------------------------------------------------------------------------------

// Carry-in of every segment of the columns of this CTA from the per-segment affine maps
// carry_out = A[s] + P[s] * carry_in, s = 0..nseg-1 chained in index order (the backward sweep stores its
// segments in reversed order).  Shared layout [col][stride] with stride = nseg + 1, so a warp reads 32
// consecutive segments of one column without bank conflicts.  Warp w scans column w, w + nwarps, ...:
// blocks of 32 segments, a shuffle scan of the affine maps inside a block, the total carried on to the
// next block.  Contains no barrier; every thread of the CTA calls it.
template <int COLS>
__device__ __forceinline__ void mg_carry_scan(const double* sA, const double* sP, double* sC, int nseg, int stride) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int col = w; col < COLS; col += nw) {
    double cb = 0.0;   // carry into the current block of 32 segments
    for (int s0 = 0; s0 < nseg; s0 += 32) {
      const int s = s0 + lane;
      double A = 0.0, P = 1.0;   // identity for lanes past the end
      if (s < nseg) { A = sA[col * stride + s]; P = sP[col * stride + s]; }
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {   // inclusive scan: map of segments s0 .. s applied in order
        const double Ao = __shfl_up_sync(0xffffffffu, A, o), Po = __shfl_up_sync(0xffffffffu, P, o);
        if (lane >= o) { A = A + P * Ao; P = P * Po; }
      }
      double Ae = __shfl_up_sync(0xffffffffu, A, 1), Pe = __shfl_up_sync(0xffffffffu, P, 1);
      if (lane == 0) { Ae = 0.0; Pe = 1.0; }
      if (s < nseg) sC[col * stride + s] = Ae + Pe * cb;
      const double At = __shfl_sync(0xffffffffu, A, 31), Pt = __shfl_sync(0xffffffffu, P, 31);
      cb = At + Pt * cb;
    }
  }
}

// The same scan with ONE (column, block of 32 segments) per warp - all warps of the CTA busy for one shuffle scan instead of
// COLS warps for nseg / 32 of them in a row (the chain a latency-bound launch waits for).  Needs nseg % 32 == 0 and
// COLS * nseg / 32 warps.  (sC, sQ)[col][s] = the map from the carry into the BLOCK to the carry into segment s,
// (tA, tP)[col][blk] = the composed map of a whole block; mg_block_carry chains the blocks before the reader's own.
template <int COLS>
__device__ __forceinline__ void mg_carry_scan_blocks(const double* sA, const double* sP, double* sC, double* sQ, double* tA, double* tP,
                                                     int stride, int nblk) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = w % COLS, blk = w / COLS;
  const int s = blk * 32 + lane;
  double A = sA[col * stride + s], P = sP[col * stride + s];
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double Ao = __shfl_up_sync(0xffffffffu, A, o), Po = __shfl_up_sync(0xffffffffu, P, o);
    if (lane >= o) { A = A + P * Ao; P = P * Po; }
  }
  double Ae = __shfl_up_sync(0xffffffffu, A, 1), Pe = __shfl_up_sync(0xffffffffu, P, 1);
  if (lane == 0) { Ae = 0.0; Pe = 1.0; }
  sC[col * stride + s] = Ae;
  sQ[col * stride + s] = Pe;
  if (lane == 31) { tA[col * nblk + blk] = A; tP[col * nblk + blk] = P; }
}
__device__ __forceinline__ double mg_block_carry(const double* sC, const double* sQ, const double* tA, const double* tP, int col, int s,
                                                 int stride, int nblk) {
  const int blk = s >> 5;
  double cb = 0.0;
  for (int b2 = 0; b2 < blk; ++b2) cb = tA[col * nblk + b2] + tP[col * nblk + b2] * cb;
  return sC[col * stride + s] + sQ[col * stride + s] * cb;
}

__device__ __forceinline__ void mg_cp_async8(double* smem_dst, const double* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void mg_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// MODE 0: z = omega T^-1 r
// MODE 1: z = z + P zc + omega T^-1 t      (zc: solution of the next coarser level)
// MODE 2: z = z + omega T^-1 t
// grid: (ceil(ny / COLS), nbatch), block: nseg * COLS threads rounded up to whole warps, nseg = ceil(nx / SEG).
// FULL: nx % SEG == 0, ny % COLS == 0 and nseg * COLS % 32 == 0 - no thread and no row needs a predicate.
// The kernel is latency bound (a CTA is one dependent chain of load - sweep - scan - sweep - scan - store
// and a level has at most ny / COLS CTAs), so the instruction count per thread is what matters: short
// segments, running pointers instead of index products, all loads of a phase issued before their use.
// `group` = which COLS adjacent columns, `batch` = which problem (the stand-alone kernel passes blockIdx.x / .y; the fused
// coarse-level kernel below loops over groups).  `state` != 0: the problem has converged, nothing to do.
// PRE (FULL shapes with nseg % 32 == 0; the host adds the shared memory): the solve as ONE exposed memory latency instead of
// three - the factors of the backward sweep (dinv, e) travel by cp.async into thread-private shared-memory slots while the forward
// sweep runs, the old iterate (and the coarse correction) follow into the same slots during the backward sweep - and the scans run one block of 32
// segments per warp (mg_carry_scan_blocks).
template <int SEG, int COLS, int MODE, bool FULL, bool PRE = false>
__device__ __forceinline__ void mg_line_body(const MgLevel& lv, const double* zc, const MgArgs& a, int group, int batch, int state, double* mg_smem) {
  static_assert(!PRE || FULL, "PRE needs a shape without predicates");
  const int nx = a.nx, ny = lv.ny;
  const int nseg = (nx + SEG - 1) / SEG;
  const int stride = nseg + 1;
  const int nblk = nseg >> 5;
  double* sA = mg_smem;
  double* sP = sA + COLS * stride;
  double* sC = sP + COLS * stride;
  double* sQ = sC + COLS * stride;            // PRE only from here on
  double* tA = sQ + COLS * stride;
  double* tP = tA + COLS * nblk;
  double* sD = tP + COLS * nblk;              // [SEG][threads]: dinv, then e
  double* sE = sD + SEG * (int)blockDim.x;
  const int col = threadIdx.x % COLS, seg = threadIdx.x / COLS;
  const int j = group * COLS + col;
  const bool in_cta = FULL || seg < nseg;
  const bool live = FULL || (in_cta && j < ny);
  const int r0 = seg * SEG;
  const int cnt = FULL ? SEG : (live ? min(SEG, nx - r0) : 0);
  const size_t n0 = live ? (size_t)batch * lv.N + (size_t)r0 * ny + j : 0;
  double y[SEG], c[SEG];
  // forward sweep y_m = b_m - l_m y_{m-1} with carry-in 0; P = product of (-l) over the segment
  if (MODE == 0 && a.fuse != 0) {   // the right-hand side is the Krylov update itself (see MgArgs)
    const Scal& sc = a.scal[batch];
    const bool first = sc.first != 0;
    const double alpha = sc.alpha, beta = sc.beta, omega = sc.omega;
    const double* pr = (first ? a.f_rhs : a.f_r) + n0;
    const double* pv = a.f_v + n0;
    const double* pp = a.f_dst + n0;
    const double* pl = lv.l + n0;
#pragma unroll
    for (int m = 0; m < SEG; ++m) {
      y[m] = 0.0; c[m] = 0.0;
      if (FULL || m < cnt) {
        const double r = *pr;
        if (a.fuse == 1) y[m] = first ? r : r + beta * (*pp - omega * *pv);
        else y[m] = r - alpha * *pv;
        c[m] = *pl;
      }
      pr += ny; pv += ny; pp += ny; pl += ny;
    }
    if (state != 0) return;
    double* pd = a.f_dst + n0;
#pragma unroll
    for (int m = 0; m < SEG; ++m) {
      if (FULL || m < cnt) *pd = y[m];
      pd += ny;
    }
  } else {
    const double* ps = (MODE == 0 ? lv.r : lv.t) + n0;
    const double* pl = lv.l + n0;
#pragma unroll
    for (int m = 0; m < SEG; ++m) {
      y[m] = 0.0; c[m] = 0.0;
      if (FULL || m < cnt) { y[m] = *ps; c[m] = *pl; }
      ps += ny; pl += ny;
    }
  }
  if (state != 0) return;   // uniform over the CTA
  if (PRE) {
    const double* pd = lv.dinv + n0;
    const double* pe = lv.e + n0;
#pragma unroll
    for (int m = 0; m < SEG; ++m) {
      mg_cp_async8(sD + m * (int)blockDim.x + threadIdx.x, pd);
      mg_cp_async8(sE + m * (int)blockDim.x + threadIdx.x, pe);
      pd += ny; pe += ny;
    }
  }
  double carry = 0.0, P = 1.0, last = 0.0;
#pragma unroll
  for (int m = 0; m < SEG; ++m) {
    if (FULL || m < cnt) {
      carry = y[m] - c[m] * carry;
      y[m] = carry;
      P = -c[m] * P;
      last = carry;
    }
  }
  if (in_cta) {
    sA[col * stride + seg] = last;
    sP[col * stride + seg] = (FULL || cnt > 0) ? P : 0.0;
  }
  __syncthreads();
  if (PRE) mg_carry_scan_blocks<COLS>(sA, sP, sC, sQ, tA, tP, stride, nblk);
  else mg_carry_scan<COLS>(sA, sP, sC, nseg, stride);
  __syncthreads();
  {
    double q = PRE ? mg_block_carry(sC, sQ, tA, tP, col, seg, stride, nblk) : (in_cta ? sC[col * stride + seg] : 0.0);
#pragma unroll
    for (int m = 0; m < SEG; ++m) { q = -c[m] * q; y[m] += q; }   // rows past cnt have c = 0
  }
  // backward sweep z_m = y_m / d_m - e_m z_{m+1} with carry-in 0; P = product of (-e)
  if (PRE) {
    mg_cp_async_wait_all();   // the thread's own copies: no barrier needed
#pragma unroll
    for (int m = 0; m < SEG; ++m) {
      y[m] *= sD[m * (int)blockDim.x + threadIdx.x];
      c[m] = sE[m * (int)blockDim.x + threadIdx.x];
    }
    if (MODE != 0) {   // the slots are free again: the old iterate (and the coarse correction) of the update travel next
      const double* pq = lv.z + n0;
      const double* pc = MODE == 1 ? zc + ((size_t)batch * (lv.N >> 1) + (size_t)r0 * (ny >> 1) + (j >> 1)) : nullptr;
#pragma unroll
      for (int m = 0; m < SEG; ++m) {
        mg_cp_async8(sD + m * (int)blockDim.x + threadIdx.x, pq);
        if (MODE == 1) mg_cp_async8(sE + m * (int)blockDim.x + threadIdx.x, pc);
        pq += ny;
        if (MODE == 1) pc += ny >> 1;
      }
    }
  } else {
    const double* pd = lv.dinv + n0;
    const double* pe = lv.e + n0;
    double dv[SEG];
#pragma unroll
    for (int m = 0; m < SEG; ++m) {
      dv[m] = 0.0; c[m] = 0.0;
      if (FULL || m < cnt) { dv[m] = *pd; c[m] = *pe; }
      pd += ny; pe += ny;
    }
#pragma unroll
    for (int m = 0; m < SEG; ++m) y[m] *= dv[m];
  }
  carry = 0.0; P = 1.0;
#pragma unroll
  for (int m = SEG - 1; m >= 0; --m) {
    if (FULL || m < cnt) {
      carry = y[m] - c[m] * carry;
      y[m] = carry;
      P = -c[m] * P;
    }
  }
  __syncthreads();   // sA/sP/sC are reused; segments are stored in reversed order: the scan runs last to first
  if (in_cta) {
    sA[col * stride + (nseg - 1 - seg)] = y[0];
    sP[col * stride + (nseg - 1 - seg)] = (FULL || cnt > 0) ? P : 0.0;
  }
  __syncthreads();
  if (PRE) mg_carry_scan_blocks<COLS>(sA, sP, sC, sQ, tA, tP, stride, nblk);
  else mg_carry_scan<COLS>(sA, sP, sC, nseg, stride);
  __syncthreads();
  {
    double q = PRE ? mg_block_carry(sC, sQ, tA, tP, col, nseg - 1 - seg, stride, nblk) : (in_cta ? sC[col * stride + (nseg - 1 - seg)] : 0.0);
#pragma unroll
    for (int m = SEG - 1; m >= 0; --m) { q = -c[m] * q; y[m] += q; }
  }
  if (a.tips && live) {   // interface values of the local solution (spike correction, row-slab mode)
    if (r0 == 0) a.tips[j] = y[0];
    if (r0 + cnt == nx) {
      double yl = y[0];
#pragma unroll
      for (int m = 1; m < SEG; ++m) yl = (m < cnt) ? y[m] : yl;
      a.tips[ny + j] = yl;
    }
  }
  // all loads of the update before the first store (z is read and written through the same pointer)
  double* pz = lv.z + n0;
  if (PRE && MODE != 0) {
    mg_cp_async_wait_all();
#pragma unroll
    for (int m = 0; m < SEG; ++m) {
      c[m] = sD[m * (int)blockDim.x + threadIdx.x];
      if (MODE == 1) c[m] += sE[m * (int)blockDim.x + threadIdx.x];
    }
  } else if (MODE != 0) {
    const double* pq = pz;
    const double* pc = MODE == 1 ? zc + ((size_t)batch * (lv.N >> 1) + (size_t)r0 * (ny >> 1) + (j >> 1)) : nullptr;
#pragma unroll
    for (int m = 0; m < SEG; ++m) {
      c[m] = 0.0;
      if (FULL || m < cnt) {
        c[m] = *pq;
        if (MODE == 1) c[m] += *pc;
      }
      pq += ny;
      if (MODE == 1) pc += ny >> 1;
    }
  }
#pragma unroll
  for (int m = 0; m < SEG; ++m) {
    if (FULL || m < cnt) *pz = MODE == 0 ? kMgOmega * y[m] : c[m] + kMgOmega * y[m];
    pz += ny;
  }
}

template <int SEG, int COLS, int MODE, bool FULL, bool PRE = false>
__global__ void __launch_bounds__(SEG <= 8 ? 1024 : 512, 1) k_mg_line(MgLevel lv, const double* __restrict__ zc, MgArgs a) {
  extern __shared__ double mg_smem[];
  mg_line_body<SEG, COLS, MODE, FULL, PRE>(lv, zc, a, (int)blockIdx.x, (int)blockIdx.y, a.scal[blockIdx.y].state, mg_smem);
}

// ---------------------------------------------------------------------------------------------
// Long columns: one column group per thread-block CLUSTER (sm_90+ clusters, distributed shared memory).
//
// A thread's registers limit a CTA to 512 threads of 16 rows: beyond 2048 rows the stand-alone kernel above can only
// keep the whole column in one CTA by taking fewer columns (4096 rows: 2, 8192 rows: 1), and with one column per CTA
// a row access is 8 bytes of a 32-byte sector - the line solves of a 16384^2 grid on 2 GPUs moved 4x their bytes.
// Here the CL CTAs of a cluster split the SEGMENTS of the same 8 columns (64-byte row accesses for any column length up
// to 8192 rows): every CTA scans the affine maps of its own segments as before, publishes the composed map of its
// whole part per column in its shared memory, and after a cluster barrier reads the parts below (forward sweep) or
// above (backward sweep) through DSMEM to get its carry-in: three cluster barriers per solve, one pass over memory.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ double ld_dsmem(const double* local_smem_ptr, unsigned rank) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(local_smem_ptr);
  unsigned ra;
  double v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra) : "memory");
  return v;
}

// Like mg_carry_scan, for the segments of ONE CTA of the cluster: sC = carry-in of a segment if the carry into the CTA's
// first segment were 0, sQ = the factor a non-zero carry into the CTA is multiplied by on its way to the segment, and
// (totA, totP)[col] = the composed map of the CTA's whole part.
template <int COLS>
__device__ __forceinline__ void mg_carry_scan_part(const double* sA, const double* sP, double* sC, double* sQ, double* totA, double* totP,
                                                   int nseg, int stride) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int col = w; col < COLS; col += nw) {
    double cb = 0.0, pb = 1.0;
    for (int s0 = 0; s0 < nseg; s0 += 32) {
      const int s = s0 + lane;
      double A = 0.0, P = 1.0;
      if (s < nseg) { A = sA[col * stride + s]; P = sP[col * stride + s]; }
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double Ao = __shfl_up_sync(0xffffffffu, A, o), Po = __shfl_up_sync(0xffffffffu, P, o);
        if (lane >= o) { A = A + P * Ao; P = P * Po; }
      }
      double Ae = __shfl_up_sync(0xffffffffu, A, 1), Pe = __shfl_up_sync(0xffffffffu, P, 1);
      if (lane == 0) { Ae = 0.0; Pe = 1.0; }
      if (s < nseg) { sC[col * stride + s] = Ae + Pe * cb; sQ[col * stride + s] = Pe * pb; }
      const double At = __shfl_sync(0xffffffffu, A, 31), Pt = __shfl_sync(0xffffffffu, P, 31);
      cb = At + Pt * cb;
      pb = Pt * pb;
    }
    if (lane == 0) { totA[col] = cb; totP[col] = pb; }
  }
}

// grid: (groups * CL, nbatch) with cluster dimension (CL, 1, 1); block: (nseg / CL) * COLS threads.
// Requires nx % (SEG * CL) == 0, ny % COLS == 0, ((nseg / CL) * COLS) % 32 == 0 (the host falls back to k_mg_line otherwise).
template <int SEG, int COLS, int MODE, int CL>
__global__ void __launch_bounds__(512, 1) k_mg_line_cluster(MgLevel lv, const double* zc, MgArgs a) {
  extern __shared__ double mg_smem[];
  const int state = a.scal[blockIdx.y].state;
  const int nx = a.nx, ny = lv.ny;
  const int nseg_c = nx / (SEG * CL);          // segments of this CTA
  const int stride = nseg_c + 1;
  double* sA = mg_smem;
  double* sP = sA + COLS * stride;
  double* sC = sP + COLS * stride;
  double* sQ = sC + COLS * stride;
  double* tot = sQ + COLS * stride;            // [2 sweeps][A, P][COLS]: read by the other CTAs of the cluster
  const unsigned q = cluster_ctarank();
  const int group = blockIdx.x / CL;
  const int col = threadIdx.x % COLS, seg = threadIdx.x / COLS;
  const int j = group * COLS + col;
  const int r0 = ((int)q * nseg_c + seg) * SEG;
  const size_t n0 = (size_t)blockIdx.y * lv.N + (size_t)r0 * ny + j;
  double y[SEG], c[SEG];
  {
    const double* ps = (MODE == 0 ? lv.r : lv.t) + n0;
    const double* pl = lv.l + n0;
#pragma unroll
    for (int m = 0; m < SEG; ++m) { y[m] = *ps; c[m] = *pl; ps += ny; pl += ny; }
  }
  if (state != 0) return;   // uniform over the cluster (one problem per cluster): nobody reaches a cluster barrier
  double carry = 0.0, P = 1.0;
#pragma unroll
  for (int m = 0; m < SEG; ++m) { carry = y[m] - c[m] * carry; y[m] = carry; P = -c[m] * P; }
  sA[col * stride + seg] = carry;
  sP[col * stride + seg] = P;
  __syncthreads();
  mg_carry_scan_part<COLS>(sA, sP, sC, sQ, tot, tot + COLS, nseg_c, stride);
  cluster_sync_all();                          // every part's map is published (and this CTA's scan is complete)
  {
    double cb = 0.0;                           // carry into this CTA: the parts below, in order
    for (unsigned r = 0; r < q; ++r) cb = ld_dsmem(tot + col, r) + ld_dsmem(tot + COLS + col, r) * cb;
    double qv = sC[col * stride + seg] + sQ[col * stride + seg] * cb;
#pragma unroll
    for (int m = 0; m < SEG; ++m) { qv = -c[m] * qv; y[m] += qv; }
  }
  {
    const double* pd = lv.dinv + n0;
    const double* pe = lv.e + n0;
    double dv[SEG];
#pragma unroll
    for (int m = 0; m < SEG; ++m) { dv[m] = *pd; c[m] = *pe; pd += ny; pe += ny; }
#pragma unroll
    for (int m = 0; m < SEG; ++m) y[m] *= dv[m];
  }
  carry = 0.0; P = 1.0;
#pragma unroll
  for (int m = SEG - 1; m >= 0; --m) { carry = y[m] - c[m] * carry; y[m] = carry; P = -c[m] * P; }
  __syncthreads();                             // sA .. sQ are reused (tot of the forward sweep stays: separate slots below)
  sA[col * stride + (nseg_c - 1 - seg)] = y[0];
  sP[col * stride + (nseg_c - 1 - seg)] = P;
  __syncthreads();
  mg_carry_scan_part<COLS>(sA, sP, sC, sQ, tot + 2 * COLS, tot + 3 * COLS, nseg_c, stride);
  cluster_sync_all();
  {
    double cb = 0.0;                           // carry into this CTA from the parts above, nearest last
    for (int r = CL - 1; r > (int)q; --r) cb = ld_dsmem(tot + 2 * COLS + col, (unsigned)r) + ld_dsmem(tot + 3 * COLS + col, (unsigned)r) * cb;
    double qv = sC[col * stride + (nseg_c - 1 - seg)] + sQ[col * stride + (nseg_c - 1 - seg)] * cb;
#pragma unroll
    for (int m = SEG - 1; m >= 0; --m) { qv = -c[m] * qv; y[m] += qv; }
  }
  if (a.tips) {   // interface values of the local solution (spike correction, row-slab mode)
    if (r0 == 0) a.tips[j] = y[0];
    if (r0 + SEG == nx) a.tips[ny + j] = y[SEG - 1];
  }
  {
    double* pz = lv.z + n0;
    if (MODE != 0) {
      const double* pq = pz;
      const double* pc = MODE == 1 ? zc + ((size_t)blockIdx.y * (lv.N >> 1) + (size_t)r0 * (ny >> 1) + (j >> 1)) : nullptr;
#pragma unroll
      for (int m = 0; m < SEG; ++m) {
        c[m] = *pq;
        if (MODE == 1) c[m] += *pc;
        pq += ny;
        if (MODE == 1) pc += ny >> 1;
      }
    }
#pragma unroll
    for (int m = 0; m < SEG; ++m) {
      *pz = MODE == 0 ? kMgOmega * y[m] : c[m] + kMgOmega * y[m];
      pz += ny;
    }
  }
  cluster_sync_all();                          // no CTA may exit while its shared memory can still be read
}

// ---------------------------------------------------------------------------------------------
// Row-slab mode: exact x-lines across ranks (SPIKE).  Rank r owns rows [0, n) of every global line; with
// T_r its own tridiagonal block, bot_{r-1} the line's value on the last row of rank r-1 and top_{r+1} the one on
// the first row of rank r+1,
//     x_r = g_r - W_r bot_{r-1} - V_r top_{r+1},    g_r = T_r^-1 b_r,
//     W_r = T_r^-1 (wW_0 e_0),  V_r = T_r^-1 (wE_{n-1} e_{n-1})          (the "spikes": once per time step).
// The interface values solve a 2P x 2P system per column whose data are the tips (first, last entry) of g, W, V
// of every rank: all-gathered (2 ny doubles per rank and solve), solved redundantly by every rank with a block
// forward elimination / back substitution over the ranks, then x_r is corrected in one streaming pass.  The
// preconditioner is then the one of the single-GPU engine up to round-off, so the iteration counts do not grow
// with the number of slabs.
// ---------------------------------------------------------------------------------------------
constexpr int kMgMaxRanks = 16;

// right-hand side of a spike solve: which = 0: wW of the first row on row 0; 1: wE of the last row on row nx-1
__global__ void __launch_bounds__(kBlock) k_mg_spike_rhs(double* __restrict__ t, const double* __restrict__ w, int which, int nx, int ny) {
  const size_t N = (size_t)nx * ny;
  const size_t hot = which == 0 ? 0 : (size_t)(nx - 1) * ny;
  for (size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (size_t)gridDim.x * blockDim.x)
    t[n] = (n >= hot && n < hot + ny) ? w[n] : 0.0;
}

// tips_all: [P][2][ny] (g first, g last); sp_all: [P][4][ny] (W first, W last, V first, V last);
// coef: [2][ny] = (bot_{rank-1}, top_{rank+1}) of the corrected lines.  One thread per column.
__global__ void __launch_bounds__(128) k_mg_spike_reduced(const double* __restrict__ tips_all, const double* __restrict__ sp_all,
                                                          double* __restrict__ coef, const Scal* scal, int rank, int P, int ny) {
  if (scal[0].state != 0) return;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ny) return;
  double al[kMgMaxRanks], be[kMgMaxRanks], ga[kMgMaxRanks], de[kMgMaxRanks];
  double ap = 0.0, bp = 0.0;
  for (int r = 0; r < P; ++r) {
    const double g0 = tips_all[((size_t)r * 2 + 0) * ny + j], g1 = tips_all[((size_t)r * 2 + 1) * ny + j];
    const double W0 = sp_all[((size_t)r * 4 + 0) * ny + j], W1 = sp_all[((size_t)r * 4 + 1) * ny + j];
    const double V0 = sp_all[((size_t)r * 4 + 2) * ny + j], V1 = sp_all[((size_t)r * 4 + 3) * ny + j];
    const double inv = 1.0 / (1.0 + W0 * bp);
    ga[r] = (g0 - W0 * ap) * inv;
    de[r] = -V0 * inv;
    al[r] = g1 - W1 * ap - W1 * bp * ga[r];
    be[r] = -W1 * bp * de[r] - V1;
    ap = al[r]; bp = be[r];
  }
  double tnext = 0.0, t_after_me = 0.0, b_before_me = 0.0;
  for (int r = P - 1; r >= 0; --r) {
    const double tr = ga[r] + de[r] * tnext;
    const double br = al[r] + be[r] * tnext;
    if (r == rank + 1) t_after_me = tr;
    if (r == rank - 1) b_before_me = br;
    tnext = tr;
  }
  coef[j] = b_before_me;
  coef[ny + j] = t_after_me;
}

// z -= spW * bot_{rank-1} + spV * top_{rank+1}   (spW, spV already carry the damping omega of the smoother)
__global__ void __launch_bounds__(kBlock) k_mg_spike_apply(double* __restrict__ z, const double* __restrict__ spW, const double* __restrict__ spV,
                                                           const double* __restrict__ coef, const Scal* scal, int nx, int ny) {
  if (scal[0].state != 0) return;
  const size_t N2 = (size_t)nx * ny / 2;
  const int ny2 = ny >> 1;
  for (size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x; n < N2; n += (size_t)gridDim.x * blockDim.x) {
    const int j2 = (int)(n % ny2);
    const double2 a = ld2(coef + 2 * j2), b = ld2(coef + ny + 2 * j2);
    const double2 w = ld2(spW + 2 * n), v = ld2(spV + 2 * n);
    double2 zz = ld2(z + 2 * n);
    zz.x -= w.x * a.x + v.x * b.x;
    zz.y -= w.y * a.y + v.y * b.y;
    st2(z + 2 * n, zz.x, zz.y);
  }
}

// ---------------------------------------------------------------------------------------------
// Residuals.  One thread per pair of cells (i, 2J), (i, 2J+1): 16-byte accesses, ny even.
//   KIND 0: t = r - A z                                          (coarsest-level sweeps)
//   KIND 1: rc(i,J) = (om_a res_a + om_b res_b) / om_c(i,J)       (restriction to the next level)
//   KIND 2: t = r - A (z + P zc)                                  (after the coarse-grid correction)
// Out-of-range neighbours are clamped: their weights are exactly zero.
// ---------------------------------------------------------------------------------------------
template <int KIND>
__device__ __forceinline__ void mg_resid_body(const MgLevel& lv, const double* zc, const double* omc, double* rc, const MgArgs& a, int batch,
                                              size_t first, size_t stride) {
  const int nx = a.nx, ny = lv.ny, nyc = ny >> 1;
  const size_t Nc = (size_t)nx * nyc;
  const size_t bf = (size_t)batch * lv.N, bc = (size_t)batch * Nc;
  const double* z = lv.z + bf;
  const double* zcp = KIND == 2 ? zc + bc : nullptr;
  for (size_t n = first; n < Nc; n += stride) {
    const int i = (int)(n / nyc), J = (int)(n - (size_t)i * nyc);
    const ptrdiff_t p = (ptrdiff_t)i * ny + 2 * J;    // in-problem index of cell a
    const ptrdiff_t pW = (i > 0 || a.halo) ? p - ny : p, pE = (i < nx - 1 || a.halo) ? p + ny : p;
    const ptrdiff_t pS = J > 0 ? p - 1 : p, pN = J < nyc - 1 ? p + 2 : p + 1;
    double2 zz = ld2(z + p), zW = ld2(z + pW), zE = ld2(z + pE);
    double zS = z[pS], zN = z[pN];
    if (KIND == 2) {
      const ptrdiff_t q = (ptrdiff_t)n;
      const double c0 = zcp[q];
      zz.x += c0; zz.y += c0;
      const double cW = zcp[(i > 0 || a.halo) ? q - nyc : q], cE = zcp[(i < nx - 1 || a.halo) ? q + nyc : q];
      zW.x += cW; zW.y += cW; zE.x += cE; zE.y += cE;
      zS += zcp[J > 0 ? n - 1 : n];
      zN += zcp[J < nyc - 1 ? n + 1 : n];
    }
    const double2 wW = ld2(lv.wW + bf + p), wE = ld2(lv.wE + bf + p), wS = ld2(lv.wS + bf + p), wN = ld2(lv.wN + bf + p);
    const double2 r = ld2(lv.r + bf + p);
    const double ra = r.x - (zz.x + ((wW.x * zW.x + wE.x * zE.x) + (wS.x * zS + wN.x * zz.y)));
    const double rb = r.y - (zz.y + ((wW.y * zW.y + wE.y * zE.y) + (wS.y * zz.x + wN.y * zN)));
    if (KIND == 1) {
      double2 om = make_double2(1.0, 1.0);
      if (lv.om) om = ld2(lv.om + bf + p);
      rc[bc + n] = sy2d_div(om.x * ra + om.y * rb, omc[bc + n]);
    } else {
      st2(lv.t + bf + p, ra, rb);
    }
  }
}

template <int KIND>
__global__ void __launch_bounds__(kBlock) k_mg_resid(MgLevel lv, const double* __restrict__ zc, const double* __restrict__ omc,
                                                     double* __restrict__ rc, MgArgs a) {
  if (a.scal[blockIdx.y].state != 0) return;
  mg_resid_body<KIND>(lv, zc, omc, rc, a, (int)blockIdx.y, (size_t)blockIdx.x * blockDim.x + threadIdx.x, (size_t)gridDim.x * blockDim.x);
}

// ---------------------------------------------------------------------------------------------
// The coarse tail of the V-cycle in ONE launch.  On the levels with few columns every kernel above is a latency-bound
// launch of 5-10 us whatever its size (a level has at most ny_l / COLS line CTAs, each one dependent chain), and a
// V-cycle spends 11 of its 19 launches on the three coarsest levels, which hold 7 / 16 of the fine grid's cells.
// k_mg_tail runs levels k0 .. L-1 - down sweep, coarsest-level sweeps, up sweep - as the stages of one kernel separated
// by a barrier over the CTAs of a problem (an atomic counter per problem; the grid is at most one CTA per SM, so all CTAs
// are resident).  A CTA has SEG x COLS-shaped work like the stand-alone line kernel (nseg x COLS threads) and loops over
// the column groups of a level; the residual stages are grid-stride over the same threads.  Plain loads only: arrays
// written in one stage are read in the next (no ld.global.nc), and the barrier's acquire invalidates the SM's L1.
// ---------------------------------------------------------------------------------------------
struct MgTailArgs {
  MgLevels L;
  MgArgs a;
  unsigned* barrier;   // [nbatch] arrival counters, zero between launches
  int k0;              // first fused level; its right-hand side lv[k0].r is ready, its iterate lv[k0].z is the result
  int coarse_sweeps;
};

__device__ __forceinline__ void mg_problem_barrier(unsigned* ctr, unsigned nblocks, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += nblocks;
    __threadfence();
    atomicAdd(ctr, 1u);
    unsigned seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
    } while (seen < target);
  }
  __syncthreads();
}

template <int SEG, int COLS>
__global__ void __launch_bounds__(512, 1) k_mg_tail(MgTailArgs t) {   // nseg x COLS <= 512 threads (the host picks COLS)
  extern __shared__ double mg_smem[];
  const int batch = blockIdx.y;
  if (t.a.scal[batch].state != 0) return;   // the whole problem is skipped by all of its CTAs alike
  unsigned* ctr = t.barrier + batch;
  unsigned target = 0;
  const unsigned nblk = gridDim.x;
  const int Lc = t.L.nlev;
  const size_t first = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  auto line = [&](auto mode, const MgLevel& lv, const double* zc) {
    constexpr int MODE = decltype(mode)::value;
    const int groups = lv.ny / COLS;
    for (int g = blockIdx.x; g < groups; g += gridDim.x) {
      mg_line_body<SEG, COLS, MODE, true>(lv, zc, t.a, g, batch, 0, mg_smem);
      __syncthreads();   // the scan buffers are reused by the next group
    }
  };
  for (int k = t.k0; k + 1 < Lc; ++k) {
    line(std::integral_constant<int, 0>{}, t.L.lv[k], nullptr);
    mg_problem_barrier(ctr, nblk, target);
    mg_resid_body<1>(t.L.lv[k], nullptr, t.L.lv[k + 1].om, const_cast<double*>(t.L.lv[k + 1].r), t.a, batch, first, stride);
    mg_problem_barrier(ctr, nblk, target);
  }
  line(std::integral_constant<int, 0>{}, t.L.lv[Lc - 1], nullptr);
  for (int sweep = 1; sweep < t.coarse_sweeps; ++sweep) {
    mg_problem_barrier(ctr, nblk, target);
    mg_resid_body<0>(t.L.lv[Lc - 1], nullptr, nullptr, nullptr, t.a, batch, first, stride);
    mg_problem_barrier(ctr, nblk, target);
    line(std::integral_constant<int, 2>{}, t.L.lv[Lc - 1], nullptr);
  }
  for (int k = Lc - 2; k >= t.k0; --k) {
    mg_problem_barrier(ctr, nblk, target);
    mg_resid_body<2>(t.L.lv[k], t.L.lv[k + 1].z, nullptr, nullptr, t.a, batch, first, stride);
    mg_problem_barrier(ctr, nblk, target);
    line(std::integral_constant<int, 1>{}, t.L.lv[k], t.L.lv[k + 1].z);
  }
  // exit arrival: the last CTA of the problem to get here leaves the counter at zero for the next launch
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned old = atomicAdd(ctr, 1u);
    if (old == target + nblk - 1) *ctr = 0u;
  }
}

// ---------------------------------------------------------------------------------------------
// Fine-level BiCGSTAB kernels of the right-preconditioned iteration
//   phat = M^-1 p, v = A phat, s = r - alpha v, shat = M^-1 s, t = A shat,
//   x += alpha phat + omega shat, r = s - omega t.
// p-update, v = A phat and s-update are the Jacobi kernels (k_p_update2, k_spmv_v2 on phat, k_s_update2).
// ---------------------------------------------------------------------------------------------

// t = A shat; (t, s), (t, t); last block: omega                        56 B/cell
__global__ void __launch_bounds__(kBlock, 6) k_mg_spmv_t(KrylovVecs k, const double* __restrict__ shat, size_t N, int ny) {
  __shared__ double red[2 * 32];
  Scal* sc = k.scal + blockIdx.y;
  if (sc->state != 0) return;
  const size_t stride = 2 * (size_t)gridDim.x * blockDim.x;
  const size_t base = (size_t)blockIdx.y * N;
  double ts = 0.0, tt = 0.0;
  for (size_t n = k.n_begin + 2 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x); n < k.n_end; n += stride) {
    const size_t g = base + n;
    const double* h = shat + base;
    double t0, t1;
    stencil_apply2(h, n, N, ny, ld2(h + n), ld2(k.wW + g), ld2(k.wE + g), ld2(k.wS + g), ld2(k.wN + g), t0, t1);
    st2(k.t + g, t0, t1);
    const double2 s = ld2(k.s + g);
    ts += t0 * s.x + t1 * s.y;
    tt += t0 * t0 + t1 * t1;
  }
  double sums[2] = {ts, tt};
  block_sums<2>(sums, red);
  {
    __shared__ int last_flag;
    double* const dst[2] = {&sc->acc_ts, &sc->acc_tt};
    if (cta_totals<2>(sc, k.part + (size_t)blockIdx.y * k.part_stride, sums, dst, nullptr, 0.0, &last_flag) && !k.defer) {
      const double a = sc->acc_ts, b = sc->acc_tt;
      sc->acc_ts = 0.0;
      sc->acc_tt = 0.0;
      sc->omega = b > 0.0 ? a / b : 0.0;
    }
  }
}

// x += alpha phat + omega shat; r = s - omega t; (rhat, r), max|r|; bookkeeping        72 B/cell
__global__ void __launch_bounds__(kBlock, 6) k_mg_xr(KrylovVecs k, const double* __restrict__ phat, const double* __restrict__ shat,
                                                  size_t N) {
  __shared__ double red[32];
  Scal* sc = k.scal + blockIdx.y;
  if (sc->state != 0) return;
  const size_t stride = 2 * (size_t)gridDim.x * blockDim.x;
  double dot = 0.0, rabs = 0.0;
  const double alpha = sc->alpha, omega = sc->omega;
  const bool first = sc->first;
  for (size_t n = k.n_begin + 2 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x); n < k.n_end; n += stride) {
    const size_t g = (size_t)blockIdx.y * N + n;
    const double2 s = ld2(k.s + g), ph = ld2(phat + g), sh = ld2(shat + g), t = ld2(k.t + g), rh = ld2(k.rhs + g);
    double2 x = make_double2(0.0, 0.0);
    if (!first) x = ld2(k.x + g);
    st2(k.x + g, x.x + (alpha * ph.x + omega * sh.x), x.y + (alpha * ph.y + omega * sh.y));
    const double r0 = s.x - omega * t.x, r1 = s.y - omega * t.y;
    st2(k.r + g, r0, r1);
    dot += rh.x * r0 + rh.y * r1;
    rabs = nmax(rabs, nmax(fabs(r0), fabs(r1)));
  }
  double sums[1] = {dot};
  block_sums<1>(sums, red);
  const double bmax = block_max(rabs, red);
  {
    __shared__ int last_flag;
    double* const dst[1] = {&sc->acc_rho};
    if (cta_totals<1>(sc, k.part + (size_t)blockIdx.y * k.part_stride, sums, dst, &sc->acc_rmax, bmax, &last_flag) && !k.defer) xr_finish_iteration(sc, k);
  }
}

}  // namespace sy2d
