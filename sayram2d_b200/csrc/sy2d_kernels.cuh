// Device kernels of the B200-native Sayram-2D time-step engine (sm_100a).
//
// Everything is IEEE fp64 and HBM/L2-bandwidth bound (no dense contraction => no
// tensor cores).  Layout: every field is [nbatch][nx][ny], j (log E) fastest, the
// reference's xtensor layout (source/common.h:28); a thread owns one cell and the
// flattened in-problem index n = i*ny + j is the coalescing axis.
//
// Linear system solved per time step (see DESIGN.md "formulation"):
//   reference:  M f^{n+1} = R                         (Solver.cc:270-278, direct LU)
//   here:       A d = rhs,  f^{n+1} = c (1 + d)
//     c   = f^n * yprev            column scale (yprev = predicted per-cell ratio, or 1)
//     A   = D_r M diag(c),  D_r = 1/(M_KK c_K)   => unit diagonal, 4 stored off-diagonals
//     rhs = D_r R - A 1            (residual of the guess "f^{n+1} = c")
//   so the unknown is O(1e-2) relative change, and max|r| <= tol bounds the
//   ELEMENTWISE relative error of f^{n+1}, which is what parity with the direct
//   solve over 17-49 orders of magnitude in f needs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sy2d {

constexpr int kBlock = 256;
constexpr double kEps = 2.220446049250313e-16;  // gEPS, source/common.h:38
// The predicted per-cell ratio f^{n+1}/f^n is only trusted inside this range; outside it (fast
// transients, e.g. loss-cone cells dropping 1000x per step) the unknown d would grow to O(100)
// and cost accuracy, so the prediction is clamped and d stays O(1).
constexpr double kPredMin = 0.5, kPredMax = 2.0;

// Column-scale predictor for the next step (the unknown of a step is d = f^{n+1} / (f^n yprev) - 1, so the closer yprev is to
// the coming ratio the closer the initial guess d = 0 is to the solution).  predictor 1: the last ratio y_n = f^{n+1} / f^n;
// predictor 2: its geometric extrapolation y_n (y_n / y_{n-1}) - quasi-steady decay and growth curve a little, and the
// extrapolated guess starts 10-20x closer (about one BiCGSTAB iteration per step).  Both clamped to [kPredMin, kPredMax].
// yl: the value of *ylast (0 until the first step after (re)setting f has been taken: nothing to extrapolate from)
__host__ __device__ __forceinline__ void predictor_update(int predictor, double fnew, double fold, double* yprev, double* ylast, double yl) {
  double y = fnew / fold;
  y = y < kPredMin ? kPredMin : (y > kPredMax ? kPredMax : y);
  if (!(y == y)) y = 1.0;
  double pred = y;
  if (predictor >= 2) {
    pred = yl > 0.0 ? y * (y / yl) : y;
    pred = pred < kPredMin ? kPredMin : (pred > kPredMax ? kPredMax : pred);
    if (!(pred == pred)) pred = y;
    *ylast = y;
  }
  *yprev = pred;
}
__host__ __device__ __forceinline__ void predictor_update(int predictor, double fnew, double fold, double* yprev, double* ylast) {
  predictor_update(predictor, fnew, fold, yprev, ylast, predictor >= 2 ? *ylast : 0.0);
}

// Per-problem Krylov scalars, written only by the last block of a kernel to
// finish that problem (threadfence reduction), read by every block of the NEXT
// kernels => no intra-kernel races and no host round trip.
struct Scal {
  double rho, alpha, omega, beta;
  double acc_rv, acc_ts, acc_tt, acc_rho;
  unsigned long long acc_rmax;  // max |r| as raw bits (non-negative doubles order like integers)
  double rmax;
  unsigned int counter;
  int it;
  int state;  // 0 active, 1 converged, 2 maxit, 3 breakdown
  int first;  // first iteration after (re)start: p = r, x = 0, r = rhs
};

struct Geometry {  // 1-D device arrays shared by the batch
  const double* wxL;  // [nx+1] weight of cell i-1 at vertex i   (Solver.cc:340; 0 at i=0, 1 at i=nx)
  const double* wxR;  // [nx+1] weight of cell i   at vertex i   (Solver.cc:341; 1 at i=0, 0 at i=nx)
  const double* wyB;  // [ny+1]
  const double* wyT;  // [ny+1]
  const double* bc_xmin;  // [ny+1] Dirichlet vertex lines (Solver.cc:385-422)
  const double* bc_xmax;
  const double* bc_ymin;  // [nx+1]
  const double* bc_ymax;
  int bc[4];
  int nx, ny;
};

// a / b for the assembly (mu, B/(f+eps), row scaling, LU pivots).  On the device: reciprocal seed
// (MUFU.RCP64H, ~20 bits), two Newton steps, one residual correction of the quotient - 8 fp64
// instructions without the special-case branches of the IEEE division sequence (denominators here
// are normal, positive numbers); the result is within 1 ulp of a / b.  On the host (tests/emul): a / b.
__host__ __device__ __forceinline__ double sy2d_div(double a, double b) {
#ifdef __CUDA_ARCH__
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  r = fma(r, fma(-b, r, 1.0), r);
  r = fma(r, fma(-b, r, 1.0), r);
  const double q = a * r;
  return fma(r, fma(-b, q, a), q);
#else
  return a / b;
#endif
}

// max that PROPAGATES NaN (fmax drops it): a residual reduction must not turn a NaN iterate into "max|r| = 0 => converged"
__host__ __device__ __forceinline__ double nmax(double a, double b) { return (b != b || b > a) ? b : a; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = nmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block reductions for up to 3 sums and 1 max; result valid in thread 0.
template <int NS>
__device__ __forceinline__ void block_sums(double (&v)[NS], double* smem /* >= NS*32 */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < NS; ++k) v[k] = warp_sum(v[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NS; ++k) smem[k * 32 + w] = v[k];
  }
  __syncthreads();
  if (w == 0) {
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      double x = lane < nw ? smem[k * 32 + lane] : 0.0;
      v[k] = warp_sum(x);
    }
  }
  __syncthreads();
}
__device__ __forceinline__ double block_max(double v, double* smem /* >= 32 */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  if (lane == 0) smem[w] = v;
  __syncthreads();
  if (w == 0) v = warp_max(lane < nw ? smem[lane] : 0.0);
  __syncthreads();
  return v;
}

// Cross-CTA sums of the lockstep engine.  `part` == NULL (default): one atomicAdd(double) per CTA and sum, the last CTA to arrive
// (atomic ticket) finishes - the sums are formed in arrival order, so two runs agree to round-off and their iteration counts can
// differ by one or two at tol = 1e-14.  `part` != NULL (SY2D_DETERMINISTIC=1 at context creation): WITHOUT floating-point atomics,
// i.e. deterministic: CTA b of a problem leaves its partial sums (and its partial
// maximum) in slot b of the problem's slot array `part`; the last CTA to arrive (atomic ticket) adds the slots IN SLOT ORDER with
// its first warp (lane l takes slots l, l + 32, ...; then the fixed shuffle tree) and adds the totals to *dst / maxes them into
// *dst_max, where the finishing code of the kernels - and, in row-slab mode, the all-gather over the ranks - reads them.  The
// result does not depend on the order in which the CTAs finished: two runs of the same problem give the same bits, iteration
// counts included.  Called by EVERY thread of the CTA (one barrier); true in thread 0 of the last CTA only.
constexpr int kPartSlot = 4;   // doubles per slot: up to three partial sums, one partial maximum
__device__ __forceinline__ bool last_block_done(Scal* sc, unsigned int nblocks);
template <int NS>
__device__ __forceinline__ bool cta_totals(Scal* sc, double* part, const double (&sums)[NS], double* const (&dst)[NS],
                                           unsigned long long* dst_max, double bmax, int* flag) {
  static_assert(NS <= 3, "three sums per slot");
  if (part == nullptr) {   // default: floating-point atomics, the last CTA to arrive finishes (5 % faster per 1024^2 time step, sums in arrival order)
    if (threadIdx.x != 0) return false;
#pragma unroll
    for (int k = 0; k < NS; ++k) atomicAdd(dst[k], sums[k]);
    if (dst_max) atomicMax(dst_max, (unsigned long long)__double_as_longlong(bmax));
    return last_block_done(sc, gridDim.x);
  }
  if (threadIdx.x == 0) {
    double* mine = part + (size_t)blockIdx.x * kPartSlot;
#pragma unroll
    for (int k = 0; k < NS; ++k) mine[k] = sums[k];
    if (dst_max) mine[3] = bmax;
    __threadfence();
    *flag = atomicAdd(&sc->counter, 1u) == gridDim.x - 1 ? 1 : 0;
  }
  __syncthreads();
  if (*flag == 0 || threadIdx.x >= 32) return false;
  __threadfence();
  double acc[NS], m = 0.0;
#pragma unroll
  for (int k = 0; k < NS; ++k) acc[k] = 0.0;
  // eight slots per lane in flight at a time (one L2 round trip per batch instead of one per slot); the order of the additions is
  // fixed by the code, not by the arrival of the loads
  for (unsigned b0 = threadIdx.x; b0 < gridDim.x; b0 += 32 * 8) {
    double v[8][NS + 1];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const unsigned b = b0 + 32 * u;
      const bool in = b < gridDim.x;
      const double* s = part + (size_t)(in ? b : 0) * kPartSlot;
#pragma unroll
      for (int k = 0; k < NS; ++k) v[u][k] = in ? __ldcg(s + k) : 0.0;
      v[u][NS] = (dst_max && in) ? __ldcg(s + 3) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int k = 0; k < NS; ++k) acc[k] += v[u][k];
      if (dst_max) m = nmax(m, v[u][NS]);
    }
  }
#pragma unroll
  for (int k = 0; k < NS; ++k) acc[k] = warp_sum(acc[k]);
  if (dst_max) m = warp_max(m);
  if (threadIdx.x != 0) return false;
#pragma unroll
  for (int k = 0; k < NS; ++k) *dst[k] += acc[k];
  if (dst_max) {   // raw bits: non-negative doubles order like integers, NaN wins
    const unsigned long long bits = (unsigned long long)__double_as_longlong(m);
    if (bits > *dst_max) *dst_max = bits;
  }
  sc->counter = 0;
  __threadfence();
  return true;
}

// true in exactly one thread (thread 0 of the last block of this problem to get here)
__device__ __forceinline__ bool last_block_done(Scal* sc, unsigned int nblocks) {
  __threadfence();
  const unsigned int prev = atomicAdd(&sc->counter, 1u);
  if (prev == nblocks - 1) {
    sc->counter = 0;
    __threadfence();
    return true;
  }
  return false;
}

__global__ void __launch_bounds__(kBlock) k_fill(double* __restrict__ a, size_t N, double value) {
  const size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n < N) a[(size_t)blockIdx.y * N + n] = value;
}

// ---------------------------------------------------------------------------
// Coefficient staging: Lambda = G*D (Solver.cc:57-65) folded with the face
// geometry into the closed forms of SURVEY.md section 0:
//   tx = Lxx*dy_j/dx_i, ty = Lyy*dx_i/dy_j, cxy = Lxy, U = G*dx*dy/dt (Solver.cc:193),
//   Ud = U*(1 + dt/tau) (Solver.cc:195).
// ---------------------------------------------------------------------------
struct CellCoeffs { double tx, ty, cxy, U, Ud; };
__host__ __device__ __forceinline__ CellCoeffs prepare_cell(double G, double Dxx, double Dxy, double Dyy, double inv_tau,
                                                            double dxi, double dyj, double dt) {
  CellCoeffs c;
  const double lxx = Dxx * G, lxy = Dxy * G, lyy = Dyy * G;  // Solver.cc:61-62
  c.tx = lxx * dyj / dxi;
  c.ty = lyy * dxi / dyj;
  c.cxy = lxy;
  c.U = G * (dxi * dyj / dt);             // Solver.cc:193, Mesh.h:61-63
  c.Ud = c.U * (1.0 + dt * inv_tau);      // Solver.cc:195
  return c;
}

__global__ void __launch_bounds__(kBlock) k_prepare_coeffs(
    const double* __restrict__ G, const double* __restrict__ Dxx, const double* __restrict__ Dxy,
    const double* __restrict__ Dyy, const double* __restrict__ inv_tau, const double* __restrict__ dx,
    const double* __restrict__ dy, double dt, int nx, int ny, double* __restrict__ tx, double* __restrict__ ty,
    double* __restrict__ cxy, double* __restrict__ U, double* __restrict__ Ud) {
  const size_t N = (size_t)nx * ny;
  const size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const size_t g = (size_t)blockIdx.y * N + n;
  const int i = (int)(n / ny), j = (int)(n - (size_t)i * ny);
  const CellCoeffs c = prepare_cell(G[g], Dxx[g], Dxy[g], Dyy[g], inv_tau ? inv_tau[g] : 0.0, dx[i], dy[j], dt);
  tx[g] = c.tx; ty[g] = c.ty; cxy[g] = c.cxy; U[g] = c.U; Ud[g] = c.Ud;
}

// vertex_f_(vi, vj) of Solver.cc:292-422 from the 2x2 cells around the vertex.
// The padded weights make the interior bilinear formula (Solver.cc:346) reproduce
// the edge lerp (:361-380) and the corner copy (:316-319); Dirichlet lines then
// override in the reference's order XMIN, XMAX, YMIN, YMAX (later wins).
__host__ __device__ __forceinline__ double vertex_value(const Geometry& g, int vi, int vj, double f00, double f10,
                                                        double f01, double f11) {
  if (vj == g.ny && g.bc[3] == 0) return g.bc_ymax[vi];
  if (vj == 0 && g.bc[2] == 0) return g.bc_ymin[vi];
  if (vi == g.nx && g.bc[1] == 0) return g.bc_xmax[vj];
  if (vi == 0 && g.bc[0] == 0) return g.bc_xmin[vj];
  const double wl = g.wxL[vi], wr = g.wxR[vi], wb = g.wyB[vj], wt = g.wyT[vj];
  return wl * wb * f00 + wr * wb * f10 + wl * wt * f01 + wr * wt * f11;
}

// Nonlinear two-point flux of one interior face (Solver.cc:99-141, Solver.h:63-67).
// K is the cell that owns the face as its west/south face (the reference's loop
// cell), L its neighbour.  Returns A_K, A_L.
__host__ __device__ __forceinline__ void face_pair(double asK, double sumK, double fK, double asL, double sumL,
                                                   double fL, double& AK, double& AL) {
  const double aK = fabs(asK), aL = fabs(asL);
  const double denom = aK + aL + 2.0 * kEps;
  const double muK = sy2d_div(aL + kEps, denom);
  const double muL = 1.0 - muK;
  const double B = muL * asL - muK * asK;
  const double Babs = fabs(B);
  const double Bp = (Babs + B) / 2.0, Bm = (Babs - B) / 2.0;
  // exactly one of Bp, Bm is non-zero and 0/(f+eps) == 0, so one division serves both
  const double q = sy2d_div(B > 0.0 ? Bp : Bm, (B > 0.0 ? fK : fL) + kEps);
  AK = muK * sumK + (B > 0.0 ? q : 0.0);
  AL = muL * sumL + (B > 0.0 ? 0.0 : q);
}

// Dirichlet boundary face (Solver.cc:143-164): returns A_K, adds B^- to R.
__host__ __device__ __forceinline__ double dirichlet_face(double asK, double sumK, double fK, double& R) {
  const double B = -asK;
  const double Babs = fabs(B);
  const double Bp = (Babs + B) / 2.0, Bm = (Babs - B) / 2.0;
  R += Bm;
  return sumK + sy2d_div(Bp, fK + kEps);
}

struct Row {  // one row of the reference's M and R (unscaled)
  double diag, oW, oE, oS, oN, R;
  double f00, fW, fE, fS, fN;     // f of the cell and its face neighbours (clamped)
  double vSW, vSE, vNW, vNE;      // vertex_f_ at the cell's four corners
};

// Row (i,j) of M(f), R(f): vertex values from the 3x3 block of f (never
// materialised in HBM), one-sided weights of the four faces and of the matching
// faces of the neighbours, nonlinear combination.  fp/tx/ty/cxy/U/Ud point at the
// start of the problem.  Replaces Solver::update_vertex_f, a_sigma_func,
// apply_inner_face_pair, apply_dirichlet_face, apply_boundary_faces and assemble
// (Solver.cc:68-267, 292-422).  __host__ too so that tests/ can run the same
// arithmetic on the CPU against the reference's (M,R) without a GPU.
__host__ __device__ __forceinline__ void assemble_row(const double* __restrict__ fp, const double* __restrict__ tx,
                                                      const double* __restrict__ ty, const double* __restrict__ cxy,
                                                      const double* __restrict__ U, const double* __restrict__ Ud,
                                                      const Geometry& g, int i, int j, Row& o) {
  const int nx = g.nx, ny = g.ny;
  const int im = i > 0 ? i - 1 : 0, ip = i < nx - 1 ? i + 1 : nx - 1;
  const int jm = j > 0 ? j - 1 : 0, jp = j < ny - 1 ? j + 1 : ny - 1;
  const size_t rm = (size_t)im * ny, r0 = (size_t)i * ny, rp = (size_t)ip * ny;
  // 3x3 block of f, clamped at the domain boundary (clamped values only meet zero weights)
  const double fmm = fp[rm + jm], fm0 = fp[rm + j], fmp = fp[rm + jp];
  const double f0m = fp[r0 + jm], f00 = fp[r0 + j], f0p = fp[r0 + jp];
  const double fpm = fp[rp + jm], fp0 = fp[rp + j], fpp = fp[rp + jp];
  // vertices of cell (i,j): SW=(i,j) SE=(i+1,j) NW=(i,j+1) NE=(i+1,j+1)   (Mesh.cc:94-149)
  const double vSW = vertex_value(g, i, j, fmm, f0m, fm0, f00);
  const double vSE = vertex_value(g, i + 1, j, f0m, fpm, f00, fp0);
  const double vNW = vertex_value(g, i, j + 1, fm0, f00, fmp, f0p);
  const double vNE = vertex_value(g, i + 1, j + 1, f00, fp0, f0p, fpp);
  o.vSW = vSW; o.vSE = vSE; o.vNW = vNW; o.vNE = vNE;
  o.f00 = f00; o.fW = fm0; o.fE = fp0; o.fS = f0m; o.fN = f0p;
  const size_t c0 = r0 + j;
  const double txP = tx[c0], tyP = ty[c0], cP = cxy[c0];
  // one-sided weights of cell P (closed form of Solver.cc:68-97)
  const double aW_A = txP - cP, aW_B = txP + cP;  // W: A=NW, B=SW
  const double aE_A = txP - cP, aE_B = txP + cP;  // E: A=SE, B=NE
  const double aN_A = tyP + cP, aN_B = tyP - cP;  // N: A=NE, B=NW
  const double aS_A = tyP + cP, aS_B = tyP - cP;  // S: A=SW, B=SE
  const double asW = aW_A * vNW + aW_B * vSW;
  const double asE = aE_A * vSE + aE_B * vNE;
  const double asN = aN_A * vNE + aN_B * vNW;
  const double asS = aS_A * vSW + aS_B * vSE;
  double diag = 0.0, R = 0.0, oW = 0.0, oE = 0.0, oS = 0.0, oN = 0.0;
  double AK, AL;
  if (i > 0) {  // west face: K = P, L = (i-1,j) through its east face (A=SE_L=SW_P, B=NE_L=NW_P)
    const size_t cl = c0 - ny;
    const double t = tx[cl], c = cxy[cl];
    const double lA = t - c, lB = t + c;
    face_pair(asW, aW_A + aW_B, f00, lA * vSW + lB * vNW, lA + lB, fm0, AK, AL);
    diag += AK;
    oW = -AL;
  } else if (g.bc[0] == 0) {
    diag += dirichlet_face(asW, aW_A + aW_B, f00, R);
  }
  if (i < nx - 1) {  // east face: K = (i+1,j) through its west face (A=NW_K=NE_P, B=SW_K=SE_P), L = P
    const size_t ck = c0 + ny;
    const double t = tx[ck], c = cxy[ck];
    const double kA = t - c, kB = t + c;
    face_pair(kA * vNE + kB * vSE, kA + kB, fp0, asE, aE_A + aE_B, f00, AK, AL);
    diag += AL;
    oE = -AK;
  } else if (g.bc[1] == 0) {
    diag += dirichlet_face(asE, aE_A + aE_B, f00, R);
  }
  if (j > 0) {  // south face: K = P, L = (i,j-1) through its north face (A=NE_L=SE_P, B=NW_L=SW_P)
    const size_t cl = c0 - 1;
    const double t = ty[cl], c = cxy[cl];
    const double lA = t + c, lB = t - c;
    face_pair(asS, aS_A + aS_B, f00, lA * vSE + lB * vSW, lA + lB, f0m, AK, AL);
    diag += AK;
    oS = -AL;
  } else if (g.bc[2] == 0) {
    diag += dirichlet_face(asS, aS_A + aS_B, f00, R);
  }
  if (j < ny - 1) {  // north face: K = (i,j+1) through its south face (A=SW_K=NW_P, B=SE_K=NE_P), L = P
    const size_t ck = c0 + 1;
    const double t = ty[ck], c = cxy[ck];
    const double kA = t + c, kB = t - c;
    face_pair(kA * vNW + kB * vNE, kA + kB, f0p, asN, aN_A + aN_B, f00, AK, AL);
    diag += AL;
    oN = -AK;
  } else if (g.bc[3] == 0) {
    diag += dirichlet_face(asN, aN_A + aN_B, f00, R);
  }
  diag += Ud[c0];    // Solver.cc:195
  R += U[c0] * f00;  // Solver.cc:197
  o.diag = diag; o.oW = oW; o.oE = oE; o.oS = oS; o.oN = oN; o.R = R;
}

struct Scaled {  // row of the scaled unit-diagonal system A d = rhs
  double wW, wE, wS, wN, rhs, cs;
};

// Column scale c = f*yprev (neighbours too), row scale 1/(M_KK c_K); rhs = D_r R - A 1.
__host__ __device__ __forceinline__ void scale_row(const Row& r, double ypC, double ypW, double ypE, double ypS,
                                                   double ypN, Scaled& s) {
  const double cs0 = r.f00 * ypC;
  const double dscale = sy2d_div(1.0, r.diag * cs0);
  s.wW = r.oW * (r.fW * ypW) * dscale;
  s.wE = r.oE * (r.fE * ypE) * dscale;
  s.wS = r.oS * (r.fS * ypS) * dscale;
  s.wN = r.oN * (r.fN * ypN) * dscale;
  s.rhs = r.R * dscale - 1.0 - ((s.wW + s.wE) + (s.wS + s.wN));
  s.cs = cs0;
}

// LU of the x-line tridiagonal T = tridiag(wW, 1, wE) along i (one column j):
//   d_0 = 1,  l_i = wW_i / d_{i-1},  d_i = 1 - l_i * wE_{i-1}
// T is an M-matrix (A with its S/N couplings dropped), so the pivots stay positive.
struct XlineFactor { double l, d, dinv; };
__host__ __device__ __forceinline__ XlineFactor xline_factor(double wW_i, double wE_prev, double d_prev, bool first_row) {
  XlineFactor f;
  f.l = first_row ? 0.0 : sy2d_div(wW_i, d_prev);
  f.d = 1.0 - f.l * wE_prev;
  f.dinv = sy2d_div(1.0, f.d);
  return f;
}

struct AssembleOut {
  // MODE 0 (solve): scaled unit-diagonal operator + Krylov start
  double *wW, *wE, *wS, *wN, *rhs, *cs;
  double* om;   // optional: row weight M_KK c_K = 1 / (row scale), the multigrid restriction weight (may be NULL)
  Scal* scal;
  double* part;         // slots of the deterministic cross-CTA sums (cta_totals): [nbatch][part_stride]
  size_t part_stride;
  int* n_active;
  double tol;
  // MODE 1 (dump): the reference's unscaled M as 5 diagonals + R
  double *diag, *oW, *oE, *oS, *oN, *R;
  // MODE 2 (dump vertex_f): [nbatch][nx+1][ny+1]
  double* vf;
  int local_rows;  // rows of the (local) arrays: nx, or nx_loc + 2 for a slab with halo rows
};

// ---------------------------------------------------------------------------
// Fused PPFV assembly, one thread per cell (see assemble_row / scale_row).
// Algorithmic HBM bytes per cell (MODE 0): read f, yprev, tx, ty, cxy, U, Ud (56) +
// write wW, wE, wS, wN, rhs, cs (48) = 104 B (96 B with predictor off).
// ---------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kBlock) k_assemble(const double* __restrict__ f, const double* __restrict__ yprev,
                                                     const double* __restrict__ tx, const double* __restrict__ ty,
                                                     const double* __restrict__ cxy, const double* __restrict__ U,
                                                     const double* __restrict__ Ud, Geometry g, AssembleOut o) {
  __shared__ double red[3 * 32];
  const int nx = g.nx, ny = g.ny;
  const size_t N = (size_t)nx * ny;
  const size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t base = (size_t)blockIdx.y * N;
  double rr = 0.0, rabs = 0.0;
  if (n < N) {
    const int i = (int)(n / ny), j = (int)(n - (size_t)i * ny);
    Row row;
    assemble_row(f + base, tx + base, ty + base, cxy + base, U + base, Ud + base, g, i, j, row);
    const size_t c0 = base + n;
    if (MODE == 2) {
      double* vf = o.vf + (size_t)blockIdx.y * (nx + 1) * (ny + 1);
      vf[(size_t)i * (ny + 1) + j] = row.vSW;
      if (i == nx - 1) vf[(size_t)nx * (ny + 1) + j] = row.vSE;
      if (j == ny - 1) vf[(size_t)i * (ny + 1) + ny] = row.vNW;
      if (i == nx - 1 && j == ny - 1) vf[(size_t)nx * (ny + 1) + ny] = row.vNE;
    } else if (MODE == 1) {
      o.diag[c0] = row.diag; o.oW[c0] = row.oW; o.oE[c0] = row.oE; o.oS[c0] = row.oS; o.oN[c0] = row.oN; o.R[c0] = row.R;
    } else {
      const double* yp = yprev + base;
      const size_t nW = i > 0 ? n - ny : n, nE = i < nx - 1 ? n + ny : n;
      const size_t nS = j > 0 ? n - 1 : n, nN = j < ny - 1 ? n + 1 : n;
      Scaled sc;
      scale_row(row, yp[n], yp[nW], yp[nE], yp[nS], yp[nN], sc);
      o.wW[c0] = sc.wW; o.wE[c0] = sc.wE; o.wS[c0] = sc.wS; o.wN[c0] = sc.wN;
      o.rhs[c0] = sc.rhs;
      o.cs[c0] = sc.cs;
      if (o.om) o.om[c0] = row.diag * sc.cs;   // M_KK c_K: what scale_row divided the row by
      rr = sc.rhs * sc.rhs;
      rabs = fabs(sc.rhs);
    }
  }
  if (MODE != 0) return;
  // Krylov start: rho = (rhat, r0) = |rhs|^2, max-norm of the initial residual
  double sums[1] = {rr};
  block_sums<1>(sums, red);
  const double bmax = block_max(rabs, red);
  Scal* sc = o.scal + blockIdx.y;
  {
    __shared__ int last_flag;
    double* const dst[1] = {&sc->acc_rho};
    if (cta_totals<1>(sc, o.part + (size_t)blockIdx.y * o.part_stride, sums, dst, &sc->acc_rmax, bmax, &last_flag)) {
      const double rmax = __longlong_as_double((long long)sc->acc_rmax);
      sc->rho = sc->acc_rho;
      sc->rmax = rmax;
      sc->alpha = 1.0; sc->omega = 1.0; sc->beta = 0.0;
      sc->acc_rv = 0.0; sc->acc_ts = 0.0; sc->acc_tt = 0.0; sc->acc_rho = 0.0; sc->acc_rmax = 0ull;
      sc->it = 0;
      sc->first = 1;
      const int active_now = !(rmax <= o.tol);  // NaN counts as active => surfaces as an error later
      sc->state = active_now ? 0 : 1;
      if (active_now) atomicAdd(o.n_active, 1);
    }
  }
}

// ---------------------------------------------------------------------------
// Tiled PPFV assembly for large grids (engine 1): a CTA of 256 threads owns a tile of
// TI x TJ = 8 x 32 cells.  Shared-memory stages, each computed ONCE per tile:
//   1. f and c = f*yprev with a one-cell halo                       (10 x 34)
//   2. the vertex values of the tile                                 (9 x 33)
//   3. the nonlinear two-point fluxes of the west faces (9 x 32) and south faces (8 x 33):
//      A_K, A_L per face (face_pair), so every interior face is evaluated once per tile
//      instead of once per adjacent cell
//   4. the operator row of every cell, gathered from its four faces, scaled, written coalesced.
// Per cell: ~1.1 vertex, ~2.2 face evaluations and 5 divisions instead of 4, 4 and 9 in the
// one-thread-per-cell kernel, 29 KB of shared memory per CTA; every global load happens in stage 1
// (coalesced rows of the halo tile), the later stages read shared memory only.  Same arithmetic per face, so the
// rows are bit-identical to k_assemble<0>.
// ---------------------------------------------------------------------------
constexpr int kTI = 8, kTJ = 32;

__global__ void __launch_bounds__(kTI * kTJ, 4) k_assemble_tiled(const double* __restrict__ f, const double* __restrict__ yprev,
                                                              const double* __restrict__ tx, const double* __restrict__ ty,
                                                              const double* __restrict__ cxy, const double* __restrict__ U,
                                                              const double* __restrict__ Ud, Geometry g, AssembleOut o, int tiles_j,
                                                              int gi0, int li_begin, int li_end, int defer) {
  // halo tiles (one cell around the 8 x 32 tile): f, c = f*yprev and the face coefficients
  __shared__ double fs[kTI + 2][kTJ + 2], cs_[kTI + 2][kTJ + 2], txs[kTI + 2][kTJ + 2], tys[kTI + 2][kTJ + 2], cxs[kTI + 2][kTJ + 2];
  __shared__ double vs[kTI + 1][kTJ + 1];
  __shared__ double WK[kTI + 1][kTJ], WL[kTI + 1][kTJ], SK[kTI][kTJ + 1], SL[kTI][kTJ + 1];
  __shared__ double red[3 * 32];
  // Rows: memory is indexed with LOCAL rows li (the array may be a slab with halo rows), geometry
  // and boundary logic with GLOBAL rows i = gi0 + li; the kernel assembles local rows [li_begin, li_end).
  // Single-GPU: gi0 = 0, li_begin = 0, li_end = nx.  All in-problem indices fit 32 bits.
  const int nx = g.nx, ny = g.ny;
  const size_t base = (size_t)blockIdx.y * ((size_t)o.local_rows * ny);
  const int tid = threadIdx.x;
  const int ta = tid >> 5, tb = tid & 31;   // thread's cell inside the tile
  const int ntiles = tiles_j * ((li_end - li_begin + kTI - 1) / kTI);
  double rr = 0.0, rabs = 0.0;
  const double* fp = f + base;
  const double* yp = yprev + base;
  const double* txp = tx + base;
  const double* typ = ty + base;
  const double* cp = cxy + base;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {  // a CTA walks over several tiles: few atomics per problem
    const int tile_i = tile / tiles_j, tile_j = tile - tile_i * tiles_j;
    const int I0 = gi0 + li_begin + tile_i * kTI, J0 = tile_j * kTJ;   // global row / column of the tile origin
    // 1. halo tiles; indices clamped at the domain boundary (clamped values only meet zero weights)
    auto stage = [&](int a, int b) {
      int i = I0 + a - 1, j = J0 + b - 1;
      i = i < 0 ? 0 : (i > nx - 1 ? nx - 1 : i);
      j = j < 0 ? 0 : (j > ny - 1 ? ny - 1 : j);
      const int n = (i - gi0) * ny + j;
      const double fv = fp[n];
      fs[a][b] = fv;
      cs_[a][b] = fv * yp[n];
      txs[a][b] = txp[n];
      tys[a][b] = typ[n];
      cxs[a][b] = cp[n];
    };
    stage(ta, tb);
    if (ta < 2) stage(ta + kTI, tb);
    if (tb < 2) stage(ta, tb + kTJ);
    if (ta < 2 && tb < 2) stage(ta + kTI, tb + kTJ);
    __syncthreads();
    // 2. vertices (I0 + a, J0 + b); tiles that do not touch the boundary skip the Dirichlet / edge logic
    const bool edge_tile = I0 == 0 || I0 + kTI >= nx || J0 == 0 || J0 + kTJ >= ny;
    auto vertex = [&](int a, int b) {
      const int vi = I0 + a, vj = J0 + b;
      double v = 0.0;
      if (!edge_tile) {
        const double wl = g.wxL[vi], wr = g.wxR[vi], wb = g.wyB[vj], wt = g.wyT[vj];
        v = wl * wb * fs[a][b] + wr * wb * fs[a + 1][b] + wl * wt * fs[a][b + 1] + wr * wt * fs[a + 1][b + 1];
      } else if (vi <= nx && vj <= ny) {
        v = vertex_value(g, vi, vj, fs[a][b], fs[a + 1][b], fs[a][b + 1], fs[a + 1][b + 1]);
      }
      vs[a][b] = v;
    };
    vertex(ta, tb);
    if (ta == 0) vertex(kTI, tb);
    if (tid <= kTI) vertex(tid, kTJ);
    __syncthreads();
    // 3a. west faces of cells (I0 + a, J0 + b), a = 0..TI: K = (i, j), L = (i-1, j)
    auto wface = [&](int a, int b) {
      const int i = I0 + a, j = J0 + b;
      double AK = 0.0, AL = 0.0;
      if (i >= 1 && i <= nx - 1 && j < ny) {
        const double tK = txs[a + 1][b + 1], cK = cxs[a + 1][b + 1], tL = txs[a][b + 1], cL = cxs[a][b + 1];
        const double vSW = vs[a][b], vNW = vs[a][b + 1];
        const double kA = tK - cK, kB = tK + cK;   // W face of K: A = NW, B = SW
        const double lA = tL - cL, lB = tL + cL;   // E face of L: A = SE_L = SW_K, B = NE_L = NW_K
        face_pair(kA * vNW + kB * vSW, kA + kB, fs[a + 1][b + 1], lA * vSW + lB * vNW, lA + lB, fs[a][b + 1], AK, AL);
      }
      WK[a][b] = AK;
      WL[a][b] = AL;
    };
    wface(ta, tb);
    if (ta == 0) wface(kTI, tb);
    // 3b. south faces of cells (I0 + a, J0 + b), b = 0..TJ: K = (i, j), L = (i, j-1)
    auto sface = [&](int a, int b) {
      const int i = I0 + a, j = J0 + b;
      double AK = 0.0, AL = 0.0;
      if (j >= 1 && j <= ny - 1 && i < nx) {
        const double tK = tys[a + 1][b + 1], cK = cxs[a + 1][b + 1], tL = tys[a + 1][b], cL = cxs[a + 1][b];
        const double vSW = vs[a][b], vSE = vs[a + 1][b];
        const double kA = tK + cK, kB = tK - cK;   // S face of K: A = SW, B = SE
        const double lA = tL + cL, lB = tL - cL;   // N face of L: A = NE_L = SE_K, B = NW_L = SW_K
        face_pair(kA * vSW + kB * vSE, kA + kB, fs[a + 1][b + 1], lA * vSE + lB * vSW, lA + lB, fs[a + 1][b], AK, AL);
      }
      SK[a][b] = AK;
      SL[a][b] = AL;
    };
    sface(ta, tb);
    if (tid < kTI) sface(tid, kTJ);
    __syncthreads();
    // 4. rows
    const int a = ta, b = tb;
    const int i = I0 + a, j = J0 + b;
    if (i < nx && i - gi0 < li_end && j < ny) {
      const int n = (i - gi0) * ny + j;
      const size_t c0 = base + n;
      const double f00 = fs[a + 1][b + 1];
      double diag = 0.0, R = 0.0, oW = 0.0, oE = 0.0, oS = 0.0, oN = 0.0;
      if (i > 0) { diag += WK[a][b]; oW = -WL[a][b]; }
      if (i < nx - 1) { diag += WL[a + 1][b]; oE = -WK[a + 1][b]; }
      if (j > 0) { diag += SK[a][b]; oS = -SL[a][b]; }
      if (j < ny - 1) { diag += SL[a][b + 1]; oN = -SK[a][b + 1]; }
      if (edge_tile && (i == 0 || i == nx - 1 || j == 0 || j == ny - 1)) {  // Dirichlet boundary faces (Solver.cc:143-164, 204-267)
        const double txP = txs[a + 1][b + 1], tyP = tys[a + 1][b + 1], cP = cxs[a + 1][b + 1];
        const double vSW = vs[a][b], vSE = vs[a + 1][b], vNW = vs[a][b + 1], vNE = vs[a + 1][b + 1];
        if (i == 0 && g.bc[0] == 0) diag += dirichlet_face((txP - cP) * vNW + (txP + cP) * vSW, (txP - cP) + (txP + cP), f00, R);
        if (i == nx - 1 && g.bc[1] == 0) diag += dirichlet_face((txP - cP) * vSE + (txP + cP) * vNE, (txP - cP) + (txP + cP), f00, R);
        if (j == 0 && g.bc[2] == 0) diag += dirichlet_face((tyP + cP) * vSW + (tyP - cP) * vSE, (tyP + cP) + (tyP - cP), f00, R);
        if (j == ny - 1 && g.bc[3] == 0) diag += dirichlet_face((tyP + cP) * vNE + (tyP - cP) * vNW, (tyP + cP) + (tyP - cP), f00, R);
      }
      diag += Ud[c0];
      R += U[c0] * f00;
      const double cs0 = cs_[a + 1][b + 1];
      const double om = diag * cs0;
      const double dscale = sy2d_div(1.0, om);
      const double wW = oW * cs_[a][b + 1] * dscale, wE = oE * cs_[a + 2][b + 1] * dscale;
      const double wS = oS * cs_[a + 1][b] * dscale, wN = oN * cs_[a + 1][b + 2] * dscale;
      const double rhs = R * dscale - 1.0 - ((wW + wE) + (wS + wN));
      o.wW[c0] = wW; o.wE[c0] = wE; o.wS[c0] = wS; o.wN[c0] = wN;
      o.rhs[c0] = rhs;
      o.cs[c0] = cs0;
      if (o.om) o.om[c0] = om;
      rr += rhs * rhs;
      rabs = nmax(rabs, fabs(rhs));
    }
    __syncthreads();  // the tile's shared arrays are rewritten by the next tile
  }
  double sums[1] = {rr};
  block_sums<1>(sums, red);
  const double bmax = block_max(rabs, red);
  Scal* sc = o.scal + blockIdx.y;
  {
    __shared__ int last_flag;
    double* const dst[1] = {&sc->acc_rho};
    if (cta_totals<1>(sc, o.part + (size_t)blockIdx.y * o.part_stride, sums, dst, &sc->acc_rmax, bmax, &last_flag) && !defer) {
      const double rmax = __longlong_as_double((long long)sc->acc_rmax);
      sc->rho = sc->acc_rho;
      sc->rmax = rmax;
      sc->alpha = 1.0; sc->omega = 1.0; sc->beta = 0.0;
      sc->acc_rv = 0.0; sc->acc_ts = 0.0; sc->acc_tt = 0.0; sc->acc_rho = 0.0; sc->acc_rmax = 0ull;
      sc->it = 0;
      sc->first = 1;
      const int active_now = !(rmax <= o.tol);
      sc->state = active_now ? 0 : 1;
      if (active_now) atomicAdd(o.n_active, 1);
    }
  }
}

// y = A x for one cell of the unit-diagonal 5-point operator.  Out-of-range
// neighbours are clamped to the cell itself: their weights are exactly zero.
__host__ __device__ __forceinline__ double stencil_apply(const double* __restrict__ x, size_t n, size_t N, int ny, double xc,
                                                double wW, double wE, double wS, double wN) {
  const double xW = x[n >= (size_t)ny ? n - ny : n];
  const double xE = x[n + ny < N ? n + ny : n];
  const double xS = x[n > 0 ? n - 1 : n];
  const double xN = x[n + 1 < N ? n + 1 : n];
  return xc + ((wW * xW + wE * xE) + (wS * xS + wN * xN));
}

struct KrylovVecs {
  const double *wW, *wE, *wS, *wN, *rhs;
  double *x, *r, *p, *v, *s, *t;
  Scal* scal;
  double* part;         // slots of the deterministic cross-CTA sums (cta_totals): [nbatch][part_stride]
  size_t part_stride;   // doubles per problem
  int* n_active;   // [0] problems still iterating, [1] problems that stopped without converging (maxit / breakdown)
  double tol;
  int maxit;
  int freeze_state;  // sy2d_bench_kernel: keep every problem active whatever the residual does
  // row-slab mode (one grid split over ranks): the kernels update cells [n_begin, n_end) of a local
  // array that carries one halo row on each side, and leave the reduction accumulators untouched
  // (defer = 1) - they are all-gathered over NCCL and turned into scalars by k_slab_scalars.
  size_t n_begin, n_end;
  int defer;
};

// End-of-iteration bookkeeping done by the last block of a problem: iteration count,
// convergence / breakdown / maxit, beta for the next p update.
__device__ __forceinline__ void xr_finish_iteration(Scal* sc, const KrylovVecs& k) {
  const double rho_new = sc->acc_rho;
  const double rmax = __longlong_as_double((long long)sc->acc_rmax);
  sc->acc_rho = 0.0;
  sc->acc_rmax = 0ull;
  sc->rmax = rmax;
  sc->it += 1;
  sc->first = 0;
  int state = 0;
  if (rmax <= k.tol) state = 1;
  else if (!(rmax == rmax) || !(rho_new == rho_new) || rho_new == 0.0 || sc->omega == 0.0) state = 3;  // NaN / breakdown
  else if (sc->it >= k.maxit) state = 2;
  sc->beta = (rho_new / sc->rho) * (sc->alpha / sc->omega);
  sc->rho = rho_new;
  if (state != 0 && !k.freeze_state) {
    sc->state = state;
    atomicSub(k.n_active, 1);
    if (state >= 2) atomicAdd(k.n_active + 1, 1);   // n_active[1]: problems that stopped without converging
  }
}

// KA: p = r + beta (p - omega v)            (first iteration: p = rhs)      32 B/cell
__global__ void __launch_bounds__(kBlock) k_p_update(KrylovVecs k, size_t N) {
  const Scal* sc = k.scal + blockIdx.y;
  if (sc->state != 0) return;
  const size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const size_t g = (size_t)blockIdx.y * N + n;
  if (sc->first) {
    k.p[g] = k.rhs[g];
  } else {
    k.p[g] = k.r[g] + sc->beta * (k.p[g] - sc->omega * k.v[g]);
  }
}

// KB: v = A p, acc (rhat, v); last block: alpha = rho / (rhat, v)            56 B/cell
__global__ void __launch_bounds__(kBlock, 6) k_spmv_v(KrylovVecs k, size_t N, int ny) {
  __shared__ double red[32];
  Scal* sc = k.scal + blockIdx.y;
  if (sc->state != 0) return;
  const size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t base = (size_t)blockIdx.y * N;
  double dot = 0.0;
  if (n < N) {
    const size_t g = base + n;
    const double* p = k.p + base;
    const double v = stencil_apply(p, n, N, ny, p[n], k.wW[g], k.wE[g], k.wS[g], k.wN[g]);
    k.v[g] = v;
    dot = k.rhs[g] * v;
  }
  double sums[1] = {dot};
  block_sums<1>(sums, red);
  {
    __shared__ int last_flag;
    double* const dst[1] = {&sc->acc_rv};
    if (cta_totals<1>(sc, k.part + (size_t)blockIdx.y * k.part_stride, sums, dst, nullptr, 0.0, &last_flag)) {
      const double rv = sc->acc_rv;
      sc->acc_rv = 0.0;
      sc->alpha = rv != 0.0 ? sc->rho / rv : 0.0;
    }
  }
}

// KC: s = r - alpha v                                                         24 B/cell
__global__ void __launch_bounds__(kBlock) k_s_update(KrylovVecs k, size_t N) {
  const Scal* sc = k.scal + blockIdx.y;
  if (sc->state != 0) return;
  const size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const size_t g = (size_t)blockIdx.y * N + n;
  const double r = sc->first ? k.rhs[g] : k.r[g];
  k.s[g] = r - sc->alpha * k.v[g];
}

// KD: t = A s, acc (t,s), (t,t); last block: omega = (t,s)/(t,t)             48 B/cell
__global__ void __launch_bounds__(kBlock, 6) k_spmv_t(KrylovVecs k, size_t N, int ny) {
  __shared__ double red[2 * 32];
  Scal* sc = k.scal + blockIdx.y;
  if (sc->state != 0) return;
  const size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t base = (size_t)blockIdx.y * N;
  double ts = 0.0, tt = 0.0;
  if (n < N) {
    const size_t g = base + n;
    const double* s = k.s + base;
    const double sc0 = s[n];
    const double t = stencil_apply(s, n, N, ny, sc0, k.wW[g], k.wE[g], k.wS[g], k.wN[g]);
    k.t[g] = t;
    ts = t * sc0;
    tt = t * t;
  }
  double sums[2] = {ts, tt};
  block_sums<2>(sums, red);
  {
    __shared__ int last_flag;
    double* const dst[2] = {&sc->acc_ts, &sc->acc_tt};
    if (cta_totals<2>(sc, k.part + (size_t)blockIdx.y * k.part_stride, sums, dst, nullptr, 0.0, &last_flag)) {
      const double a = sc->acc_ts, b = sc->acc_tt;
      sc->acc_ts = 0.0;
      sc->acc_tt = 0.0;
      sc->omega = b > 0.0 ? a / b : 0.0;
    }
  }
}

// KE: x += alpha p + omega s; r = s - omega t; acc (rhat, r), max|r|;
// last block: iteration bookkeeping, convergence, beta                        56 B/cell
__global__ void __launch_bounds__(kBlock, 6) k_xr_update(KrylovVecs k, size_t N) {
  __shared__ double red[32];
  Scal* sc = k.scal + blockIdx.y;
  if (sc->state != 0) return;
  const size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  double dot = 0.0, rabs = 0.0;
  if (n < N) {
    const size_t g = (size_t)blockIdx.y * N + n;
    const double alpha = sc->alpha, omega = sc->omega;
    const double s = k.s[g];
    const double x0 = sc->first ? 0.0 : k.x[g];
    k.x[g] = x0 + (alpha * k.p[g] + omega * s);
    const double r = s - omega * k.t[g];
    k.r[g] = r;
    dot = k.rhs[g] * r;
    rabs = fabs(r);
  }
  double sums[1] = {dot};
  block_sums<1>(sums, red);
  const double bmax = block_max(rabs, red);
  {
    __shared__ int last_flag;
    double* const dst[1] = {&sc->acc_rho};
    if (cta_totals<1>(sc, k.part + (size_t)blockIdx.y * k.part_stride, sums, dst, &sc->acc_rmax, bmax, &last_flag)) xr_finish_iteration(sc, k);
  }
}

// ---------------------------------------------------------------------------
// Two-cells-per-thread variants of the five Krylov kernels (ny even): every array access is a
// 16-byte ld/st.global.v2.f64, the two cells share their inner S/N neighbours in registers, and
// the grid is half as large (half the block reductions and atomics).  Same arithmetic per cell.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }

__global__ void __launch_bounds__(kBlock) k_p_update2(KrylovVecs k, size_t N) {
  const Scal* sc = k.scal + blockIdx.y;
  if (sc->state != 0) return;
  const size_t stride = 2 * (size_t)gridDim.x * blockDim.x;
  const bool first = sc->first;
  const double beta = sc->beta, omega = sc->omega;
  for (size_t n = k.n_begin + 2 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x); n < k.n_end; n += stride) {
    const size_t g = (size_t)blockIdx.y * N + n;
    if (first) {
      const double2 r = ld2(k.rhs + g);
      st2(k.p + g, r.x, r.y);
    } else {
      const double2 r = ld2(k.r + g), p = ld2(k.p + g), v = ld2(k.v + g);
      st2(k.p + g, r.x + beta * (p.x - omega * v.x), r.y + beta * (p.y - omega * v.y));
    }
  }
}

// y = A x for cells n, n+1 (n even, ny even => both in the same row)
__device__ __forceinline__ void stencil_apply2(const double* __restrict__ x, size_t n, size_t N, int ny, double2 xc, double2 wW,
                                               double2 wE, double2 wS, double2 wN, double& y0, double& y1) {
  const double2 xW = ld2(x + (n >= (size_t)ny ? n - ny : n));
  const double2 xE = ld2(x + (n + ny < N ? n + ny : n));
  const double xS = x[n > 0 ? n - 1 : n];
  const double xN = x[n + 2 < N ? n + 2 : n + 1];
  y0 = xc.x + ((wW.x * xW.x + wE.x * xE.x) + (wS.x * xS + wN.x * xc.y));
  y1 = xc.y + ((wW.y * xW.y + wE.y * xE.y) + (wS.y * xc.x + wN.y * xN));
}

__global__ void __launch_bounds__(kBlock, 6) k_spmv_v2(KrylovVecs k, size_t N, int ny) {
  __shared__ double red[32];
  Scal* sc = k.scal + blockIdx.y;
  if (sc->state != 0) return;
  const size_t stride = 2 * (size_t)gridDim.x * blockDim.x;
  const size_t base = (size_t)blockIdx.y * N;
  double dot = 0.0;
  for (size_t n = k.n_begin + 2 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x); n < k.n_end; n += stride) {
    const size_t g = base + n;
    const double* p = k.p + base;
    double v0, v1;
    stencil_apply2(p, n, N, ny, ld2(p + n), ld2(k.wW + g), ld2(k.wE + g), ld2(k.wS + g), ld2(k.wN + g), v0, v1);
    st2(k.v + g, v0, v1);
    const double2 rh = ld2(k.rhs + g);
    dot += rh.x * v0 + rh.y * v1;
  }
  double sums[1] = {dot};
  block_sums<1>(sums, red);
  {
    __shared__ int last_flag;
    double* const dst[1] = {&sc->acc_rv};
    if (cta_totals<1>(sc, k.part + (size_t)blockIdx.y * k.part_stride, sums, dst, nullptr, 0.0, &last_flag) && !k.defer) {
      const double rv = sc->acc_rv;
      sc->acc_rv = 0.0;
      sc->alpha = rv != 0.0 ? sc->rho / rv : 0.0;
    }
  }
}

__global__ void __launch_bounds__(kBlock) k_s_update2(KrylovVecs k, size_t N) {
  const Scal* sc = k.scal + blockIdx.y;
  if (sc->state != 0) return;
  const size_t stride = 2 * (size_t)gridDim.x * blockDim.x;
  const double alpha = sc->alpha;
  const double* rsrc = sc->first ? k.rhs : k.r;
  for (size_t n = k.n_begin + 2 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x); n < k.n_end; n += stride) {
    const size_t g = (size_t)blockIdx.y * N + n;
    const double2 r = ld2(rsrc + g), v = ld2(k.v + g);
    st2(k.s + g, r.x - alpha * v.x, r.y - alpha * v.y);
  }
}

__global__ void __launch_bounds__(kBlock, 6) k_spmv_t2(KrylovVecs k, size_t N, int ny) {
  __shared__ double red[2 * 32];
  Scal* sc = k.scal + blockIdx.y;
  if (sc->state != 0) return;
  const size_t stride = 2 * (size_t)gridDim.x * blockDim.x;
  const size_t base = (size_t)blockIdx.y * N;
  double ts = 0.0, tt = 0.0;
  for (size_t n = k.n_begin + 2 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x); n < k.n_end; n += stride) {
    const size_t g = base + n;
    const double* s = k.s + base;
    const double2 sc0 = ld2(s + n);
    double t0, t1;
    stencil_apply2(s, n, N, ny, sc0, ld2(k.wW + g), ld2(k.wE + g), ld2(k.wS + g), ld2(k.wN + g), t0, t1);
    st2(k.t + g, t0, t1);
    ts += t0 * sc0.x + t1 * sc0.y;
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

__global__ void __launch_bounds__(kBlock, 6) k_xr_update2(KrylovVecs k, size_t N) {
  __shared__ double red[32];
  Scal* sc = k.scal + blockIdx.y;
  if (sc->state != 0) return;
  const size_t stride = 2 * (size_t)gridDim.x * blockDim.x;
  double dot = 0.0, rabs = 0.0;
  const double alpha = sc->alpha, omega = sc->omega;
  const bool first = sc->first;
  for (size_t n = k.n_begin + 2 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x); n < k.n_end; n += stride) {
    const size_t g = (size_t)blockIdx.y * N + n;
    const double2 s = ld2(k.s + g), p = ld2(k.p + g), t = ld2(k.t + g), rh = ld2(k.rhs + g);
    double2 x = make_double2(0.0, 0.0);
    if (!first) x = ld2(k.x + g);
    st2(k.x + g, x.x + (alpha * p.x + omega * s.x), x.y + (alpha * p.y + omega * s.y));
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

// Row-slab mode: every rank contributed its five accumulators {acc_rv, acc_ts, acc_tt, acc_rho,
// acc_rmax} through ncclAllGather into gathered[nranks][5]; one thread reduces them in rank order
// (identical on every rank, so all ranks take the same decisions) and does what the last block of
// the single-GPU kernels does.  phase: 0 after the assembly, 1 after v = A p, 2 after t = A s,
// 3 after the x, r update.
__global__ void k_slab_scalars(int phase, Scal* sc, const double* __restrict__ gathered, int nranks, KrylovVecs k) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (phase != 0 && sc->state != 0) return;   // stopped: the rest of the polling chunk must not count iterations or touch rmax
  double rv = 0.0, ts = 0.0, tt = 0.0, rho = 0.0, rmax = 0.0;
  for (int r = 0; r < nranks; ++r) {
    const double* q = gathered + 5 * r;
    rv += q[0]; ts += q[1]; tt += q[2]; rho += q[3];
    rmax = nmax(rmax, __longlong_as_double(__double_as_longlong(q[4])));  // raw bits of a non-negative double
  }
  sc->acc_rv = 0.0; sc->acc_ts = 0.0; sc->acc_tt = 0.0;
  if (phase == 0) {
    sc->rho = rho; sc->rmax = rmax;
    sc->alpha = 1.0; sc->omega = 1.0; sc->beta = 0.0;
    sc->acc_rho = 0.0; sc->acc_rmax = 0ull;
    sc->it = 0; sc->first = 1;
    const int active_now = !(rmax <= k.tol);
    sc->state = active_now ? 0 : 1;
    if (active_now) atomicAdd(k.n_active, 1);
  } else if (phase == 1) {
    sc->alpha = rv != 0.0 ? sc->rho / rv : 0.0;
  } else if (phase == 2) {
    sc->omega = tt > 0.0 ? ts / tt : 0.0;
  } else {
    sc->acc_rho = rho;
    sc->acc_rmax = (unsigned long long)__double_as_longlong(rmax);
    xr_finish_iteration(sc, k);
  }
}

// max over the ranks' all-gathered values (non-negative doubles; NaN propagates)
__global__ void k_slab_max(const double* __restrict__ gathered, int nranks, double* out) {   // pairs (absolute, relative)
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double m0 = 0.0, m1 = 0.0;
  for (int r = 0; r < nranks; ++r) { m0 = nmax(m0, gathered[2 * r]); m1 = nmax(m1, gathered[2 * r + 1]); }
  out[0] = m0;
  out[1] = m1;
}

// True residual of the accepted solution (verification of the recursive BiCGSTAB residual), two max-norms over batch
// and cells: out_max[0] = max |rhs - A d| (reported), out_max[1] = max |rhs - A d|_K / (1 + |d_K| + sum_L |w_KL d_L|),
// the componentwise backward error the commit decision uses - when a cell jumps by orders of magnitude in one step
// (|d| >> 1: a tiny cell next to O(1) Dirichlet data) round-off alone puts |rhs - A d| at eps |A| |d| >> tol.
__global__ void __launch_bounds__(kBlock) k_true_residual(KrylovVecs k, size_t N, int ny, double* out_max) {
  __shared__ double red[32];
  const Scal* sc = k.scal + blockIdx.y;
  const size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t base = (size_t)blockIdx.y * N;
  double rabs = 0.0, rrel = 0.0;
  if (n >= k.n_begin && n < k.n_end) {
    const size_t g = base + n;
    double ax = 0.0, scale = 1.0;
    if (sc->it > 0) {
      const double* x = k.x + base;
      const double xW = x[n >= (size_t)ny ? n - ny : n], xE = x[n + ny < N ? n + ny : n], xS = x[n > 0 ? n - 1 : n], xN = x[n + 1 < N ? n + 1 : n];
      const double wW = k.wW[g], wE = k.wE[g], wS = k.wS[g], wN = k.wN[g];
      ax = x[n] + ((wW * xW + wE * xE) + (wS * xS + wN * xN));   // stencil_apply
      scale += fabs(x[n]) + ((fabs(wW * xW) + fabs(wE * xE)) + (fabs(wS * xS) + fabs(wN * xN)));
    }
    rabs = fabs(k.rhs[g] - ax);
    rrel = rabs / scale;
  }
  const double bmax = block_max(rabs, red);
  const double bmax_rel = block_max(rrel, red);
  if (threadIdx.x == 0) {
    atomicMax(reinterpret_cast<unsigned long long*>(out_max), (unsigned long long)__double_as_longlong(bmax));
    atomicMax(reinterpret_cast<unsigned long long*>(out_max + 1), (unsigned long long)__double_as_longlong(bmax_rel));
  }
}

struct StepStats {  // device-resident, copied back with the step result
  double fmin;
  unsigned long long negatives;
  double resid_max;      // max |rhs - A d| (true residual, absolute)
  double resid_rel_max;  // max componentwise backward error |rhs - A d|_K / (1 + |d_K| + sum |w d|): what a step is committed on
  int it_max;
  int n_bad;  // problems that ended in state 2/3
  int it_total_max;                // engine B: max over problems of iterations summed over the call's steps
  int steps_min;                   // engine B: smallest number of time steps of the call a problem completed (< nsteps only after a failure)
  unsigned long long it_sum_all;   // engine B: iterations summed over problems and steps
};

// KF: f^{n+1} = c (1 + d); yprev = clamp(f^{n+1}/f^n); statistics.           40 B/cell
// The step is COMMITTED only when the true residual of the solve (st->resid_rel_max, the componentwise backward error
// written by k_true_residual earlier on the stream; the max over the batch) is finite and <= resid_limit: otherwise f and yprev are left
// untouched and every problem counts as bad (the recursive BiCGSTAB residual alone can drift from the true one).
__global__ void __launch_bounds__(kBlock) k_finish(const double* __restrict__ x, const double* __restrict__ cs,
                                                   double* __restrict__ f, double* __restrict__ yprev, double* __restrict__ ylast,
                                                   const Scal* __restrict__ scal, size_t N, int predictor,
                                                   StepStats* st, double resid_limit) {
  const Scal* sc = scal + blockIdx.y;
  const size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool commit = st->resid_rel_max <= resid_limit;   // false for NaN
  double fmin = 1.0e300;
  int neg = 0;
  if (n < N) {
    const size_t g = (size_t)blockIdx.y * N + n;
    const double d = sc->it > 0 ? x[g] : 0.0;
    const double fold = f[g];
    const double fnew = commit ? cs[g] * (1.0 + d) : fold;
    if (commit) {
      f[g] = fnew;
      if (predictor) predictor_update(predictor, fnew, fold, yprev + g, ylast + g);
      if (!(fabs(fnew) <= 1.0e300)) atomicAdd(&st->n_bad, 1);   // non-finite f (cannot happen after a verified solve of finite data)
    }
    fmin = fnew;
    neg = fnew < 0.0;
  }
  double mn = warp_min(fmin);
  __shared__ double smin[32];
  __shared__ int sneg[32];
  int wn = __reduce_add_sync(0xffffffffu, neg);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) { smin[w] = mn; sneg[w] = wn; }
  __syncthreads();
  if (w == 0) {
    mn = warp_min(lane < nw ? smin[lane] : 1.0e300);
    wn = __reduce_add_sync(0xffffffffu, lane < nw ? sneg[lane] : 0);
    if (lane == 0) {
      if (wn) atomicAdd(&st->negatives, (unsigned long long)wn);
      // atomic min on a double through CAS (values may be negative)
      unsigned long long* addr = reinterpret_cast<unsigned long long*>(&st->fmin);
      unsigned long long old = *addr;
      while (mn < __longlong_as_double((long long)old)) {
        const unsigned long long assumed = old;
        old = atomicCAS(addr, assumed, (unsigned long long)__double_as_longlong(mn));
        if (old == assumed) break;
      }
      if (blockIdx.x == 0) {
        atomicMax(&st->it_max, sc->it);
        if (sc->state >= 2 || !commit) atomicAdd(&st->n_bad, 1);
      }
    }
  }
}

// A solve that stopped without converging commits nothing: only the statistics of the failure are collected.
__global__ void k_fail_stats(const Scal* __restrict__ scal, int nbatch, StepStats* st) {
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nbatch; b += gridDim.x * blockDim.x) {
    atomicMax(&st->it_max, scal[b].it);
    if (scal[b].state != 1) atomicAdd(&st->n_bad, 1);
  }
}

// Input check of sy2d_set_f / sy2d_put_f: the engine solves for the per-cell ratio f^{n+1} / (f^n yprev), so f must be
// finite and strictly positive (the reference's initial conditions add gEPS for the same reason - Albert_Young.h:39 - and
// the PPFV scheme keeps f positive).  Counts the offending cells.
__global__ void __launch_bounds__(kBlock) k_check_f(const double* __restrict__ f, size_t n_total, unsigned long long* bad) {
  unsigned long long mine = 0;
  for (size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x; n < n_total; n += (size_t)gridDim.x * blockDim.x) {
    const double v = f[n];
    mine += !(v > 0.0 && v <= 1.0e300);
  }
  if (mine) atomicAdd(bad, mine);
}

}  // namespace sy2d
