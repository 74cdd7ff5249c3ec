// Engine B: one persistent CTA per independent problem (ensemble mode and small grids).
//
// A CTA owns problem b for the whole call: for every time step it assembles the
// PPFV operator, runs the complete BiCGSTAB solve and updates f, with block-level
// barriers and block reductions only - no grid-wide synchronisation, no host
// round trip, no lockstep between problems (a member that converges in 40
// iterations does not wait for one that needs 200; the next queued CTA takes the
// SM).  The problem's working set (11 arrays x N x 8 B = 563 KB at 80x80) is
// touched only by this CTA, so with one CTA per SM the 148 concurrent working sets
// (83 MB) stay L2-resident and HBM sees each array about once per time step.
//
// Same arithmetic as the lockstep kernels of sy2d_kernels.cuh (assemble_row,
// scale_row, stencil_apply), same formulation A d = rhs, same stopping rule.
#pragma once
#include "sy2d_kernels.cuh"

namespace sy2d {

struct ProblemArgs {
  const double *tx, *ty, *cxy, *U, *Ud;  // read-only coefficients [nbatch][N]
  double *f, *yprev, *ylast, *cs;          // ylast: the last per-cell ratio (predictor 2 only)
  double *wW, *wE, *wS, *wN, *rhs;       // scaled operator (scratch, rewritten every step)
  double *x, *r, *p, *v, *s, *t;         // Krylov vectors (scratch)
  Scal* scal;                            // per-problem outcome
  StepStats* stats;                      // batch-wide statistics of the LAST step of the call
  Geometry g;
  double tol;
  int maxit, predictor, nsteps;
  // Scheduling: CTA b works on problem order[b]; the host sorts problems by the iterations they
  // needed in the previous call (longest first), so the last wave is filled with cheap problems
  // (longest-processing-time list scheduling: matters when nbatch is only a few waves, e.g. 512
  // members per GPU on 148 SMs).  cost[problem] receives this call's iteration total.
  const int* order;
  int* cost;
};

constexpr int kProblemThreads = 1024;
static_assert(kProblemThreads == 1024, "cta_allreduce pads partial warps with 0: keep 32 full warps");

// all-reduce of NS sums (+ optionally one max in slot NS) over the CTA; every thread gets the result
template <int NS, bool WITH_MAX>
__device__ __forceinline__ void cta_allreduce(double (&v)[NS + (WITH_MAX ? 1 : 0)], double* smem /* >= (NS+1)*32 */) {
  constexpr int NV = NS + (WITH_MAX ? 1 : 0);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < NS; ++k) v[k] = warp_sum(v[k]);
  if (WITH_MAX) v[NS] = warp_max(v[NS]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) smem[k * 32 + w] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NS; ++k) v[k] = warp_sum(lane < nw ? smem[k * 32 + lane] : 0.0);
  if (WITH_MAX) v[NS] = warp_max(lane < nw ? smem[NS * 32 + lane] : 0.0);
  __syncthreads();  // smem reusable
}

__global__ void __launch_bounds__(kProblemThreads, 1) k_problem_steps(ProblemArgs a) {
  __shared__ double red[4 * 32];
  const int nx = a.g.nx, ny = a.g.ny;
  const int N = nx * ny;
  const int prob = a.order ? a.order[blockIdx.x] : blockIdx.x;
  const size_t base = (size_t)prob * N;
  const int tid = threadIdx.x;
  const double* __restrict__ tx = a.tx + base;
  const double* __restrict__ ty = a.ty + base;
  const double* __restrict__ cxy = a.cxy + base;
  const double* __restrict__ U = a.U + base;
  const double* __restrict__ Ud = a.Ud + base;
  // arrays written inside this kernel: plain (coherent) accesses only
  double* f = a.f + base; double* yprev = a.yprev + base; double* cs = a.cs + base;
  double* wW = a.wW + base; double* wE = a.wE + base; double* wS = a.wS + base; double* wN = a.wN + base;
  double* rhs = a.rhs + base; double* x = a.x + base; double* r = a.r + base; double* p = a.p + base;
  double* v = a.v + base; double* s = a.s + base; double* t = a.t + base;

  int it_total = 0, it = 0, state = 1, steps_ok = 0;
  double rmax = 0.0, res_true = 0.0, res_rel = 0.0;
  for (int step = 0; step < a.nsteps; ++step) {
    // ---------------- assembly (Solver.cc:167-267 fused with :292-422) ----------------
    double acc[2] = {0.0, 0.0};
    for (int n = tid; n < N; n += kProblemThreads) {
      const int i = n / ny, j = n - i * ny;
      Row row;
      assemble_row(f, tx, ty, cxy, U, Ud, a.g, i, j, row);
      const int nW = i > 0 ? n - ny : n, nE = i < nx - 1 ? n + ny : n, nS = j > 0 ? n - 1 : n, nN = j < ny - 1 ? n + 1 : n;
      Scaled sc;
      scale_row(row, yprev[n], yprev[nW], yprev[nE], yprev[nS], yprev[nN], sc);
      wW[n] = sc.wW; wE[n] = sc.wE; wS[n] = sc.wS; wN[n] = sc.wN; rhs[n] = sc.rhs; cs[n] = sc.cs;
      acc[0] += sc.rhs * sc.rhs;
      acc[1] = nmax(acc[1], fabs(sc.rhs));
    }
    cta_allreduce<1, true>(acc, red);
    double rho = acc[0];
    rmax = acc[1];
    double alpha = 1.0, omega = 1.0, beta = 0.0;
    bool first = true;
    it = 0;
    state = (rmax <= a.tol) ? 1 : 0;
    // ---------------- BiCGSTAB on A d = rhs, d0 = 0, rhat = rhs ----------------
    while (state == 0) {
      for (int n = tid; n < N; n += kProblemThreads) p[n] = first ? rhs[n] : r[n] + beta * (p[n] - omega * v[n]);
      __syncthreads();
      double a1[1] = {0.0};
      for (int n = tid; n < N; n += kProblemThreads) {
        const double vv = stencil_apply(p, (size_t)n, (size_t)N, ny, p[n], wW[n], wE[n], wS[n], wN[n]);
        v[n] = vv;
        a1[0] += rhs[n] * vv;
      }
      cta_allreduce<1, false>(a1, red);
      alpha = a1[0] != 0.0 ? rho / a1[0] : 0.0;
      for (int n = tid; n < N; n += kProblemThreads) s[n] = (first ? rhs[n] : r[n]) - alpha * v[n];
      __syncthreads();
      double a2[2] = {0.0, 0.0};
      for (int n = tid; n < N; n += kProblemThreads) {
        const double sn = s[n];
        const double tt = stencil_apply(s, (size_t)n, (size_t)N, ny, sn, wW[n], wE[n], wS[n], wN[n]);
        t[n] = tt;
        a2[0] += tt * sn;
        a2[1] += tt * tt;
      }
      cta_allreduce<2, false>(a2, red);
      omega = a2[1] > 0.0 ? a2[0] / a2[1] : 0.0;
      double a3[2] = {0.0, 0.0};
      for (int n = tid; n < N; n += kProblemThreads) {
        const double sn = s[n];
        x[n] = (first ? 0.0 : x[n]) + (alpha * p[n] + omega * sn);
        const double rn = sn - omega * t[n];
        r[n] = rn;
        a3[0] += rhs[n] * rn;
        a3[1] = nmax(a3[1], fabs(rn));
      }
      cta_allreduce<1, true>(a3, red);
      const double rho_new = a3[0];
      rmax = a3[1];
      ++it;
      first = false;
      if (rmax <= a.tol) state = 1;
      else if (!(rmax == rmax) || rho_new == 0.0 || omega == 0.0) state = 3;
      else if (it >= a.maxit) state = 2;
      beta = (rho_new / rho) * (alpha / omega);
      rho = rho_new;
    }
    it_total += it;
    if (state >= 2) break;   // stopped without converging (maxit, breakdown, NaN): f and yprev stay those of t^n
    // ---------------- true residual of the accepted solution ----------------
    const bool last = step == a.nsteps - 1;
    if (last) {
      double m = 0.0, mr = 0.0;   // absolute, and componentwise-relative (k_true_residual) true residual
      for (int n = tid; n < N; n += kProblemThreads) {
        double ax = 0.0, scale = 1.0;
        if (it > 0) {
          const double xW = x[n >= ny ? n - ny : n], xE = x[n + ny < N ? n + ny : n], xS = x[n > 0 ? n - 1 : n], xN = x[n + 1 < N ? n + 1 : n];
          ax = x[n] + ((wW[n] * xW + wE[n] * xE) + (wS[n] * xS + wN[n] * xN));
          scale += fabs(x[n]) + ((fabs(wW[n] * xW) + fabs(wE[n] * xE)) + (fabs(wS[n] * xS) + fabs(wN[n] * xN)));
        }
        const double ra = fabs(rhs[n] - ax);
        m = nmax(m, ra);
        mr = nmax(mr, ra / scale);
      }
      double mm[1] = {m};
      cta_allreduce<0, true>(mm, red);
      res_true = mm[0];
      mm[0] = mr;
      cta_allreduce<0, true>(mm, red);
      res_rel = mm[0];
    }
    // ---------------- f^{n+1} = c (1 + d), predictor, statistics ----------------
    double fmin_l = 1.0e300;
    int neg = 0;
    for (int n = tid; n < N; n += kProblemThreads) {
      const double fold = f[n];
      const double fnew = cs[n] * (1.0 + (it > 0 ? x[n] : 0.0));
      f[n] = fnew;
      if (a.predictor) predictor_update(a.predictor, fnew, fold, yprev + n, a.ylast + base + n);
      fmin_l = ::fmin(fmin_l, fnew);
      neg += fnew < 0.0;
    }
    if (last) {
      double mm[2] = {(double)neg, -fmin_l};
      cta_allreduce<1, true>(mm, red);  // sum of negatives, max of -f  (mm[1] >= -1e300)
      if (tid == 0) {
        if (mm[0] > 0.0) atomicAdd(&a.stats->negatives, (unsigned long long)mm[0]);
        const double mn = -mm[1];
        unsigned long long* addr = reinterpret_cast<unsigned long long*>(&a.stats->fmin);
        unsigned long long old = *addr;
        while (mn < __longlong_as_double((long long)old)) {
          const unsigned long long assumed = old;
          old = atomicCAS(addr, assumed, (unsigned long long)__double_as_longlong(mn));
          if (old == assumed) break;
        }
      }
    }
    __syncthreads();  // f complete before the next step's assembly reads neighbours
    steps_ok = step + 1;
  }
  if (tid == 0) {
    Scal* sc = a.scal + prob;
    if (a.cost) a.cost[prob] = it_total;
    sc->it = it;
    sc->state = state;
    sc->rmax = rmax;
    atomicMax(&a.stats->it_max, it);
    atomicMax(&a.stats->it_total_max, it_total);
    atomicAdd(&a.stats->it_sum_all, (unsigned long long)it_total);
    atomicMax(reinterpret_cast<unsigned long long*>(&a.stats->resid_max), (unsigned long long)__double_as_longlong(res_true));
    atomicMax(reinterpret_cast<unsigned long long*>(&a.stats->resid_rel_max), (unsigned long long)__double_as_longlong(res_rel));
    atomicMin(&a.stats->steps_min, steps_ok);
    if (state >= 2 || !(res_rel == res_rel)) atomicAdd(&a.stats->n_bad, 1);
  }
}

}  // namespace sy2d
