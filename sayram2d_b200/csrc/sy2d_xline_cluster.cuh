// Engine 2, x-line variant on a thread-block CLUSTER: one problem per pair of CTAs (two SMs), every array of the
// iteration on chip.
//
// The single-CTA kernel (sy2d_xline_kernel.cuh) keeps five of a problem's arrays (wS, wN, v, y, rhat) in an L2-backed
// scratch because one SM cannot hold them: 80 B of L2 traffic per cell and iteration, and the L2 latency of those
// loads is what its 20 warps wait for.  Here the problem is cut in two along j (the direction WITHOUT the line solve):
// CTA `rank` of the pair owns the columns j = rank*CW .. rank*CW + CW-1 of all nx rows, 16 lanes per column with R rows
// each (80 x 80: R = 5, 640 threads per CTA as before, but half the cells per thread).  Per CTA
//   registers : r/s, the Thomas work vector z (phat / shat / t), v, and the factors l', e of the x-lines (5 x R doubles)
//   shared    : hat (phat / shat in natural layout, with one halo column on each side), p, y, wS', wN', rhat
// and nothing of the iteration touches global memory.  What crosses the pair:
//   * the S/N exchange of the boundary columns: the owner writes its phat / shat of column CW-1 (rank 0) or 0 (rank 1)
//     into the halo column of the partner's hat through distributed shared memory (80 doubles per publish);
//   * the three reductions of an iteration: every warp writes its partial into BOTH CTAs' buffers, a cluster barrier,
//     and both CTAs add the 2 x NW partials in the same order - bitwise the same scalars, so both take the same
//     decisions (convergence, breakdown) without any further exchange;
//   * per time step, the S faces of column CW (rank 1) that are the N faces of column CW-1 (rank 0): read remotely.
// Every CTA barrier of the single-CTA iteration becomes a cluster barrier (barrier.cluster arrive.release /
// wait.acquire); f, yprev, ylast live in global memory between time steps as before (the work queue hands a problem
// from pair to pair).  Same arithmetic per cell as the single-CTA kernel (face expressions, scaling, pivot scaling,
// BiCGSTAB recurrences); the reductions add in a different order, so iterates agree to rounding, not bitwise.
#pragma once
#include <cooperative_groups.h>

#include "sy2d_assemble_tma.cuh"   // smem_u32, mbar_init
#include "sy2d_xline_kernel.cuh"

namespace sy2d {

namespace cg = cooperative_groups;


__device__ __forceinline__ void cl_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cl_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// Full cluster barrier with release / acquire semantics (ptxas: MEMBAR.ALL.GPU + UCGABAR + CCTL.IVALL, microseconds):
// used a few times per time step, where GLOBAL memory (f, yprev) or plain remote loads have to be ordered.  The five
// synchronisation points of an iteration use the mbarrier path below instead.
__device__ __forceinline__ void cl_sync() {
  __syncwarp();
  cl_arrive();
  cl_wait();
}

// Remote store that carries its own completion: the value lands in the partner's shared memory and `bytes` are
// subtracted from the transaction count of the partner's mbarrier - no fence, no cluster barrier.
__device__ __forceinline__ unsigned cl_mapa(unsigned addr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cl_st_async(unsigned raddr, double v, unsigned rmbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(raddr), "l"(__double_as_longlong(v)), "r"(rmbar)
               : "memory");
}
__device__ __forceinline__ void cl_expect_tx(unsigned mbar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cl_mbar_wait(unsigned mbar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "SY2D_CL_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra SY2D_CL_DONE;\n"
      "bra SY2D_CL_WAIT;\n"
      "SY2D_CL_DONE:\n"
      "}\n" ::"r"(mbar), "r"(parity)
      : "memory");
}

// The partials of the 2 x NW warps of the pair sit in slots [rank * NW + w] of BOTH CTAs' buffers; every warp adds them in
// the same order.
template <int NW>
__device__ __forceinline__ double cl_slots(const double* red, int lane) {
  static_assert(2 * NW <= 64, "two slots per lane at most");
  double x = lane < 2 * NW ? red[lane] : 0.0;
  if (lane + 32 < 2 * NW) x += red[lane + 32];
  return x;
}
template <int NW>
__device__ __forceinline__ double cl_slots_max(const double* red, int lane) {
  double x = lane < 2 * NW ? red[lane] : 0.0;
  if (lane + 32 < 2 * NW) x = nmax(x, red[lane + 32]);
  return x;
}

// One reduction point of the pair: its own partial buffer (`red` here, `red_r` = the same buffer in the partner, as a
// shared::cluster address) and its own mbarrier (`mb` here, `mb_r` in the partner).  A warp writes its partials into both
// buffers; the local ones are ordered by the CTA barrier, the partner's arrive as transactions on `mb`.
struct ClPoint {
  double* red;
  unsigned red_r, mb, mb_r;
};

// NS sums followed by NM maxima (of non-negative values) over the pair; every thread of both CTAs gets the result.
template <int NW, int NS, int NM>
__device__ __forceinline__ void cl_reduce(double (&v)[NS + NM], const ClPoint& pt, unsigned parity, int slot, int lane) {
#pragma unroll
  for (int q = 0; q < NS + NM; ++q) v[q] = q < NS ? warp_sum(v[q]) : warp_max(v[q]);
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < NS + NM; ++q) {
      pt.red[q * 64 + slot] = v[q];
      cl_st_async(pt.red_r + (unsigned)(q * 64 + slot) * 8u, v[q], pt.mb_r);
    }
  }
  if (threadIdx.x == 0) cl_expect_tx(pt.mb, (unsigned)(NW * (NS + NM) * 8));
  __syncthreads();
  cl_mbar_wait(pt.mb, parity);
#pragma unroll
  for (int q = 0; q < NS + NM; ++q)
    v[q] = q < NS ? warp_sum(cl_slots<NW>(pt.red + q * 64, lane)) : warp_max(cl_slots_max<NW>(pt.red + q * 64, lane));
}

// R rows per lane, NT threads per CTA, HS = row stride of hat (odd multiple R*HS mod 16 => the 16 lanes of a column
// hit 16 distinct 8-byte banks).  Full tiles only: nx = 16 R, ny = 2 * (NT / 16).
template <int R, int NCH, int NT, int HS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NT, 1) k_problem_xline_cl(XlineArgs xa) {
  constexpr int CPW = 32 / NCH;          // columns per warp
  // 320 threads are 10 warps, three of them on one scheduler: 16384 / 96 = 170 registers per thread at most (ptxas: 168).
  // l', e, r/s, z, v (100 registers at R = 10) fit; wS' next to them spills around every line solve.
  constexpr bool WREG = false;
  constexpr int NW = NT / 32;
  constexpr int CW = NW * CPW;   // columns per CTA
  constexpr int nx = NCH * R, ny = 2 * CW, N = nx * ny;
  constexpr int S = R * NT;
  static_assert(HS >= CW + 2 && nx * HS >= S, "hat holds the halo columns and doubles as a face buffer");
  extern __shared__ double sm[];
  const ProblemArgs& a = xa.a;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int k = lane % NCH, jj = lane / NCH;
  const int jl = w * CPW + jj, j = rank * CW + jl;
  const int i0 = k * R;
  double* hat = sm;             // [nx][HS]: local column jl at offset jl + 1, offsets 0 and CW + 1 are the partner's boundary columns
  double* l_s = hat + nx * HS;  // assembly: vertex line, then raw wW; iteration: y
  double* e_s = l_s + S;        // assembly: raw wE
  double* p_s = e_s + S;        // assembly: A_L of the south faces; iteration: p, then p - omega v
  double* wS_s = p_s + S;
  double* wN_s = wS_s + S;
  double* rh_s = wN_s + S;      // rhat = rhs'
  double* red = rh_s + S;       // three buffers of 2 x 64 partials
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(red + 384);   // [4]: halo exchange, three reduction points
  int* s_item = reinterpret_cast<int*>(mbar + 4);
  double* y_s = l_s;
  const double* hat_r = cluster.map_shared_rank(hat, rank ^ 1);   // plain remote loads (assembly, once per time step)
  const double* p_r = cluster.map_shared_rank(p_s, rank ^ 1);
  int* s_item_r = cluster.map_shared_rank(s_item, rank ^ 1);
  const unsigned full = 0xffffffffu;
  const int slot = rank * NW + w;
  const bool send_up = rank == 0 && jl == CW - 1;   // my column is the partner's south halo
  const bool send_dn = rank == 1 && jl == 0;        // my column is the partner's north halo
  const bool halo_reader = send_up || send_dn;      // ... and the partner's boundary column is my halo
  const unsigned mbP = smem_u32(mbar), mbP_r = cl_mapa(mbP, rank ^ 1);
  const unsigned hat_r32 = cl_mapa(smem_u32(hat), rank ^ 1);
  ClPoint pt[3];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    pt[q].red = red + 128 * q;
    pt[q].red_r = cl_mapa(smem_u32(red + 128 * q), rank ^ 1);
    pt[q].mb = smem_u32(mbar + 1 + q);
    pt[q].mb_r = cl_mapa(pt[q].mb, rank ^ 1);
  }
  unsigned ph = 0;   // phase parities: bit 0 halo exchange, bits 1-3 the reduction points

  for (int n = tid; n < nx * HS; n += NT) hat[n] = 0.0;   // halo columns at the domain boundary are read (times a zero weight)
  if (tid == 0) {
    for (int q = 0; q < 4; ++q) mbar_init(mbar + q, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cl_sync();   // the partner's mbarriers exist before anything is sent to them

  double rs[R], z[R], v[R], lr[R], er[R];
  constexpr bool WNREG = false;   // wN' in registers as well: spills five values around every line solve
  constexpr bool RHREG = false;   // rhat in registers too: 8 arrays of R doubles do not fit in 200 registers without spills
  double wSr[WREG ? R : 1] = {}, wNr[WNREG ? R : 1] = {}, rhr[RHREG ? R : 1] = {};
  auto wS_of = [&](int m) { return WREG ? wSr[WREG ? m : 0] : wS_s[m * NT + tid]; };
  auto wN_of = [&](int m) { return WNREG ? wNr[WNREG ? m : 0] : wN_s[m * NT + tid]; };
  auto rh_of = [&](int m) { return RHREG ? rhr[RHREG ? m : 0] : rh_s[m * NT + tid]; };

  for (;;) {   // work items: (problem, chunk of time steps)
    if (rank == 0 && tid == 0) {
      const int ticket = atomicAdd(&xa.q->head, 1);
      int item = -1;
      while (ticket < ld_acquire_gpu(&xa.q->total)) {
        item = ld_acquire_gpu(xa.slots + ticket);
        if (item >= 0) break;
        __nanosleep(200);
      }
      const int sb = item >= 0 ? ld_acquire_gpu(xa.steps_done + item) : 0;
      s_item[0] = item; s_item[1] = sb;
      s_item_r[0] = item; s_item_r[1] = sb;
    }
    cl_sync();
    const int item = s_item[0];
    if (item < 0) break;
    const int prob = item;
    const size_t base = (size_t)prob * N;
    const double* __restrict__ tx = a.tx + base;
    const double* __restrict__ ty = a.ty + base;
    const double* __restrict__ cxy = a.cxy + base;
    const double* __restrict__ U = a.U + base;
    const double* __restrict__ Ud = a.Ud + base;
    double* f = a.f + base;
    double* yprev = a.yprev + base;
    const int step_begin = s_item[1];
    const int step_end = min(step_begin + xa.chunk, a.nsteps);
    if (xa.hin && step_begin == 0) {   // first item of the problem: f from the host buffer, half per CTA
      const double2* src = reinterpret_cast<const double2*>(xa.hin + base);
      for (int n = rank * NT + tid; n < N / 2; n += 2 * NT) reinterpret_cast<double2*>(f)[n] = src[n];
      __threadfence();
      cl_sync();
    }
    int it_total = 0, it = 0, state = 1, steps_ok = step_begin;
    double rmax = 0.0, res_true = 0.0, res_rel = 0.0;

    for (int step = step_begin; step < step_end; ++step) {
      // ------------- assembly (same face-once scheme as the single-CTA kernel's full tile) -------------
      const Geometry& g = a.g;
      double* SK_s = hat;    // A_K, A_L of the south face of slot q
      double* SL_s = p_s;
      auto vertex_at = [&](int vi, int vj) {
        const int il = vi > 0 ? vi - 1 : 0, ih = vi < nx ? vi : nx - 1;
        const int jb = vj > 0 ? vj - 1 : 0, jh = vj < ny ? vj : ny - 1;
        return vertex_value(g, vi, vj, f[il * ny + jb], f[ih * ny + jb], f[il * ny + jh], f[ih * ny + jh]);
      };
      double vprev = vertex_at(i0, j);
#pragma unroll 1
      for (int m = 0; m < R; ++m) {
        const int n = (i0 + m) * ny + j;
        const double vnext = vertex_at(i0 + m + 1, j);
        double AK = 0.0, AL = 0.0;
        if (j > 0) {   // south face: K = (i, j), L = (i, j-1)
          const double tyP = ty[n], cP = cxy[n], t = ty[n - 1], c = cxy[n - 1];
          const double aS_A = tyP + cP, aS_B = tyP - cP, lA = t + c, lB = t - c;
          face_pair(aS_A * vprev + aS_B * vnext, aS_A + aS_B, f[n], lA * vnext + lB * vprev, lA + lB, f[n - 1], AK, AL);
        }
        SK_s[m * NT + tid] = AK;
        SL_s[m * NT + tid] = AL;
        l_s[m * NT + tid] = vprev;
        vprev = vnext;
      }
      const double vtop = vprev;   // V(i0 + R, j)
      cl_sync();                   // south faces of column j + 1 complete (rank 0, last column: in the partner's memory)
      {
        const int jn = j + 1;
        const int jln = jl + 1 < CW ? jl + 1 : 0;
        const int tidN = (jln / CPW) * 32 + (jln % CPW) * NCH + k;   // same rows, column j + 1 (in the partner for jl = CW - 1)
        const bool remote = rank == 0 && jl == CW - 1;
        const double* SKn = remote ? hat_r : SK_s;
        const double* SLn = remote ? p_r : SL_s;
        double vNW = vertex_at(i0, jn);
        double AKw = 0.0, ALw = 0.0;   // west face of the current row: K = current cell, L = the row below
        if (i0 > 0) {
          const int n = i0 * ny + j;
          const double v0 = l_s[tid];
          const double txP = tx[n], cP = cxy[n], t = tx[n - ny], c = cxy[n - ny];
          const double aW_A = txP - cP, aW_B = txP + cP, lA = t - c, lB = t + c;
          face_pair(aW_A * vNW + aW_B * v0, aW_A + aW_B, f[n], lA * v0 + lB * vNW, lA + lB, f[n - ny], AKw, ALw);
        }
#pragma unroll 1
        for (int m = 0; m < R; ++m) {
          const int i = i0 + m, n = i * ny + j, q = m * NT + tid;
          const double vSW = l_s[q], vSE = m + 1 < R ? l_s[q + NT] : vtop, vNE = vertex_at(i + 1, jn);
          const double txP = tx[n], tyP = ty[n], cP = cxy[n], f00 = f[n];
          Row row;
          double diag = 0.0, Rr = 0.0;
          row.oW = 0.0; row.oE = 0.0; row.oS = 0.0; row.oN = 0.0;
          if (i > 0) { diag += AKw; row.oW = -ALw; }
          else if (g.bc[0] == 0) diag += dirichlet_face((txP - cP) * vNW + (txP + cP) * vSW, (txP - cP) + (txP + cP), f00, Rr);
          double AKe = 0.0, ALe = 0.0;
          if (i < nx - 1) {   // east face: K = (i+1, j), L = this cell
            const double t = tx[n + ny], c = cxy[n + ny];
            const double kA = t - c, kB = t + c, aE_A = txP - cP, aE_B = txP + cP;
            face_pair(kA * vNE + kB * vSE, kA + kB, f[n + ny], aE_A * vSE + aE_B * vNE, aE_A + aE_B, f00, AKe, ALe);
            diag += ALe;
            row.oE = -AKe;
          } else if (g.bc[1] == 0) {
            diag += dirichlet_face((txP - cP) * vSE + (txP + cP) * vNE, (txP - cP) + (txP + cP), f00, Rr);
          }
          if (j > 0) { diag += SK_s[q]; row.oS = -SL_s[q]; }
          else if (g.bc[2] == 0) diag += dirichlet_face((tyP + cP) * vSW + (tyP - cP) * vSE, (tyP + cP) + (tyP - cP), f00, Rr);
          if (j < ny - 1) { diag += SLn[m * NT + tidN]; row.oN = -SKn[m * NT + tidN]; }
          else if (g.bc[3] == 0) diag += dirichlet_face((tyP + cP) * vNE + (tyP - cP) * vNW, (tyP + cP) + (tyP - cP), f00, Rr);
          diag += Ud[n];
          Rr += U[n] * f00;
          row.diag = diag; row.R = Rr; row.f00 = f00;
          const int nW = i > 0 ? n - ny : n, nE = i < nx - 1 ? n + ny : n, nS = j > 0 ? n - 1 : n, nN = j < ny - 1 ? n + 1 : n;
          row.fW = f[nW]; row.fE = f[nE]; row.fS = f[nS]; row.fN = f[nN];
          Scaled sc;
          scale_row(row, yprev[n], yprev[nW], yprev[nE], yprev[nS], yprev[nN], sc);
          l_s[q] = sc.wW;   // raw wW, wE until the factorisation below
          e_s[q] = sc.wE;
          wS_s[q] = sc.wS; wN_s[q] = sc.wN; rh_s[q] = sc.rhs;
          AKw = AKe; ALw = ALe; vNW = vNE;
        }
      }
      // LU of T down each column (chain over the 16 lanes of the column), factors into registers:
      // d_i = 1 - wW_i e_{i-1}, l' = wW / d, e = wE / d; then the pivot scaling of the rest of the row.
      double acc[2] = {0.0, 0.0};
      {
        double dv[R];
#pragma unroll
        for (int m = 0; m < R; ++m) { lr[m] = l_s[m * NT + tid]; er[m] = e_s[m * NT + tid]; dv[m] = 1.0; }
        double elast = 0.0;
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
          const double ein = __shfl_up_sync(full, elast, 1, NCH);
          if (k == c) {
            double eprev = k == 0 ? 0.0 : ein;   // wW of the first row of a column is 0
#pragma unroll
            for (int m = 0; m < R; ++m) {
              const double dinv = sy2d_div(1.0, 1.0 - lr[m] * eprev);
              eprev = er[m] * dinv;
              lr[m] *= dinv; dv[m] = dinv; er[m] = eprev;
            }
            elast = eprev;
          }
        }
#pragma unroll
        for (int m = 0; m < R; ++m) {
          const int q = m * NT + tid;
          const double rh = rh_s[q] * dv[m];
          if (WREG) {
            wSr[WREG ? m : 0] = wS_s[q] * dv[m];
          } else {
            wS_s[q] *= dv[m];
          }
          if (WNREG) wNr[WNREG ? m : 0] = wN_s[q] * dv[m];
          else wN_s[q] *= dv[m];
          if (RHREG) rhr[RHREG ? m : 0] = rh;
          else rh_s[q] = rh;
          rs[m] = rh;
          acc[0] += rh * rh;
          acc[1] = nmax(acc[1], fabs(rh));
        }
      }
      // (this barrier also ends the face exchange through hat and the p region, and the use of l_s as a staging area)
      cl_reduce<NW, 1, 1>(acc, pt[2], (ph >> 3) & 1u, slot, lane);
      ph ^= 8u;
      double rho = acc[0];
      rmax = acc[1];
      double alpha = 1.0, omega = 1.0, beta = 0.0;
      bool first = true;
      it = 0;
      state = (rmax <= a.tol) ? 1 : 0;

      // z <- T^-1 b as a partitioned solve over the 16 lanes of the column (see the single-CTA kernel), factors in registers
      auto tsolve = [&](auto bget) {
        double A = 0.0, B = 1.0;
#pragma unroll
        for (int m = 0; m < R; ++m) {
          A = bget(m) - lr[m] * A;
          z[m] = A;
          B = -lr[m] * B;
        }
#pragma unroll
        for (int d = 1; d < NCH; d <<= 1) {
          const double Au = __shfl_up_sync(full, A, d, NCH), Bu = __shfl_up_sync(full, B, d, NCH);
          if (k >= d) { A = A + B * Au; B = B * Bu; }
        }
        double cin = __shfl_up_sync(full, A, 1, NCH);
        if (k == 0) cin = 0.0;
        double P = 1.0;
#pragma unroll
        for (int m = 0; m < R; ++m) {
          P = -lr[m] * P;
          z[m] += P * cin;
        }
        A = 0.0; B = 1.0;
#pragma unroll
        for (int m = R - 1; m >= 0; --m) {
          A = z[m] - er[m] * A;
          z[m] = A;
          B = -er[m] * B;
        }
#pragma unroll
        for (int d = 1; d < NCH; d <<= 1) {
          const double Ad = __shfl_down_sync(full, A, d, NCH), Bd = __shfl_down_sync(full, B, d, NCH);
          if (k + d < NCH) { A = A + B * Ad; B = B * Bd; }
        }
        cin = __shfl_down_sync(full, A, 1, NCH);
        if (k == NCH - 1) cin = 0.0;
        P = 1.0;
#pragma unroll
        for (int m = R - 1; m >= 0; --m) {
          P = -er[m] * P;
          z[m] += P * cin;
        }
      };
      // z into hat (natural layout) for the S/N neighbours; the boundary columns also into the partner's halo column
      auto publish = [&]() {
#pragma unroll
        for (int m = 0; m < R; ++m) hat[(i0 + m) * HS + jl + 1] = z[m];
        if (halo_reader) {   // my boundary column is the partner's halo column (rank 0 -> its offset 0, rank 1 -> its offset CW + 1)
          const unsigned dst = hat_r32 + (unsigned)((i0 * HS + (send_up ? 0 : CW + 1)) * 8);
#pragma unroll
          for (int m = 0; m < R; ++m) cl_st_async(dst + (unsigned)(m * HS * 8), z[m], mbP_r);
          if (k == 0) cl_expect_tx(mbP, (unsigned)(nx * 8));   // ... and the partner's boundary column arrives in mine
        }
        __syncthreads();
        if (halo_reader) cl_mbar_wait(mbP, ph & 1u);
        ph ^= 1u;
      };
      const double* hS = hat + i0 * HS + jl;       // S neighbour of row i0 (local column jl - 1); N neighbour: + 2

      while (state == 0) {
        // p = r + beta (p - omega v): the last iteration left p - omega v in the p array
        tsolve([&](int m) {
          const int q = m * NT + tid;
          const double pm = first ? rs[m] : rs[m] + beta * p_s[q];
          p_s[q] = pm;
          v[m] = pm;
          return pm;
        });
        publish();
        // v = p + wS phat_S + wN phat_N ; (rhat, v)
        double a1[1] = {0.0};
#pragma unroll
        for (int m = 0; m < R; ++m) {
          const double val = v[m] + (wS_of(m) * hS[m * HS] + wN_of(m) * hS[m * HS + 2]);
          v[m] = val;
          a1[0] += rh_of(m) * val;
        }
        cl_reduce<NW, 1, 0>(a1, pt[0], (ph >> 1) & 1u, slot, lane);
        ph ^= 2u;
        alpha = a1[0] != 0.0 ? rho / a1[0] : 0.0;
#pragma unroll
        for (int m = 0; m < R; ++m) rs[m] -= alpha * v[m];   // s
        tsolve([&](int m) { return rs[m]; });
        publish();
        // t = s + wS shat_S + wN shat_N ; (t,s), (t,t)   (t reuses z)
        double a2[2] = {0.0, 0.0};
#pragma unroll
        for (int m = 0; m < R; ++m) {
          const double val = rs[m] + (wS_of(m) * hS[m * HS] + wN_of(m) * hS[m * HS + 2]);
          z[m] = val;
          a2[0] += val * rs[m];
          a2[1] += val * val;
        }
        cl_reduce<NW, 2, 0>(a2, pt[1], (ph >> 2) & 1u, slot, lane);
        ph ^= 4u;
        omega = a2[1] > 0.0 ? a2[0] / a2[1] : 0.0;
        // y += alpha p + omega s ; p <- p - omega v ; r = s - omega t ; (rhat, r), max|r|
        double a3[2] = {0.0, 0.0};
#pragma unroll
        for (int m = 0; m < R; ++m) {
          const int q = m * NT + tid;
          const double pm = p_s[q];
          y_s[q] = (first ? 0.0 : y_s[q]) + (alpha * pm + omega * rs[m]);
          p_s[q] = pm - omega * v[m];
          rs[m] -= omega * z[m];
          a3[0] += rh_of(m) * rs[m];
          a3[1] = nmax(a3[1], fabs(rs[m]));
        }
        cl_reduce<NW, 1, 1>(a3, pt[2], (ph >> 3) & 1u, slot, lane);
        ph ^= 8u;
        const double rho_new = a3[0];
        rmax = a3[1];
        ++it;
        first = false;
        if (rmax <= a.tol) state = 1;
        else if (!(rmax == rmax) || !(rho_new == rho_new) || rho_new == 0.0 || omega == 0.0) state = 3;
        else if (it >= a.maxit) state = 2;
        beta = (rho_new / rho) * (alpha / omega);
        rho = rho_new;
      }
      it_total += it;
      // x = T^-1 y  (left in z)
      if (it > 0) {
        tsolve([&](int m) { return y_s[m * NT + tid]; });
      } else {
#pragma unroll
        for (int m = 0; m < R; ++m) z[m] = 0.0;
      }

      // ------------- f^{n+1} = c (1 + d) ; predictor ; true residual of the last step -------------
      const bool last = step == a.nsteps - 1;
      if (last) {
        // r'_i = rhs'_i - (l'_i x_W + (1/d_i) x_i + e_i x_E + wS'_i x_S + wN'_i x_N),  1/d_i = 1 + l'_i e_{i-1}
        publish();
        const double e_below = __shfl_up_sync(full, er[R - 1], 1, NCH);   // e of the last row of the lane below
        double mm[2] = {0.0, 0.0};   // absolute, componentwise-relative
#pragma unroll
        for (int m = 0; m < R; ++m) {
          const int i = i0 + m;
          const double lp = lr[m];
          const double dinv_i = i > 0 ? 1.0 + lp * (m > 0 ? er[m > 0 ? m - 1 : 0] : e_below) : 1.0;
          const double dW = i > 0 ? hS[(m - 1) * HS + 1] : 0.0, dE = i < nx - 1 ? hS[(m + 1) * HS + 1] : 0.0;
          const double tW = lp * dW, tE = er[m] * dE, tS = wS_of(m) * hS[m * HS], tN = wN_of(m) * hS[m * HS + 2];
          const double ax = dinv_i * z[m] + ((tW + tE) + (tS + tN));
          const double ra = fabs(rh_of(m) - ax);
          mm[0] = nmax(mm[0], ra / dinv_i);
          mm[1] = nmax(mm[1], ra / (dinv_i * (1.0 + fabs(z[m])) + ((fabs(tW) + fabs(tE)) + (fabs(tS) + fabs(tN)))));
        }
        cl_reduce<NW, 0, 2>(mm, pt[0], (ph >> 1) & 1u, slot, lane);
        ph ^= 2u;
        res_true = mm[0];
        res_rel = mm[1];
      }
      if (state >= 2) break;   // the solve stopped without converging: f and yprev stay those of t^n
      double fneg = 0.0, fmin_neg = -1.0e300;
#pragma unroll
      for (int m = 0; m < R; ++m) {
        const int n = (i0 + m) * ny + j;
        const double fold = f[n];
        const double yp = yprev[n];
        const double fnew = (fold * yp) * (1.0 + z[m]);
        f[n] = fnew;
        if (a.predictor) predictor_update(a.predictor, fnew, fold, yprev + n, a.ylast + base + n);
        fneg += fnew < 0.0 ? 1.0 : 0.0;
        fmin_neg = nmax(fmin_neg, -fnew);
      }
      if (last) {
        double mm[2] = {fneg, fmin_neg + 1.0e300};   // max of (-f), shifted to be non-negative
        cl_reduce<NW, 1, 1>(mm, pt[1], (ph >> 2) & 1u, slot, lane);
        ph ^= 4u;
        if (rank == 0 && tid == 0) {
          if (mm[0] > 0.0) atomicAdd(&a.stats->negatives, (unsigned long long)mm[0]);
          if (!(mm[1] == mm[1]) || !(res_rel == res_rel)) atomicAdd(&a.stats->n_bad, 1);
          const double mn = -(mm[1] - 1.0e300);
          unsigned long long* addr = reinterpret_cast<unsigned long long*>(&a.stats->fmin);
          unsigned long long old = *addr;
          while (mn < __longlong_as_double((long long)old)) {
            const unsigned long long assumed = old;
            old = atomicCAS(addr, assumed, (unsigned long long)__double_as_longlong(mn));
            if (old == assumed) break;
          }
        }
      }
      __threadfence();
      cl_sync();   // f and yprev of both halves complete before the next step's assembly reads neighbours
      steps_ok = step + 1;
    }
    __threadfence();
    cl_sync();   // (the failure path leaves the step loop without the barrier above); the problem's state is complete
    if (xa.hout && (state >= 2 || step_end == a.nsteps)) {
      double2* dst = reinterpret_cast<double2*>(xa.hout + base);
      for (int n = rank * NT + tid; n < N / 2; n += 2 * NT) dst[n] = reinterpret_cast<const double2*>(f)[n];
    }
    if (rank == 0 && tid == 0) {
      Scal* sc = a.scal + prob;
      const int cost = (a.cost ? a.cost[prob] : 0) + it_total;
      if (a.cost) a.cost[prob] = cost;
      sc->it = it;
      sc->state = state;
      sc->rmax = rmax;
      if (state >= 2 || step_end == a.nsteps) atomicMax(&a.stats->it_max, it);
      atomicAdd(&a.stats->it_sum_all, (unsigned long long)it_total);
      xa.steps_done[prob] = steps_ok;
      if (state >= 2) {
        atomicAdd(&a.stats->n_bad, 1);
        atomicMin(&a.stats->steps_min, steps_ok);
        atomicMax(&a.stats->it_total_max, cost);
        const int item_no = step_begin / xa.chunk;
        atomicSub(&xa.q->total, xa.nchunks - 1 - item_no);
      } else if (step_end < a.nsteps) {
        const int u = atomicAdd(&xa.q->tail, 1);
        st_release_gpu(xa.slots + u, prob);
      } else {
        atomicMin(&a.stats->steps_min, steps_ok);
        atomicMax(&a.stats->it_total_max, cost);
        atomicMax(reinterpret_cast<unsigned long long*>(&a.stats->resid_max), (unsigned long long)__double_as_longlong(res_true));
        atomicMax(reinterpret_cast<unsigned long long*>(&a.stats->resid_rel_max), (unsigned long long)__double_as_longlong(res_rel));
      }
    }
  }
  cl_sync();   // no CTA of the pair exits while the partner may still address its shared memory
}

}  // namespace sy2d
