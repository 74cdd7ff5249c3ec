// PPFV assembly, warp-marching variant (sm_100a): no shared memory, no CTA barrier.
//
// A WARP owns a strip of kMarchCols = 30 columns x kMarchRows rows.  Lane L holds column J0 - 1 + L (lanes 0 and 31
// are the halo columns of the strip) and marches up the rows: one coalesced 256-byte row segment per array and row,
// loaded one row ahead of its use, everything else in registers.
//   * the E face of row i is the W face of row i + 1: computed once, carried in registers (engine 2's march);
//   * the S face of a cell is computed by its own lane, the N face is the S face of the lane to the right and comes
//     over with two shuffles; so do the vertex value of the right-hand vertex column and the neighbours' column scale;
//   * per cell: 1 vertex, 1 W face, 1 S face, 1 row (the tile kernels: 1.1, 1.1 + 1.1, 1 with ~25 shared-memory
//     accesses per cell and three CTA barriers per tile) - about half the instructions of k_assemble_tma, and a warp
//     never waits for another warp.
// Lanes 0 and 31 do the same arithmetic without owning a cell (30 / 32 of the lanes store); a strip re-reads the row
// below and the row above it (18 rows read per 16 assembled: from L2).  Same per-face / per-row expressions and the
// same accumulation order as k_assemble_tiled; the compiler contracts them into FMAs differently in this kernel body,
// so the rows agree with the tile kernels' to the last bits (1e-14 relative, tests/test_gpu_parity.py), not bit for bit.
//
// Rows: memory is indexed with LOCAL rows li (a slab's arrays carry one halo row per side), geometry and boundary
// logic with GLOBAL rows i = gi0 + li; the kernel assembles local rows [li_begin, li_end).  Single GPU: gi0 = 0,
// li_begin = 0, li_end = nx.  Algorithmic HBM bytes: 104 per cell (112 with the multigrid row weights).
#pragma once
#include "sy2d_kernels.cuh"

namespace sy2d {

constexpr int kMarchCols = 30;      // owned columns per warp (lanes 1..30)
constexpr int kMarchRows = 16;      // rows per strip
constexpr int kMarchWarps = 4;      // warps per CTA (independent of each other)
constexpr int kMarchCtasPerSm = 5;
// Rows are prefetched with cp.async (LDGSTS: global -> shared memory without a destination register) into a per-lane
// ring of kMarchRing rows x 7 arrays: a lane reads back only the slots it filled itself, so shared memory serves as
// asynchronously filled extra registers and no warp or CTA synchronisation is involved.  kMarchAhead rows are in flight
// per lane (16 warps x 3 rows x 1.8 KB = 86 KB per SM): with the loads in registers one row ahead (29 KB per SM in flight)
// the kernel sat at 37 us for 1024^2 / 4.1 TB/s at 4096^2, waiting for memory latency once per row.
constexpr int kMarchAhead = 3;                  // rows in flight per lane
constexpr int kMarchRing = kMarchAhead + 1;     // the slot refilled in iteration vl was read in iteration vl - 1
constexpr int kMarchSlot = 256;                 // doubles per ring slot: 7 arrays x 32 lanes, padded to a power of two
constexpr size_t kMarchSmemBytes = (size_t)kMarchWarps * kMarchRing * kMarchSlot * sizeof(double);   // 32 KB per CTA
static_assert((kMarchRing & (kMarchRing - 1)) == 0, "the ring index wraps with a mask");

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct MarchRow {   // what a lane holds of one row of its column
  double f, yp, tx, cxy, ty, U, Ud;
};

struct MarchPtrs {   // per problem (batch offset applied)
  const double *f, *yp, *tx, *ty, *cxy, *U, *Ud;
  double *wW, *wE, *wS, *wN, *rhs, *cs, *om;
};

// One strip.  EDGE = false: the strip and its halo rows / columns lie strictly inside the domain - no clamp, no
// boundary branch, no predicate but `own` in the row loop (the instruction stream of the loop is what bounds this
// kernel: running offsets instead of index products, a power-of-two ring).
template <bool EDGE>
__device__ __forceinline__ void march_strip(const MarchPtrs& P, const Geometry& g, double* ring, int lane, int gi0, int L0, int L1, int J0,
                                            double& rr, double& rabs) {
  const unsigned full = 0xffffffffu;
  const int nx = g.nx, ny = g.ny;
  const int jraw = J0 - 1 + lane;                            // the lane's column (may be -1 or >= ny: halo / padding lanes)
  const int j = EDGE ? (jraw < 0 ? 0 : (jraw > ny - 1 ? ny - 1 : jraw)) : jraw;   // clamped for loads (clamped values only meet zero weights)
  const bool own = lane >= 1 && lane <= kMarchCols && (!EDGE || jraw < ny);
  const int jv = EDGE ? (jraw < 0 ? 0 : (jraw > ny ? ny : jraw)) : jraw;
  const double wb = g.wyB[jv], wt = g.wyT[jv];               // vertex column jraw
  // Prefetch cursor: row li_p (local), offset n_p of (row, column j) in the problem's arrays, ring slot s_p (doubles).
  // Rows outside the domain (global row -1 under the first strip, nx above the last one) are clamped to the nearest row.
  int li_p = L0 - 1;
  int n_p;
  {
    int i = gi0 + li_p;
    if (EDGE) i = i < 0 ? 0 : i;
    n_p = (i - gi0) * ny + j;
  }
  int s_p = 0;
  auto prefetch = [&]() {   // row li_p -> slot s_p (one commit group per row, empty past the strip), then advance the cursor
    if (li_p <= L1) {
      double* d = ring + s_p;
      cp_async8(d, P.f + n_p); cp_async8(d + 32, P.yp + n_p); cp_async8(d + 64, P.tx + n_p); cp_async8(d + 96, P.cxy + n_p);
      if (li_p < L1) { cp_async8(d + 128, P.ty + n_p); cp_async8(d + 160, P.U + n_p); cp_async8(d + 192, P.Ud + n_p); }
    }
    cp_async_commit();
    const int gi = gi0 + li_p;           // the next row exists in the domain unless this one is the last (or the clamped row -1)
    n_p += (!EDGE || (gi >= 0 && gi < nx - 1)) ? ny : 0;
    li_p += 1;
    s_p = (s_p + kMarchSlot) & (kMarchRing * kMarchSlot - 1);
  };
  int s_f = 0;
  auto fetch = [&](MarchRow& r) {   // the lane's own slots of the next row of the march (complete: see the wait before the call)
    const double* d = ring + s_f;
    r.f = d[0]; r.yp = d[32]; r.tx = d[64]; r.cxy = d[96]; r.ty = d[128]; r.U = d[160]; r.Ud = d[192];
    s_f = (s_f + kMarchSlot) & (kMarchRing * kMarchSlot - 1);
  };
  MarchRow below, cur;                   // rows vl - 1 and vl of the march, in registers
#pragma unroll
  for (int d = 0; d < kMarchAhead; ++d) prefetch();
  cp_async_wait<kMarchAhead - 1>();      // row L0 - 1 has landed
  fetch(cur);
  double cs_below = 0.0;                 // column scale of the row under `below`
  double vL_prev = 0.0, vR_prev = 0.0;   // vertex row vl - 1: V(vl-1, j), V(vl-1, j+1)
  double AKw = 0.0, ALw = 0.0;           // W face of row vl - 1 (K = row vl-1, L = row vl-2)
  const double* pwl = g.wxL + (gi0 + L0);
  const double* pwr = g.wxR + (gi0 + L0);
  int c0 = (L0 - 1) * ny + jraw;         // store offset of row vl - 1
  // march over the vertex rows vl = L0 .. L1 (local); iteration vl finalises cell row vl - 1
  for (int vl = L0; vl <= L1; ++vl, c0 += ny) {
    below = cur;
    prefetch();                          // refills the slot of row vl - 2, read two iterations ago
    cp_async_wait<kMarchAhead - 1>();    // row vl has landed (rows vl + 1 .. vl + kMarchAhead - 1 may still be in flight)
    fetch(cur);                          // cur = row vl (the upper row of this vertex row), below = row vl - 1
    const int vi = gi0 + vl;             // global vertex row
    // vertices V(vi, jraw) and V(vi, jraw + 1)
    const double fL_below = __shfl_up_sync(full, below.f, 1), fL_cur = __shfl_up_sync(full, cur.f, 1);   // column jraw - 1
    double vL;
    if (!EDGE) {
      const double wl = *pwl++, wr = *pwr++;
      vL = wl * wb * fL_below + wr * wb * fL_cur + wl * wt * below.f + wr * wt * cur.f;
    } else {
      vL = (jraw >= 0 && jraw <= ny && vi <= nx) ? vertex_value(g, vi, jraw, fL_below, fL_cur, below.f, cur.f) : 0.0;
    }
    const double vR = __shfl_down_sync(full, vL, 1);
    // W face between row vi (K) and row vi - 1 (L)
    double AKn = 0.0, ALn = 0.0;
    if (!EDGE || (vi >= 1 && vi <= nx - 1)) {
      const double kA = cur.tx - cur.cxy, kB = cur.tx + cur.cxy;        // W face of K: A = NW, B = SW
      const double lA = below.tx - below.cxy, lB = below.tx + below.cxy;  // E face of L: A = SE_L = SW_K, B = NE_L = NW_K
      face_pair(kA * vR + kB * vL, kA + kB, cur.f, lA * vL + lB * vR, lA + lB, below.f, AKn, ALn);
    }
    const double cs0 = below.f * below.yp;
    if (vl > L0) {
      // ---- finalise cell row r = vl - 1 (global i): `below` ----
      const int i = vi - 1;
      const double f00 = below.f;
      // S face of (i, jraw): K = own cell, L = (i, jraw - 1); vertices SW = V(i, jraw) = vL_prev, SE = V(i+1, jraw) = vL
      const double tyL = __shfl_up_sync(full, below.ty, 1), cL = __shfl_up_sync(full, below.cxy, 1);
      double SKo = 0.0, SLo = 0.0;
      if (!EDGE || (jraw >= 1 && jraw <= ny - 1)) {
        const double kA = below.ty + below.cxy, kB = below.ty - below.cxy;   // S face of K: A = SW, B = SE
        const double lA = tyL + cL, lB = tyL - cL;                           // N face of L: A = NE_L = SE_K, B = NW_L = SW_K
        face_pair(kA * vL_prev + kB * vL, kA + kB, f00, lA * vL + lB * vL_prev, lA + lB, fL_below, SKo, SLo);
      }
      const double SKn = __shfl_down_sync(full, SKo, 1), SLn = __shfl_down_sync(full, SLo, 1);   // S face of the cell to the right = my N face
      const double csS = __shfl_up_sync(full, cs0, 1), csN = __shfl_down_sync(full, cs0, 1);
      if (own) {
        double diag = 0.0, R = 0.0, oW = 0.0, oE = 0.0, oS = 0.0, oN = 0.0;
        if (!EDGE) {
          diag = AKw; oW = -ALw;
          diag += ALn; oE = -AKn;
          diag += SKo; oS = -SLo;
          diag += SLn; oN = -SKn;
        } else {
          if (i > 0) { diag += AKw; oW = -ALw; }
          if (i < nx - 1) { diag += ALn; oE = -AKn; }
          if (jraw > 0) { diag += SKo; oS = -SLo; }
          if (jraw < ny - 1) { diag += SLn; oN = -SKn; }
          if (i == 0 || i == nx - 1 || jraw == 0 || jraw == ny - 1) {  // Dirichlet boundary faces (Solver.cc:143-164, 204-267)
            const double txP = below.tx, tyP = below.ty, cP = below.cxy;
            const double vSW = vL_prev, vSE = vL, vNW = vR_prev, vNE = vR;
            if (i == 0 && g.bc[0] == 0) diag += dirichlet_face((txP - cP) * vNW + (txP + cP) * vSW, (txP - cP) + (txP + cP), f00, R);
            if (i == nx - 1 && g.bc[1] == 0) diag += dirichlet_face((txP - cP) * vSE + (txP + cP) * vNE, (txP - cP) + (txP + cP), f00, R);
            if (jraw == 0 && g.bc[2] == 0) diag += dirichlet_face((tyP + cP) * vSW + (tyP - cP) * vSE, (tyP + cP) + (tyP - cP), f00, R);
            if (jraw == ny - 1 && g.bc[3] == 0) diag += dirichlet_face((tyP + cP) * vNE + (tyP - cP) * vNW, (tyP + cP) + (tyP - cP), f00, R);
          }
        }
        diag += below.Ud;
        R += below.U * f00;
        const double om = diag * cs0;
        const double dscale = sy2d_div(1.0, om);
        const double wW = oW * cs_below * dscale, wE = oE * (cur.f * cur.yp) * dscale;
        const double wS = oS * csS * dscale, wN = oN * csN * dscale;
        const double rhs = R * dscale - 1.0 - ((wW + wE) + (wS + wN));
        P.wW[c0] = wW; P.wE[c0] = wE; P.wS[c0] = wS; P.wN[c0] = wN;
        P.rhs[c0] = rhs;
        P.cs[c0] = cs0;
        if (P.om) P.om[c0] = om;
        rr += rhs * rhs;
        rabs = nmax(rabs, fabs(rhs));
      }
    }
    cs_below = cs0;                      // row vl - 1 becomes the W neighbour of row vl
    vL_prev = vL; vR_prev = vR;
    AKw = AKn; ALw = ALn;
  }
}

__global__ void __launch_bounds__(kMarchWarps * 32, kMarchCtasPerSm) k_assemble_march(const double* __restrict__ f, const double* __restrict__ yprev,
                                                                       const double* __restrict__ tx, const double* __restrict__ ty,
                                                                       const double* __restrict__ cxy, const double* __restrict__ U,
                                                                       const double* __restrict__ Ud, Geometry g, AssembleOut o, int strips_j,
                                                                       int nstrips, int gi0, int li_begin, int li_end, int defer) {
  __shared__ double red[3 * 32];
  extern __shared__ double march_ring[];   // [warp][slot][array][lane], slots padded to kMarchSlot doubles
  const int nx = g.nx, ny = g.ny;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t base = (size_t)blockIdx.y * ((size_t)o.local_rows * ny);
  MarchPtrs P;
  P.f = f + base; P.yp = yprev + base; P.tx = tx + base; P.ty = ty + base; P.cxy = cxy + base; P.U = U + base; P.Ud = Ud + base;
  P.wW = o.wW + base; P.wE = o.wE + base; P.wS = o.wS + base; P.wN = o.wN + base; P.rhs = o.rhs + base; P.cs = o.cs + base;
  P.om = o.om ? o.om + base : nullptr;
  double* ring = march_ring + (size_t)warp * (kMarchRing * kMarchSlot) + lane;
  double rr = 0.0, rabs = 0.0;
  const int warps_total = gridDim.x * kMarchWarps;
  for (int strip = blockIdx.x * kMarchWarps + warp; strip < nstrips; strip += warps_total) {
    const int si = strip / strips_j, sj = strip - si * strips_j;
    const int L0 = li_begin + si * kMarchRows;                 // first local row of the strip
    const int L1 = min(L0 + kMarchRows, li_end);               // one past its last row
    const int J0 = sj * kMarchCols;
    // the strip (with its halo rows L0 - 1, L1 and halo columns J0 - 1, J0 + 30) touches the domain boundary
    const bool edge = gi0 + L0 == 0 || gi0 + L1 >= nx || J0 == 0 || J0 + kMarchCols >= ny - 1;
    if (edge) march_strip<true>(P, g, ring, lane, gi0, L0, L1, J0, rr, rabs);
    else march_strip<false>(P, g, ring, lane, gi0, L0, L1, J0, rr, rabs);
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

}  // namespace sy2d
