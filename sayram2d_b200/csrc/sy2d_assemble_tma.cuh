// PPFV assembly with TMA-staged halo tiles (sm_100a): the fast path of engine 1.
//
// Same tile decomposition and the same per-vertex / per-face / per-row arithmetic as k_assemble_tiled
// (bit-identical rows), but the staging is done by the TMA engine instead of ~25 % of the kernel's
// instructions: one thread issues seven cp.async.bulk.tensor loads per tile - the 10 x 36 halo boxes of
// f, yprev, tx, ty, cxy and the 8 x 32 interior boxes of U, Ud - into a two-stage shared-memory ring, and
// the loads of tile k+1 are in flight while the CTA computes tile k (mbarrier complete_tx signalling).
// Out-of-domain halo cells are zero-filled by the TMA unit; they only ever meet zero weights, exactly
// like the clamped reads of the other kernels.
//
// Row-slab contexts use the same kernel on their LOCAL arrays (one halo row on each side, filled by the halo
// exchange): memory and TMA coordinates are local rows, geometry and boundary logic use global rows i = gi0 + li,
// and only the owned rows [li_begin, li_end) are written.
//
// Requirements: ny even (global strides of a tensor map are multiples of 16 bytes).  Algorithmic HBM bytes: 104 per
// cell (112 with the multigrid row weights), halo re-reads come from L2.  Tried in round 2 and reverted: a three-stage
// ring with U and Ud loaded straight into registers (14.7 KB per stage, still four CTAs per SM): 362 -> 398 us at
// 4096^2 - the two direct loads per cell sit exposed on the long scoreboard (3.7 against 1.7 stall cycles per issue).
#pragma once
#include <cuda.h>

#include "sy2d_kernels.cuh"

namespace sy2d {

// The first element of a TMA box must sit on a 16-byte boundary of global memory (a misaligned start
// raises an illegal-instruction fault: profiles/tma_probe.cu), so the halo box starts two columns left of
// the tile (J0 - 2, even) and is 36 columns wide; column J0 - 1 + b' of the halo is box column b' + 1.
constexpr int kTmaHaloI = kTI + 2, kTmaHaloJ = kTJ + 4;
constexpr int kTmaHaloElems = kTmaHaloI * kTmaHaloJ;                  // 360 doubles = 2880 B
constexpr int kTmaHaloPad = (kTmaHaloElems * 8 + 127) / 128 * 16;     // doubles per halo buffer, 128-B multiple (368)
constexpr int kTmaInnerElems = kTI * kTJ;                             // 256 doubles = 2048 B
constexpr int kTmaStageDoubles = 5 * kTmaHaloPad + 2 * kTmaInnerElems;
constexpr unsigned kTmaStageBytes = 5u * kTmaHaloElems * 8u + 2u * kTmaInnerElems * 8u;  // bytes the TMA engine delivers per tile
constexpr int kTmaStages = 2;

struct AsmMaps {   // f, yprev, tx, ty, cxy: box (36, 10, 1); U, Ud: box (32, 8, 1); dims (ny, rows of the local array, nbatch)
  CUtensorMap m[14];   // [7 .. 13]: the same arrays with the boxes of the two-cells-per-thread kernel, (68, 10, 1) and (64, 8, 1)
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "SY2D_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra SY2D_DONE;\n"
      "bra SY2D_WAIT;\n"
      "SY2D_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
               : "memory");
}

// grid: (CTAs per problem, nbatch); block: 256 threads; dynamic shared memory: kTmaSmemBytes.
struct TmaSmem {
  double stage[kTmaStages][kTmaStageDoubles];
  double vs[kTI + 1][kTJ + 1];
  double WK[kTI + 1][kTJ], WL[kTI + 1][kTJ], SK[kTI][kTJ + 1], SL[kTI][kTJ + 1];
  double red[3 * 32];
  unsigned long long full[kTmaStages];
};
constexpr size_t kTmaSmemBytes = sizeof(TmaSmem);

// The seven tensor maps live in global memory (written once by the host at context creation).
// Rows: li = local row of the arrays (tensor-map coordinate), i = gi0 + li the global row; the kernel assembles local
// rows [li_begin, li_end).  Single-GPU: gi0 = 0, li_begin = 0, li_end = nx, defer = 0.
__global__ void __launch_bounds__(kTI * kTJ, 4) k_assemble_tma(const AsmMaps* __restrict__ maps_ptr, Geometry g, AssembleOut o, int tiles_j, int ntiles,
                                                               int gi0, int li_begin, int li_end, int defer) {
  const AsmMaps& maps = *maps_ptr;
  extern __shared__ __align__(128) unsigned char tma_raw[];   // the only shared memory of the kernel: starts at a 128-B boundary
  TmaSmem& sm = *reinterpret_cast<TmaSmem*>(tma_raw);
  const int nx = g.nx, ny = g.ny;
  const int tid = threadIdx.x;
  const int a = tid >> 5, b = tid & 31;   // the thread's cell inside the tile
  const size_t base = (size_t)blockIdx.y * ((size_t)o.local_rows * ny);
  if (tid == 0) {
    for (int s = 0; s < kTmaStages; ++s) mbar_init(&sm.full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int tile_i, int tile_j, int s) {   // one thread: arm the barrier, launch the seven box loads of a tile
    const int L0 = li_begin + tile_i * kTI, J0 = tile_j * kTJ;   // local row / column of the tile origin
    double* d = sm.stage[s];
    mbar_expect_tx(&sm.full[s], kTmaStageBytes);
#pragma unroll
    for (int k = 0; k < 5; ++k) tma_load_3d(d + k * kTmaHaloPad, &maps.m[k], J0 - 2, L0 - 1, (int)blockIdx.y, &sm.full[s]);
    tma_load_3d(d + 5 * kTmaHaloPad, &maps.m[5], J0, L0, (int)blockIdx.y, &sm.full[s]);
    tma_load_3d(d + 5 * kTmaHaloPad + kTmaInnerElems, &maps.m[6], J0, L0, (int)blockIdx.y, &sm.full[s]);
  };
  // tile coordinates advance by the grid stride without a division per tile (28 instructions per warp and tile, ncu)
  const int step_i = (int)gridDim.x / tiles_j, step_j = (int)gridDim.x - step_i * tiles_j;
  int tile_i = (int)blockIdx.x / tiles_j, tile_j = (int)blockIdx.x - tile_i * tiles_j;
  if (tid == 0 && (int)blockIdx.x < ntiles) issue(tile_i, tile_j, 0);
  double rr = 0.0, rabs = 0.0;
  int it = 0;
  // vertex weights of the thread's own vertex: fetched one tile ahead (at the end of the tile before), so that their
  // latency is never exposed in the vertex stage (the long-scoreboard stall of that stage in ncu); indices clamped into the tables
  double wl = g.wxL[min(gi0 + li_begin + tile_i * kTI + a, nx)], wr = g.wxR[min(gi0 + li_begin + tile_i * kTI + a, nx)];
  double wb = g.wyB[min(tile_j * kTJ + b, ny)], wt = g.wyT[min(tile_j * kTJ + b, ny)];
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int s = it & 1;
    const unsigned parity = (unsigned)(it >> 1) & 1u;
    int next_i = tile_i + step_i, next_j = tile_j + step_j;
    if (next_j >= tiles_j) { next_j -= tiles_j; ++next_i; }
    // the other stage was fully consumed in the previous iteration (trailing __syncthreads)
    if (tid == 0 && tile + (int)gridDim.x < ntiles) issue(next_i, next_j, s ^ 1);
    const int L0 = li_begin + tile_i * kTI;          // local row of the tile origin
    const int I0 = gi0 + L0, J0 = tile_j * kTJ;      // global row / column of the tile origin
    const double* fs = sm.stage[s];                  // [10][36] halo boxes; H(arr, a', b') <-> cell (I0 + a' - 1, J0 + b' - 1)
    const double* ys = fs + kTmaHaloPad;
    const double* txs = ys + kTmaHaloPad;
    const double* tys = txs + kTmaHaloPad;
    const double* cxs = tys + kTmaHaloPad;
    const double* Us = cxs + kTmaHaloPad;            // [8][32] interior tiles
    const double* Uds = Us + kTmaInnerElems;
#define H(arr, aa, bb) arr[(aa) * kTmaHaloJ + (bb) + 1]
    const bool edge_tile = I0 == 0 || I0 + kTI >= nx || J0 == 0 || J0 + kTJ >= ny || L0 + kTI > li_end;
    mbar_wait(&sm.full[s], parity);
#ifdef SY2D_ASM_NULL
    {   // structure-only experiment (build with -DSY2D_ASM_NULL; never the product): the staged boxes straight to the outputs, no
        // vertex / face / row arithmetic - what the TMA ring + stores cost by themselves: 21.3 / 88 / 343 us at 1024^2 / 2048^2 /
        // 4096^2 against 33.6 / 100.7 / 348-365 us with the arithmetic
      const int i = I0 + a, j = J0 + b;
      if (!edge_tile || (i < nx && L0 + a < li_end && j < ny)) {
        const size_t c0 = base + (size_t)(L0 + a) * ny + j;
        const double f00 = H(fs, a + 1, b + 1), y0 = H(ys, a + 1, b + 1);
        o.wW[c0] = H(txs, a + 1, b + 1) + H(fs, a, b + 1); o.wE[c0] = H(tys, a + 1, b + 1) + H(fs, a + 2, b + 1);
        o.wS[c0] = H(cxs, a + 1, b + 1) + H(fs, a + 1, b); o.wN[c0] = Us[a * kTJ + b] + H(fs, a + 1, b + 2);
        o.rhs[c0] = Uds[a * kTJ + b]; o.cs[c0] = f00 * y0;
        if (o.om) o.om[c0] = y0;
        rr += f00; rabs = nmax(rabs, fabs(y0));
      }
      tile_i = next_i; tile_j = next_j;
      __syncthreads();
      continue;
    }
#endif
    // 2. vertices (I0 + a', J0 + b').  The extra row / column of every stage goes to a different warp
    // (1 .. 4), so that no warp does more than one extra pass between two barriers.
    auto vertex = [&](int va, int vb, bool own) {
      const int vi = I0 + va, vj = J0 + vb;
      double v = 0.0;
      if (!edge_tile) {
        double xl = wl, xr = wr, yb = wb, yt = wt;
        if (!own) { xl = g.wxL[vi]; xr = g.wxR[vi]; yb = g.wyB[vj]; yt = g.wyT[vj]; }
        v = xl * yb * H(fs, va, vb) + xr * yb * H(fs, va + 1, vb) + xl * yt * H(fs, va, vb + 1) + xr * yt * H(fs, va + 1, vb + 1);
      } else if (vi <= nx && vj <= ny) {
        v = vertex_value(g, vi, vj, H(fs, va, vb), H(fs, va + 1, vb), H(fs, va, vb + 1), H(fs, va + 1, vb + 1));
      }
      sm.vs[va][vb] = v;
    };
    vertex(a, b, true);
    if (a == 1) vertex(kTI, b, false);
    if (a == 2 && b <= kTI) vertex(b, kTJ, false);
    __syncthreads();
    // 3a. west faces of cells (I0 + a', J0 + b'), a' = 0..TI: K = (i, j), L = (i-1, j)
    auto wface = [&](int fa, int fb) {
      const int i = I0 + fa, j = J0 + fb;
      double AK = 0.0, AL = 0.0;
      if (!edge_tile || (i >= 1 && i <= nx - 1 && j < ny)) {
        const double tK = H(txs, fa + 1, fb + 1), cK = H(cxs, fa + 1, fb + 1), tL = H(txs, fa, fb + 1), cL = H(cxs, fa, fb + 1);
        const double vSW = sm.vs[fa][fb], vNW = sm.vs[fa][fb + 1];
        const double kA = tK - cK, kB = tK + cK;   // W face of K: A = NW, B = SW
        const double lA = tL - cL, lB = tL + cL;   // E face of L: A = SE_L = SW_K, B = NE_L = NW_K
        face_pair(kA * vNW + kB * vSW, kA + kB, H(fs, fa + 1, fb + 1), lA * vSW + lB * vNW, lA + lB, H(fs, fa, fb + 1), AK, AL);
      }
      sm.WK[fa][fb] = AK;
      sm.WL[fa][fb] = AL;
    };
    wface(a, b);
    if (a == 3) wface(kTI, b);
    // 3b. south faces of cells (I0 + a', J0 + b'), b' = 0..TJ: K = (i, j), L = (i, j-1)
    auto sface = [&](int fa, int fb) {
      const int i = I0 + fa, j = J0 + fb;
      double AK = 0.0, AL = 0.0;
      if (!edge_tile || (j >= 1 && j <= ny - 1 && i < nx)) {
        const double tK = H(tys, fa + 1, fb + 1), cK = H(cxs, fa + 1, fb + 1), tL = H(tys, fa + 1, fb), cL = H(cxs, fa + 1, fb);
        const double vSW = sm.vs[fa][fb], vSE = sm.vs[fa + 1][fb];
        const double kA = tK + cK, kB = tK - cK;   // S face of K: A = SW, B = SE
        const double lA = tL + cL, lB = tL - cL;   // N face of L: A = NE_L = SE_K, B = NW_L = SW_K
        face_pair(kA * vSW + kB * vSE, kA + kB, H(fs, fa + 1, fb + 1), lA * vSE + lB * vSW, lA + lB, H(fs, fa + 1, fb), AK, AL);
      }
      sm.SK[fa][fb] = AK;
      sm.SL[fa][fb] = AL;
    };
    sface(a, b);
    if (a == 4 && b < kTI) sface(b, kTJ);
    __syncthreads();
    // 4. rows
    const int i = I0 + a, j = J0 + b;
    if (!edge_tile || (i < nx && L0 + a < li_end && j < ny)) {
      const size_t c0 = base + (size_t)(L0 + a) * ny + j;
      const double f00 = H(fs, a + 1, b + 1);
      double diag = 0.0, R = 0.0, oW = 0.0, oE = 0.0, oS = 0.0, oN = 0.0;
      if (!edge_tile) {
        diag = sm.WK[a][b]; oW = -sm.WL[a][b];
        diag += sm.WL[a + 1][b]; oE = -sm.WK[a + 1][b];
        diag += sm.SK[a][b]; oS = -sm.SL[a][b];
        diag += sm.SL[a][b + 1]; oN = -sm.SK[a][b + 1];
      } else {
        if (i > 0) { diag += sm.WK[a][b]; oW = -sm.WL[a][b]; }
        if (i < nx - 1) { diag += sm.WL[a + 1][b]; oE = -sm.WK[a + 1][b]; }
        if (j > 0) { diag += sm.SK[a][b]; oS = -sm.SL[a][b]; }
        if (j < ny - 1) { diag += sm.SL[a][b + 1]; oN = -sm.SK[a][b + 1]; }
        if (i == 0 || i == nx - 1 || j == 0 || j == ny - 1) {  // Dirichlet boundary faces (Solver.cc:143-164, 204-267)
          const double txP = H(txs, a + 1, b + 1), tyP = H(tys, a + 1, b + 1), cP = H(cxs, a + 1, b + 1);
          const double vSW = sm.vs[a][b], vSE = sm.vs[a + 1][b], vNW = sm.vs[a][b + 1], vNE = sm.vs[a + 1][b + 1];
          if (i == 0 && g.bc[0] == 0) diag += dirichlet_face((txP - cP) * vNW + (txP + cP) * vSW, (txP - cP) + (txP + cP), f00, R);
          if (i == nx - 1 && g.bc[1] == 0) diag += dirichlet_face((txP - cP) * vSE + (txP + cP) * vNE, (txP - cP) + (txP + cP), f00, R);
          if (j == 0 && g.bc[2] == 0) diag += dirichlet_face((tyP + cP) * vSW + (tyP - cP) * vSE, (tyP + cP) + (tyP - cP), f00, R);
          if (j == ny - 1 && g.bc[3] == 0) diag += dirichlet_face((tyP + cP) * vNE + (tyP - cP) * vNW, (tyP + cP) + (tyP - cP), f00, R);
        }
      }
      diag += Uds[a * kTJ + b];
      R += Us[a * kTJ + b] * f00;
      const double cs0 = f00 * H(ys, a + 1, b + 1);
      const double om = diag * cs0;
      const double dscale = sy2d_div(1.0, om);
      const double wW = oW * (H(fs, a, b + 1) * H(ys, a, b + 1)) * dscale, wE = oE * (H(fs, a + 2, b + 1) * H(ys, a + 2, b + 1)) * dscale;
      const double wS = oS * (H(fs, a + 1, b) * H(ys, a + 1, b)) * dscale, wN = oN * (H(fs, a + 1, b + 2) * H(ys, a + 1, b + 2)) * dscale;
      const double rhs = R * dscale - 1.0 - ((wW + wE) + (wS + wN));
      o.wW[c0] = wW; o.wE[c0] = wE; o.wS[c0] = wS; o.wN[c0] = wN;
      o.rhs[c0] = rhs;
      o.cs[c0] = cs0;
      if (o.om) o.om[c0] = om;
      rr += rhs * rhs;
      rabs = nmax(rabs, fabs(rhs));
    }
#undef H
    tile_i = next_i; tile_j = next_j;
    if (tile + (int)gridDim.x < ntiles) {   // weights of the next tile's vertex
      const int vi = min(gi0 + li_begin + tile_i * kTI + a, nx), vj = min(tile_j * kTJ + b, ny);
      wl = g.wxL[vi]; wr = g.wxR[vi]; wb = g.wyB[vj]; wt = g.wyT[vj];
    }
    __syncthreads();  // vs / face arrays and this stage are rewritten from here on
  }
  double sums[1] = {rr};
  block_sums<1>(sums, sm.red);
  const double bmax = block_max(rabs, sm.red);
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
