// PPFV assembly, TMA-staged tiles walked UP a column strip (sm_100a): the default fast path of engine 1 since round 2.
//
// Same 8 x 32 tiles, the same TMA boxes and bit-identical per-vertex / per-face / per-row arithmetic as k_assemble_tma
// (sy2d_assemble_tma.cuh), restructured around what ncu showed that kernel waits for (profiles/r02_ncu_full.md, capture
// a1024: 449 warp instructions per warp and tile, barrier stalls 3.1 cycles per issue):
//   * a CTA works through a RUN of consecutive tiles of one column strip, bottom to top.  The vertex row and the west-face
//     row on top of a tile are the bottom rows of the next one: they stay in shared memory (ring of 9 rows), so a warp
//     computes exactly one vertex row and one west-face row per tile and the extra pass of one warp per stage (the warp
//     that made the others wait at the barrier) is gone; the tile index arithmetic (two integer divisions per warp and tile)
//     is gone with the strided tile order;
//   * the column-32 leftovers (8 south faces, 9 vertices per tile) and the bottom rows at the start of a run belong to a
//     ninth HELPER warp, so the eight worker warps do the same work between two barriers;
//   * two barriers per tile instead of three: the next tile's loads are issued after the first barrier of a tile (every
//     thread is then done with the rows of the tile before), the face arrays are rewritten after that same barrier;
//   * vertex weights are fetched one tile ahead (they were the long-scoreboard stall of the vertex stage);
//   * runs are cut on the host so that every CTA gets the same COST, boundary tiles (predicated path) weighing more.
#pragma once
#include "sy2d_assemble_tma.cuh"

namespace sy2d {

constexpr int kColThreads = (kTI + 1) * 32;   // eight worker warps (warp a = row a of the tile) + the helper warp

struct ColSmem {
  double stage[kTmaStages][kTmaStageDoubles];
  double vs[kTI + 1][kTJ + 2];                                   // ring of vertex rows (physical row = (r + base) mod 9)
  double WK[kTI + 1][kTJ], WL[kTI + 1][kTJ];                     // ring of west-face rows
  double SK[kTI][kTJ + 1], SL[kTI][kTJ + 1];
  double red[3 * 32];
  unsigned long long full[kTmaStages];
};
constexpr size_t kColSmemBytes = sizeof(ColSmem);

// run_start[c] .. run_start[c + 1]: the tiles of CTA c in strip-major order t = tile_j * tiles_i + tile_i.
// grid: (CTAs per problem, nbatch); block: kColThreads; dynamic shared memory: kColSmemBytes.
__global__ void __launch_bounds__(kColThreads, 4) k_assemble_col(const AsmMaps* __restrict__ maps_ptr, Geometry g, AssembleOut o,
                                                                  const int* __restrict__ run_start, int tiles_i, int gi0, int li_begin,
                                                                  int li_end, int defer) {
  const AsmMaps& maps = *maps_ptr;
  extern __shared__ __align__(128) unsigned char col_raw[];
  ColSmem& sm = *reinterpret_cast<ColSmem*>(col_raw);
  const int nx = g.nx, ny = g.ny;
  const int tid = threadIdx.x;
  const int a = tid >> 5, b = tid & 31;   // worker: the thread's cell inside the tile; a == kTI: helper warp
  const bool helper = a == kTI;
  const size_t base = (size_t)blockIdx.y * ((size_t)o.local_rows * ny);
  const int t_begin = run_start[blockIdx.x], t_end = run_start[blockIdx.x + 1];
  if (tid == 0) {
    for (int s = 0; s < kTmaStages; ++s) mbar_init(&sm.full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  int tile_j = t_begin / tiles_i, tile_i = t_begin - tile_j * tiles_i;
  auto issue = [&](int ti, int tj, int s) {   // one thread: arm the barrier, launch the seven box loads of a tile
    const int L0 = li_begin + ti * kTI, J0 = tj * kTJ;
    double* d = sm.stage[s];
    mbar_expect_tx(&sm.full[s], kTmaStageBytes);
#pragma unroll
    for (int k = 0; k < 5; ++k) tma_load_3d(d + k * kTmaHaloPad, &maps.m[k], J0 - 2, L0 - 1, (int)blockIdx.y, &sm.full[s]);
    tma_load_3d(d + 5 * kTmaHaloPad, &maps.m[5], J0, L0, (int)blockIdx.y, &sm.full[s]);
    tma_load_3d(d + 5 * kTmaHaloPad + kTmaInnerElems, &maps.m[6], J0, L0, (int)blockIdx.y, &sm.full[s]);
  };
  if (tid == 0 && t_begin < t_end) {
    issue(tile_i, tile_j, 0);
    if (t_begin + 1 < t_end) {
      const bool wrap = tile_i + 1 == tiles_i;
      issue(wrap ? 0 : tile_i + 1, wrap ? tile_j + 1 : tile_j, 1);
    }
  }
  double rr = 0.0, rabs = 0.0;
  int rbase = 0;        // physical row of logical row 0 in the rings
  bool fresh = true;    // no rows carried over: first tile of the run or of a strip
  // vertex weights of the row this thread computes (vertex row a + 1; helper lane l: l + 1), fetched one tile ahead, and of
  // its column (fixed along a strip)
  const int vrow = helper ? (b < kTI ? b + 1 : 0) : a + 1;
  double wl_n = 0.0, wr_n = 0.0, wb = 0.0, wt = 0.0;
  int wcol = -1;
  {
    const int vi = min(gi0 + li_begin + tile_i * kTI + vrow, nx);
    wl_n = g.wxL[vi]; wr_n = g.wxR[vi];
  }
  for (int t = t_begin, it = 0; t < t_end; ++t, ++it) {
    const int s = it & 1;
    const unsigned parity = (unsigned)(it >> 1) & 1u;
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
    const bool last_of_strip = tile_i + 1 == tiles_i;
    const double wl = wl_n, wr = wr_n;
    if (wcol != tile_j) {
      const int vj = min(J0 + (helper ? kTJ : b), ny);
      wb = g.wyB[vj]; wt = g.wyT[vj];
      wcol = tile_j;
    }
    {  // weights of the next tile's row
      const int ni = last_of_strip ? 0 : tile_i + 1;
      const int vi = min(gi0 + li_begin + ni * kTI + vrow, nx);
      wl_n = g.wxL[vi]; wr_n = g.wxR[vi];
    }
    // physical ring rows of the logical rows a, a + 1 (workers)
    int pa = a + rbase; if (pa >= kTI + 1) pa -= kTI + 1;
    int pa1 = pa + 1; if (pa1 >= kTI + 1) pa1 -= kTI + 1;
    auto prow = [&](int r) { int p = r + rbase; return p >= kTI + 1 ? p - (kTI + 1) : p; };
    mbar_wait(&sm.full[s], parity);
    // 1. vertices (I0 + va, J0 + vb)
    auto vertex = [&](int va, int vb, int pr, double xl, double xr, double yb, double yt, bool have_w) {
      const int vi = I0 + va, vj = J0 + vb;
      double v = 0.0;
      if (!edge_tile) {
        if (!have_w) { xl = g.wxL[vi]; xr = g.wxR[vi]; yb = g.wyB[vj]; yt = g.wyT[vj]; }
        v = xl * yb * H(fs, va, vb) + xr * yb * H(fs, va + 1, vb) + xl * yt * H(fs, va, vb + 1) + xr * yt * H(fs, va + 1, vb + 1);
      } else if (vi <= nx && vj <= ny) {
        v = vertex_value(g, vi, vj, H(fs, va, vb), H(fs, va + 1, vb), H(fs, va, vb + 1), H(fs, va + 1, vb + 1));
      }
      sm.vs[pr][vb] = v;
    };
    if (!helper) {
      vertex(a + 1, b, pa1, wl, wr, wb, wt, true);
    } else {
      if (b < kTI) vertex(b + 1, kTJ, prow(b + 1), wl, wr, wb, wt, true);
      if (fresh) {
        vertex(0, b, rbase, 0.0, 0.0, 0.0, 0.0, false);
        if (b == 0) vertex(0, kTJ, rbase, 0.0, 0.0, 0.0, 0.0, false);
      }
    }
    __syncthreads();
    // every thread is done with the tile before: its stage is free for the tile after this one
    if (tid == 0 && it >= 1 && t + 1 < t_end) issue(last_of_strip ? 0 : tile_i + 1, last_of_strip ? tile_j + 1 : tile_j, s ^ 1);
    // 2a. west faces of cells (I0 + fa, J0 + fb), fa = 0..TI: K = (i, j), L = (i-1, j)
    auto wface = [&](int fa, int fb, int pr) {
      const int i = I0 + fa, j = J0 + fb;
      double AK = 0.0, AL = 0.0;
      if (!edge_tile || (i >= 1 && i <= nx - 1 && j < ny)) {
        const double tK = H(txs, fa + 1, fb + 1), cK = H(cxs, fa + 1, fb + 1), tL = H(txs, fa, fb + 1), cL = H(cxs, fa, fb + 1);
        const double vSW = sm.vs[pr][fb], vNW = sm.vs[pr][fb + 1];
        const double kA = tK - cK, kB = tK + cK;   // W face of K: A = NW, B = SW
        const double lA = tL - cL, lB = tL + cL;   // E face of L: A = SE_L = SW_K, B = NE_L = NW_K
        face_pair(kA * vNW + kB * vSW, kA + kB, H(fs, fa + 1, fb + 1), lA * vSW + lB * vNW, lA + lB, H(fs, fa, fb + 1), AK, AL);
      }
      sm.WK[pr][fb] = AK;
      sm.WL[pr][fb] = AL;
    };
    // 2b. south faces of cells (I0 + fa, J0 + fb), fb = 0..TJ: K = (i, j), L = (i, j-1); p0 / p1: ring rows of vertex rows fa, fa + 1
    auto sface = [&](int fa, int fb, int p0, int p1) {
      const int i = I0 + fa, j = J0 + fb;
      double AK = 0.0, AL = 0.0;
      if (!edge_tile || (j >= 1 && j <= ny - 1 && i < nx)) {
        const double tK = H(tys, fa + 1, fb + 1), cK = H(cxs, fa + 1, fb + 1), tL = H(tys, fa + 1, fb), cL = H(cxs, fa + 1, fb);
        const double vSW = sm.vs[p0][fb], vSE = sm.vs[p1][fb];
        const double kA = tK + cK, kB = tK - cK;   // S face of K: A = SW, B = SE
        const double lA = tL + cL, lB = tL - cL;   // N face of L: A = NE_L = SE_K, B = NW_L = SW_K
        face_pair(kA * vSW + kB * vSE, kA + kB, H(fs, fa + 1, fb + 1), lA * vSE + lB * vSW, lA + lB, H(fs, fa + 1, fb), AK, AL);
      }
      sm.SK[fa][fb] = AK;
      sm.SL[fa][fb] = AL;
    };
    if (!helper) {
      wface(a + 1, b, pa1);
      sface(a, b, pa, pa1);
    } else {
      if (b < kTI) { const int p0 = prow(b); sface(b, kTJ, p0, p0 + 1 >= kTI + 1 ? 0 : p0 + 1); }
      if (fresh) wface(0, b, rbase);
    }
    __syncthreads();
    // 3. rows
    if (!helper) {
      const int i = I0 + a, j = J0 + b;
      if (!edge_tile || (i < nx && L0 + a < li_end && j < ny)) {
        const size_t c0 = base + (size_t)(L0 + a) * ny + j;
        const double f00 = H(fs, a + 1, b + 1);
        double diag = 0.0, R = 0.0, oW = 0.0, oE = 0.0, oS = 0.0, oN = 0.0;
        if (!edge_tile) {
          diag = sm.WK[pa][b]; oW = -sm.WL[pa][b];
          diag += sm.WL[pa1][b]; oE = -sm.WK[pa1][b];
          diag += sm.SK[a][b]; oS = -sm.SL[a][b];
          diag += sm.SL[a][b + 1]; oN = -sm.SK[a][b + 1];
        } else {
          if (i > 0) { diag += sm.WK[pa][b]; oW = -sm.WL[pa][b]; }
          if (i < nx - 1) { diag += sm.WL[pa1][b]; oE = -sm.WK[pa1][b]; }
          if (j > 0) { diag += sm.SK[a][b]; oS = -sm.SL[a][b]; }
          if (j < ny - 1) { diag += sm.SL[a][b + 1]; oN = -sm.SK[a][b + 1]; }
          if (i == 0 || i == nx - 1 || j == 0 || j == ny - 1) {  // Dirichlet boundary faces (Solver.cc:143-164, 204-267)
            const double txP = H(txs, a + 1, b + 1), tyP = H(tys, a + 1, b + 1), cP = H(cxs, a + 1, b + 1);
            const double vSW = sm.vs[pa][b], vSE = sm.vs[pa1][b], vNW = sm.vs[pa][b + 1], vNE = sm.vs[pa1][b + 1];
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
    }
#undef H
    if (edge_tile) __syncthreads();   // boundary rows read the vertex ring, which the next tile's first stage rewrites
    // the top rows of this tile are the bottom rows of the next one
    rbase += kTI; if (rbase >= kTI + 1) rbase -= kTI + 1;
    fresh = last_of_strip;
    if (last_of_strip) { tile_i = 0; ++tile_j; } else { ++tile_i; }
  }
  __syncthreads();
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
