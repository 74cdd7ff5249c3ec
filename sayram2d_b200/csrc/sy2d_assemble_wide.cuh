// PPFV assembly with TMA-staged halo tiles, TWO cells per thread (sm_100a).
//
// k_assemble_tma (one cell per thread, 8 x 32 tiles) runs at ~7 SM-cycles per cell whatever is trimmed from its instruction
// stream: ~40 eight-byte shared-memory loads and ~450 thread instructions per cell at 32 warps per SM keep the L1TEX pipe
// (58 %), the issue slots (49 %) and the FP64 pipe (29 %) all half busy and none of them full.  Here a thread owns the two cells
// (a, 2b), (a, 2b + 1) of an 8 x 64 tile: the same TMA boxes (68 x 10 halo boxes, 64 x 8 interior boxes), the same per-vertex /
// per-face / per-row expressions (bit-identical rows), but
//   * 16-byte shared-memory accesses wherever the pair is aligned (the halo box starts two columns left of the tile, so the
//     pair (2b, 2b + 1) sits on a 16-byte boundary): ~16 loads per cell instead of ~40;
//   * the south face between the thread's two cells and its own west faces never leave the registers;
//   * four independent faces (eight divisions) in flight per thread in the face stage;
//   * index arithmetic, predicates and the three barriers of a tile are shared by 512 cells instead of 256.
// 94 KB of shared memory per CTA: two CTAs per SM (128 registers per thread).
#pragma once
#include "sy2d_assemble_tma.cuh"

namespace sy2d {

constexpr int kWJ = 2 * kTJ;                                        // 64 columns per tile
constexpr int kWHaloJ = kWJ + 4;                                    // 68
constexpr int kWHaloElems = kTmaHaloI * kWHaloJ;                    // 680 doubles
constexpr int kWHaloPad = (kWHaloElems * 8 + 127) / 128 * 16;       // 688 doubles (128-byte multiple)
constexpr int kWInnerElems = kTI * kWJ;                             // 512 doubles
constexpr int kWStageDoubles = 5 * kWHaloPad + 2 * kWInnerElems;
constexpr unsigned kWStageBytes = 5u * kWHaloElems * 8u + 2u * kWInnerElems * 8u;
constexpr int kWVS = kWJ + 2;                                       // row stride of the vertex array (even: 16-byte pairs)
constexpr int kWSS = kTJ + 2;                                       // row stride of the south-face arrays (one face per pair + column 64)

struct WideSmem {
  double stage[kTmaStages][kWStageDoubles];
  double vs[kTI + 1][kWVS];
  double WK[kTI + 1][kWJ], WL[kTI + 1][kWJ];
  double SK[kTI][kWSS], SL[kTI][kWSS];     // south face of cell (a, 2b') at [a][b'], b' = 0..32
  double red[3 * 32];
  unsigned long long full[kTmaStages];
};
constexpr size_t kWideSmemBytes = sizeof(WideSmem);

__device__ __forceinline__ double2 lds2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void sts2(double* p, double x, double y) { *reinterpret_cast<double2*>(p) = make_double2(x, y); }
__device__ __forceinline__ void stg2(double* p, double x, double y) { *reinterpret_cast<double2*>(p) = make_double2(x, y); }

// maps: m[7..13] of AsmMaps (wide boxes).  grid: (CTAs per problem, nbatch); block: 256; dynamic shared memory: kWideSmemBytes.
__global__ void __launch_bounds__(kTI * kTJ, 2) k_assemble_wide(const AsmMaps* __restrict__ maps_ptr, Geometry g, AssembleOut o, int tiles_j, int ntiles,
                                                                int gi0, int li_begin, int li_end, int defer) {
  const AsmMaps& maps = *maps_ptr;
  extern __shared__ __align__(128) unsigned char wide_raw[];
  WideSmem& sm = *reinterpret_cast<WideSmem*>(wide_raw);
  const int nx = g.nx, ny = g.ny;
  const int tid = threadIdx.x;
  const int a = tid >> 5, b = tid & 31;   // the thread's cells inside the tile: (a, 2b), (a, 2b + 1)
  const int j0 = 2 * b;
  const size_t base = (size_t)blockIdx.y * ((size_t)o.local_rows * ny);
  if (tid == 0) {
    for (int s = 0; s < kTmaStages; ++s) mbar_init(&sm.full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int tile_i, int tile_j, int s) {
    const int L0 = li_begin + tile_i * kTI, J0 = tile_j * kWJ;
    double* d = sm.stage[s];
    mbar_expect_tx(&sm.full[s], kWStageBytes);
#pragma unroll
    for (int k = 0; k < 5; ++k) tma_load_3d(d + k * kWHaloPad, &maps.m[7 + k], J0 - 2, L0 - 1, (int)blockIdx.y, &sm.full[s]);
    tma_load_3d(d + 5 * kWHaloPad, &maps.m[12], J0, L0, (int)blockIdx.y, &sm.full[s]);
    tma_load_3d(d + 5 * kWHaloPad + kWInnerElems, &maps.m[13], J0, L0, (int)blockIdx.y, &sm.full[s]);
  };
  const int step_i = (int)gridDim.x / tiles_j, step_j = (int)gridDim.x - step_i * tiles_j;
  int tile_i = (int)blockIdx.x / tiles_j, tile_j = (int)blockIdx.x - tile_i * tiles_j;
  if (tid == 0 && (int)blockIdx.x < ntiles) issue(tile_i, tile_j, 0);
  double rr = 0.0, rabs = 0.0;
  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int s = it & 1;
    const unsigned parity = (unsigned)(it >> 1) & 1u;
    int next_i = tile_i + step_i, next_j = tile_j + step_j;
    if (next_j >= tiles_j) { next_j -= tiles_j; ++next_i; }
    if (tid == 0 && tile + (int)gridDim.x < ntiles) issue(next_i, next_j, s ^ 1);
    const int L0 = li_begin + tile_i * kTI;
    const int I0 = gi0 + L0, J0 = tile_j * kWJ;
    const double* fs = sm.stage[s];                  // [10][68] halo boxes; H(arr, a', b') <-> cell (I0 + a' - 1, J0 + b' - 1)
    const double* ys = fs + kWHaloPad;
    const double* txs = ys + kWHaloPad;
    const double* tys = txs + kWHaloPad;
    const double* cxs = tys + kWHaloPad;
    const double* Us = cxs + kWHaloPad;              // [8][64] interior tiles
    const double* Uds = Us + kWInnerElems;
#define H(arr, aa, bb) arr[(aa) * kWHaloJ + (bb) + 1]
    const bool edge_tile = I0 == 0 || I0 + kTI >= nx || J0 == 0 || J0 + kWJ >= ny || L0 + kTI > li_end;
    // vertex weights (interior tiles): row a, columns 2b, 2b + 1
    double wl = 0.0, wr = 0.0, wb0 = 0.0, wt0 = 0.0, wb1 = 0.0, wt1 = 0.0;
    if (!edge_tile) {
      wl = g.wxL[I0 + a]; wr = g.wxR[I0 + a];
      wb0 = g.wyB[J0 + j0]; wt0 = g.wyT[J0 + j0]; wb1 = g.wyB[J0 + j0 + 1]; wt1 = g.wyT[J0 + j0 + 1];
    }
    mbar_wait(&sm.full[s], parity);
    // generic per-item code (boundary tiles, extra rows / columns): the expressions of k_assemble_tma
    auto vertex = [&](int va, int vb) {
      const int vi = I0 + va, vj = J0 + vb;
      double v = 0.0;
      if (!edge_tile) {
        const double xl = g.wxL[vi], xr = g.wxR[vi], yb = g.wyB[vj], yt = g.wyT[vj];
        v = xl * yb * H(fs, va, vb) + xr * yb * H(fs, va + 1, vb) + xl * yt * H(fs, va, vb + 1) + xr * yt * H(fs, va + 1, vb + 1);
      } else if (vi <= nx && vj <= ny) {
        v = vertex_value(g, vi, vj, H(fs, va, vb), H(fs, va + 1, vb), H(fs, va, vb + 1), H(fs, va + 1, vb + 1));
      }
      sm.vs[va][vb] = v;
    };
    auto wface = [&](int fa, int fb, double& AK, double& AL) {
      const int i = I0 + fa, j = J0 + fb;
      AK = 0.0; AL = 0.0;
      if (!edge_tile || (i >= 1 && i <= nx - 1 && j < ny)) {
        const double tK = H(txs, fa + 1, fb + 1), cK = H(cxs, fa + 1, fb + 1), tL = H(txs, fa, fb + 1), cL = H(cxs, fa, fb + 1);
        const double vSW = sm.vs[fa][fb], vNW = sm.vs[fa][fb + 1];
        const double kA = tK - cK, kB = tK + cK;
        const double lA = tL - cL, lB = tL + cL;
        face_pair(kA * vNW + kB * vSW, kA + kB, H(fs, fa + 1, fb + 1), lA * vSW + lB * vNW, lA + lB, H(fs, fa, fb + 1), AK, AL);
      }
    };
    auto sface = [&](int fa, int fb, double& AK, double& AL) {
      const int i = I0 + fa, j = J0 + fb;
      AK = 0.0; AL = 0.0;
      if (!edge_tile || (j >= 1 && j <= ny - 1 && i < nx)) {
        const double tK = H(tys, fa + 1, fb + 1), cK = H(cxs, fa + 1, fb + 1), tL = H(tys, fa + 1, fb), cL = H(cxs, fa + 1, fb);
        const double vSW = sm.vs[fa][fb], vSE = sm.vs[fa + 1][fb];
        const double kA = tK + cK, kB = tK - cK;
        const double lA = tL + cL, lB = tL - cL;
        face_pair(kA * vSW + kB * vSE, kA + kB, H(fs, fa + 1, fb + 1), lA * vSE + lB * vSW, lA + lB, H(fs, fa + 1, fb), AK, AL);
      }
    };
    // ---- 1. vertices (I0 + a, J0 + 2b), (I0 + a, J0 + 2b + 1) ----
    if (!edge_tile) {
      // f of the cells (a-1 .. a) x (2b-1 .. 2b+1): halo rows a, a + 1, halo columns 2b, 2b + 1, 2b + 2 (box index + 1)
      const double fm0 = H(fs, a, j0), fm1 = H(fs, a + 1, j0);
      const double2 f0 = lds2(&H(fs, a, j0 + 1)), f1 = lds2(&H(fs, a + 1, j0 + 1));
      const double v0 = wl * wb0 * fm0 + wr * wb0 * fm1 + wl * wt0 * f0.x + wr * wt0 * f1.x;
      const double v1 = wl * wb1 * f0.x + wr * wb1 * f1.x + wl * wt1 * f0.y + wr * wt1 * f1.y;
      sts2(&sm.vs[a][j0], v0, v1);
    } else {
      vertex(a, j0);
      vertex(a, j0 + 1);
    }
    if (a == 1) { vertex(kTI, j0); vertex(kTI, j0 + 1); }
    if (a == 2 && b <= kTI) vertex(b, kWJ);
    __syncthreads();
    // ---- 2. faces: west faces of both cells, south faces of both cells ----
    double wK0, wL0, wK1, wL1, sK0, sL0, sK1, sL1;
    if (!edge_tile) {
      const double2 tK = lds2(&H(txs, a + 1, j0 + 1)), cK = lds2(&H(cxs, a + 1, j0 + 1)), fK = lds2(&H(fs, a + 1, j0 + 1));
      const double2 tL = lds2(&H(txs, a, j0 + 1)), cL = lds2(&H(cxs, a, j0 + 1)), fL = lds2(&H(fs, a, j0 + 1));
      const double2 v01 = lds2(&sm.vs[a][j0]);
      const double v2 = sm.vs[a][j0 + 2];
      {   // W face of (a, 2b): vertices SW = v01.x, NW = v01.y
        const double kA = tK.x - cK.x, kB = tK.x + cK.x, lA = tL.x - cL.x, lB = tL.x + cL.x;
        face_pair(kA * v01.y + kB * v01.x, kA + kB, fK.x, lA * v01.x + lB * v01.y, lA + lB, fL.x, wK0, wL0);
      }
      {   // W face of (a, 2b + 1): SW = v01.y, NW = v2
        const double kA = tK.y - cK.y, kB = tK.y + cK.y, lA = tL.y - cL.y, lB = tL.y + cL.y;
        face_pair(kA * v2 + kB * v01.y, kA + kB, fK.y, lA * v01.y + lB * v2, lA + lB, fL.y, wK1, wL1);
      }
      const double2 yK = lds2(&H(tys, a + 1, j0 + 1));
      const double yS = H(tys, a + 1, j0), cS = H(cxs, a + 1, j0), fS = H(fs, a + 1, j0);
      const double2 u01 = lds2(&sm.vs[a + 1][j0]);
      {   // S face of (a, 2b): K = (a, 2b), L = (a, 2b - 1); SW = v01.x, SE = u01.x
        const double kA = yK.x + cK.x, kB = yK.x - cK.x, lA = yS + cS, lB = yS - cS;
        face_pair(kA * v01.x + kB * u01.x, kA + kB, fK.x, lA * u01.x + lB * v01.x, lA + lB, fS, sK0, sL0);
      }
      {   // S face of (a, 2b + 1): K = (a, 2b + 1), L = (a, 2b); SW = v01.y, SE = u01.y
        const double kA = yK.y + cK.y, kB = yK.y - cK.y, lA = yK.x + cK.x, lB = yK.x - cK.x;
        face_pair(kA * v01.y + kB * u01.y, kA + kB, fK.y, lA * u01.y + lB * v01.y, lA + lB, fK.x, sK1, sL1);
      }
    } else {
      wface(a, j0, wK0, wL0);
      wface(a, j0 + 1, wK1, wL1);
      sface(a, j0, sK0, sL0);
      sface(a, j0 + 1, sK1, sL1);
    }
    sts2(&sm.WK[a][j0], wK0, wK1);
    sts2(&sm.WL[a][j0], wL0, wL1);
    sm.SK[a][b] = sK0;
    sm.SL[a][b] = sL0;
    if (a == 3) {   // the ninth west-face row
      double k0, l0, k1, l1;
      wface(kTI, j0, k0, l0);
      wface(kTI, j0 + 1, k1, l1);
      sts2(&sm.WK[kTI][j0], k0, k1);
      sts2(&sm.WL[kTI][j0], l0, l1);
    }
    if (a == 4 && b < kTI) {   // the south faces of column 64
      double k0, l0;
      sface(b, kWJ, k0, l0);
      sm.SK[b][kTJ] = k0;
      sm.SL[b][kTJ] = l0;
    }
    __syncthreads();
    // ---- 3. rows ----
    {
      const int i = I0 + a, j = J0 + j0;
      const bool ok0 = !edge_tile || (i < nx && L0 + a < li_end && j < ny);
      const bool ok1 = !edge_tile || (i < nx && L0 + a < li_end && j + 1 < ny);
      const size_t c0 = base + (size_t)(L0 + a) * ny + j;
      const double2 eK = lds2(&sm.WK[a + 1][j0]), eL = lds2(&sm.WL[a + 1][j0]);   // east faces = west faces of the row above
      const double nK = sm.SK[a][b + 1], nL = sm.SL[a][b + 1];                     // north face of the second cell
      const double2 f00 = lds2(&H(fs, a + 1, j0 + 1)), y00 = lds2(&H(ys, a + 1, j0 + 1));
      const double2 fW = lds2(&H(fs, a, j0 + 1)), yW = lds2(&H(ys, a, j0 + 1));
      const double2 fE = lds2(&H(fs, a + 2, j0 + 1)), yE = lds2(&H(ys, a + 2, j0 + 1));
      const double fSo = H(fs, a + 1, j0), ySo = H(ys, a + 1, j0), fNo = H(fs, a + 1, j0 + 3), yNo = H(ys, a + 1, j0 + 3);
      const double2 Uv = lds2(&Us[a * kWJ + j0]), Udv = lds2(&Uds[a * kWJ + j0]);
      // one cell: its four faces (west, east, south, north as (A_K own, A_L other) pairs), the neighbours' scales, U, Ud
      auto row = [&](int jj, double WKo, double WLo, double EKo, double ELo, double SKo, double SLo, double NKo, double NLo, double f0, double y0,
                     double csW, double csE, double csS, double csN, double Uc, double Udc, double& wW, double& wE, double& wS, double& wN,
                     double& rhs, double& cs0, double& om) {
        const int jc = J0 + jj, fb = jj;
        double diag = 0.0, R = 0.0, oW = 0.0, oE = 0.0, oS = 0.0, oN = 0.0;
        if (!edge_tile) {
          diag = WKo; oW = -WLo;
          diag += ELo; oE = -EKo;
          diag += SKo; oS = -SLo;
          diag += NLo; oN = -NKo;
        } else {
          if (i > 0) { diag += WKo; oW = -WLo; }
          if (i < nx - 1) { diag += ELo; oE = -EKo; }
          if (jc > 0) { diag += SKo; oS = -SLo; }
          if (jc < ny - 1) { diag += NLo; oN = -NKo; }
          if (i == 0 || i == nx - 1 || jc == 0 || jc == ny - 1) {  // Dirichlet boundary faces (Solver.cc:143-164, 204-267)
            const double txP = H(txs, a + 1, fb + 1), tyP = H(tys, a + 1, fb + 1), cP = H(cxs, a + 1, fb + 1);
            const double vSW = sm.vs[a][fb], vSE = sm.vs[a + 1][fb], vNW = sm.vs[a][fb + 1], vNE = sm.vs[a + 1][fb + 1];
            if (i == 0 && g.bc[0] == 0) diag += dirichlet_face((txP - cP) * vNW + (txP + cP) * vSW, (txP - cP) + (txP + cP), f0, R);
            if (i == nx - 1 && g.bc[1] == 0) diag += dirichlet_face((txP - cP) * vSE + (txP + cP) * vNE, (txP - cP) + (txP + cP), f0, R);
            if (jc == 0 && g.bc[2] == 0) diag += dirichlet_face((tyP + cP) * vSW + (tyP - cP) * vSE, (tyP + cP) + (tyP - cP), f0, R);
            if (jc == ny - 1 && g.bc[3] == 0) diag += dirichlet_face((tyP + cP) * vNE + (tyP - cP) * vNW, (tyP + cP) + (tyP - cP), f0, R);
          }
        }
        diag += Udc;
        R += Uc * f0;
        cs0 = f0 * y0;
        om = diag * cs0;
        const double dscale = sy2d_div(1.0, om);
        wW = oW * csW * dscale; wE = oE * csE * dscale;
        wS = oS * csS * dscale; wN = oN * csN * dscale;
        rhs = R * dscale - 1.0 - ((wW + wE) + (wS + wN));
      };
      double wWa, wEa, wSa, wNa, rha, csa, oma, wWb, wEb, wSb, wNb, rhb, csb, omb;
      // cell (a, 2b): north face = south face of the thread's second cell
      row(j0, wK0, wL0, eK.x, eL.x, sK0, sL0, sK1, sL1, f00.x, y00.x, fW.x * yW.x, fE.x * yE.x, fSo * ySo, f00.y * y00.y, Uv.x, Udv.x,
          wWa, wEa, wSa, wNa, rha, csa, oma);
      // cell (a, 2b + 1): south neighbour = the first cell
      row(j0 + 1, wK1, wL1, eK.y, eL.y, sK1, sL1, nK, nL, f00.y, y00.y, fW.y * yW.y, fE.y * yE.y, f00.x * y00.x, fNo * yNo, Uv.y, Udv.y,
          wWb, wEb, wSb, wNb, rhb, csb, omb);
      if (ok0 && ok1) {
        stg2(o.wW + c0, wWa, wWb); stg2(o.wE + c0, wEa, wEb); stg2(o.wS + c0, wSa, wSb); stg2(o.wN + c0, wNa, wNb);
        stg2(o.rhs + c0, rha, rhb);
        stg2(o.cs + c0, csa, csb);
        if (o.om) stg2(o.om + c0, oma, omb);
        rr += rha * rha; rabs = nmax(rabs, fabs(rha));
        rr += rhb * rhb; rabs = nmax(rabs, fabs(rhb));
      } else if (ok0) {
        o.wW[c0] = wWa; o.wE[c0] = wEa; o.wS[c0] = wSa; o.wN[c0] = wNa; o.rhs[c0] = rha; o.cs[c0] = csa;
        if (o.om) o.om[c0] = oma;
        rr += rha * rha; rabs = nmax(rabs, fabs(rha));
      }
    }
#undef H
    tile_i = next_i; tile_j = next_j;
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
