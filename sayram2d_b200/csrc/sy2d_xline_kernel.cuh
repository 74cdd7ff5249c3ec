// Engine 2, x-line variant: one persistent CTA per problem, BiCGSTAB right-preconditioned with
// the tridiagonal T = tridiag(wW, 1, wE) along i (alpha0, the stiff direction), state resident
// in registers / shared memory.
//
//   phat = T^-1 p   (Thomas sweeps along i),   v = A phat = p + wS phat_S + wN phat_N
// (T phat = p holds by construction, so the W/E couplings never have to be applied again and
// wW, wE are needed only as the LU factors of T).  Same for s.
//
// Pivot scaling: the rows of the system are divided once per time step by the pivots d_i of T's LU,
//   A' = D^-1 A,  rhs' = D^-1 rhs,  T' = D^-1 T = (I + L')(I + U'),  l'_i = wW_i / d_i,  e_i = wE_i / d_i,
// so both triangular factors have a UNIT diagonal: a solve is z_i = b_i - l'_i z_{i-1}, x_i = z_i - e_i x_{i+1}
// (two factor arrays and two FMAs per row instead of three and three), wS' = wS / d and wN' = wN / d carry the
// scaling, and the shared-memory array that held 1/d holds the search direction p instead, which with r/s and
// the Thomas work vector does not fit in the 96 registers of a 640-thread CTA.  0 < d <= 1 (M-matrix), so the
// stopping rule max|r'| <= tol on the scaled residual r' = r / d is at least as strict as the one on r.  Right
// preconditioning leaves the residual - and therefore the stopping rule max|r| <= tol and the
// accuracy of f - exactly those of the unpreconditioned engine; iterations drop ~4x
// (80x80 AY: 59 -> 14.5 per step).
//
// Bank conflicts: lane (k, jj) touches hat[(k*R + m)*hs + j]; 64-bit accesses are served per half-warp
// (lanes 0-15 = columns jj, jj+1 x the eight k), so with R*hs = 2 (mod 16) the 16 lanes of a half-warp
// land on 16 distinct eight-byte slots (hs = 85 for the 80 x 80 instance).
//
// Thread layout: a warp owns CPW = 32/NCH columns j; the NCH lanes of a column own R consecutive
// rows each (lane = jj*NCH + k, rows k*R .. k*R+R-1).  A Thomas sweep is then a chain over the
// NCH lanes of a column, handed from lane to lane with a warp shuffle: no block barrier inside a
// sweep, every warp sweeps its own columns independently.
//
// Placement per problem (N = nx*ny cells, S = R*NT slots in thread-private [m][tid] layout):
//   registers : r/s of the owned cells (resident), phat/shat, t, v (transient)
//   shared    : hat (phat/shat, natural (i,j) layout: the S/N neighbour exchange),
//               l', e (sweep coefficients) and p, private layout, conflict-free
//   L2        : wS, wN, v, y, rhs (private layout => fully coalesced 128 B lines per warp)
// Traffic per cell and iteration: 10 L2 accesses (80 B: v r/w, wS wN twice, rhat twice, y r/w) + 18 shared
// accesses (two Thomas solves of 4, two publishes, four neighbour reads, p written once and read three times).
#pragma once
#include "sy2d_assemble_tma.cuh"   // smem_u32, mbar_init / mbar_expect_tx / mbar_wait
#include "sy2d_problem_kernel.cuh"
#include "sy2d_tmem.cuh"

#ifndef SY2D_XLINE_RHREG
#define SY2D_XLINE_RHREG 0   // experiment, slower (see RHREG below)
#endif

namespace sy2d {

constexpr int kXlineNCH = 8;          // lanes per column
constexpr int kXlineCPW = 32 / kXlineNCH;
constexpr int kXlineScratchArrays = 6;   // per CTA: wS, wN, v, y, rhs (thread-private layout) + one staging array of the block assembly

// Work queue of one launch.  An item is "the next `chunk` time steps of problem p"; the persistent CTAs (one per SM)
// pop tickets from `head`, wait until the ticket's slot holds a problem, run the chunk and - when the problem has steps
// left - append it at `tail`.  The first nprob slots are filled by the host in issue order (most expensive problems
// first), so the queue is a round robin over the problems: nprob * nsteps / chunk work units over the SMs instead of
// nprob units of nsteps steps each - with 512 members on 148 SMs (the per-GPU share of the 4096-member ensemble on 8
// GPUs) the last, partly filled wave costs 1 / 70 of the call instead of 0.54 / 4.  A problem is worked on by one CTA
// at a time; its state between steps is f and yprev in global memory, handed over with a release store of the slot /
// an acquire load by the popping thread followed by the CTA barrier.
struct XlineQueue {
  int head, tail, total, pad;
};

struct XlineArgs {
  ProblemArgs a;
  double* scratch;   // [CTAs of the launch][5][S]: per-step temporaries of the CTA (not of the problem)
  int NT;            // threads per CTA = ny_pad * NCH
  int S;             // R * NT
  int hs;            // row stride of hat in shared memory (>= ny)
  XlineQueue* q;     // control words of this launch
  int* slots;        // [nprob * chunks_per_problem] problem ids (local to the launch's first problem), -1 = not yet pushed
  int* steps_done;   // [nbatch] time steps of the call completed per problem (zeroed by the host)
  int chunk;         // time steps per work item
  int nchunks;       // work items per problem = ceil(nsteps / chunk)
  // Host-resident f without copy engines (sy2d_step_host with pinned, device-accessible buffers): the CTA that starts a
  // problem's first item pulls its f straight from host memory over PCIe / C2C, the CTA that finishes its last item
  // pushes the result back - in the queue's issue order (most expensive problems first), so that no problem waits for
  // the upload of a sub-batch it happens to sit in and nothing is left to download when the last CTA ends.
  const double* hin;   // [nbatch][N] device-accessible host pointer or NULL
  double* hout;        // [nbatch][N] device-accessible host pointer or NULL
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// TRAIL = false drops the trailing barrier: the caller must then not reuse `red` before every thread has
// passed a later barrier (the iteration loop rotates over three buffers).
template <int NV, bool TRAIL = true>
__device__ __forceinline__ void cta_reduce_x(double (&v)[NV], int nsum, double* red) {
  // first nsum entries are sums, the rest maxima; all threads get the result
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = k < nsum ? warp_sum(v[k]) : warp_max(v[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) red[k * 32 + w] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const double y = lane < nw ? red[k * 32 + lane] : 0.0;   // maxima here are of non-negative values
    v[k] = k < nsum ? warp_sum(y) : warp_max(y);
  }
  if (TRAIL) __syncthreads();
}

// Two sums with ONE shuffle tree: after the first exchange the lower half-warp carries the partial sums of a, the upper
// half-warp those of b, and the remaining four steps reduce a single double (a 64-bit shuffle is two SHFL instructions,
// so two sums cost 10 + 4 of them per phase instead of 20).  No trailing barrier: the caller rotates its buffers.
__device__ __forceinline__ void cta_reduce_pair(double& a, double& b, double* red /* >= 64 */) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const bool up = (lane & 16) != 0;
  double keep = up ? b : a;
  keep += __shfl_xor_sync(full, up ? a : b, 16);
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) keep += __shfl_xor_sync(full, keep, o);
  if ((lane & 15) == 0) red[(up ? 32 : 0) + w] = keep;
  __syncthreads();
  const double xa = lane < nw ? red[lane] : 0.0, xb = lane < nw ? red[32 + lane] : 0.0;
  keep = up ? xb : xa;
  keep += __shfl_xor_sync(full, up ? xa : xb, 16);
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) keep += __shfl_xor_sync(full, keep, o);
  a = __shfl_sync(full, keep, 0);
  b = __shfl_sync(full, keep, 16);
}

// One sum and the maximum of non-negative values.  The maximum goes through REDUX on the upper 32 bits of the doubles
// (non-negative doubles order like their bit patterns; NaN - upper word 0x7ff8.... - wins, as it must): one instruction
// per phase instead of a five-step tree, and the result is an UPPER bound (lower word all ones, relative excess
// < 2^-20), so the stopping test max|r| <= tol it feeds can only become stricter.
__device__ __forceinline__ void cta_reduce_sum_max(double& sum, double& mx, double* red /* >= 64 */) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  sum = warp_sum(sum);
  unsigned hi = __reduce_max_sync(full, (unsigned)__double2hiint(mx));
  if (lane == 0) { red[w] = sum; reinterpret_cast<unsigned*>(red + 32)[w] = hi; }
  __syncthreads();
  sum = warp_sum(lane < nw ? red[lane] : 0.0);
  hi = __reduce_max_sync(full, lane < nw ? reinterpret_cast<const unsigned*>(red + 32)[lane] : 0u);
  mx = __hiloint2double((int)hi, (int)0xffffffffu);
}

// one scaled row; kept out of line so that the R-times unrolled assembly loop stays small
__device__ __noinline__ Scaled assemble_scaled_cell(const double* f, const double* yprev, const double* __restrict__ tx,
                                                    const double* __restrict__ ty, const double* __restrict__ cxy,
                                                    const double* __restrict__ U, const double* __restrict__ Ud,
                                                    const Geometry& g, int i, int j) {
  const int nx = g.nx, ny = g.ny;
  Row row;
  assemble_row(f, tx, ty, cxy, U, Ud, g, i, j, row);
  const int n = i * ny + j;
  const int nW = i > 0 ? n - ny : n, nE = i < nx - 1 ? n + ny : n, nS = j > 0 ? n - 1 : n, nN = j < ny - 1 ? n + 1 : n;
  Scaled sc;
  scale_row(row, yprev[n], yprev[nW], yprev[nE], yprev[nS], yprev[nN], sc);
  return sc;
}

// ---------------------------------------------------------------------------------------------------------------
// Assembly of a FULL tile (nx = 8 R rows, ny = NT / 8 columns) by the whole CTA, block by block.
//
// The marching assembly (a thread walks up the R rows it owns, loading what each row needs when it gets there) spends
// ~20 cycles per cell waiting for L2: ten-odd dependent round trips per row with one row in flight per thread.  Here
// the CTA works through the grid in blocks of BR = 8 rows x ny columns - exactly one cell per thread - whose inputs
// (f, yprev, tx, cxy with one halo row on each side; ty, U, Ud) one thread fetches with seven bulk copies
// (cp.async.bulk global -> shared, mbarrier complete_tx) into a two-stage ring, so the loads of the next block are in
// flight while this one is computed from shared memory:
//   1. vertex values of the block's 9 x (ny+1) vertices                      -> V
//   2. per cell its west and south face (A_K, A_L), each face evaluated once  -> XK/XL, YK/YL
//   3. per cell the row: its four faces, boundary faces, U, the scaling       -> wW, wE (shared), wS, wN, rhs (scratch)
// Blocks run from the top of the grid down, so the east faces of a block's last row are the west faces of the block
// done just before (kept as row BR of XK/XL).  The ring, V and the face tiles live in the hat + p regions (unused during
// the assembly); outputs go to NATURAL (i, j) layouts: wW, wE with row stride LS (LS * R = 10 (mod 16): the 16 lanes of
// a half-warp of the solver - eight row chunks x two columns - hit 16 distinct banks), wS, wN, rhs with row stride ny
// (consecutive lanes of a row write consecutive doubles).  Same per-face expressions, operand and accumulation order as
// assemble_row.
template <int NX, int NY>
struct XlAsm {
  static constexpr int BR = 8;
  static constexpr int NBLK = NX / BR;
  static constexpr int HROWS = BR + 2;                                   // f, yprev, tx, cxy: with halo rows
  static constexpr int STAGE = 4 * HROWS * NY + 3 * BR * NY;             // doubles per ring stage
  static constexpr int oF = 0, oY = HROWS * NY, oTX = 2 * HROWS * NY, oCX = 3 * HROWS * NY, oTY = 4 * HROWS * NY,
                       oU = 4 * HROWS * NY + BR * NY, oUD = 4 * HROWS * NY + 2 * BR * NY;
  static constexpr int VW = NY + 1;                                      // vertex tile row length
  static constexpr int oV = 2 * STAGE;
  static constexpr int oXK = oV + (((BR + 1) * VW + 1) & ~1);
  static constexpr int oXL = oXK + (BR + 1) * NY;
  static constexpr int oYK = oXL + (BR + 1) * NY;
  static constexpr int oYL = oYK + BR * NY;
  static constexpr int END = oYL + BR * NY;                              // doubles of shared memory the assembly needs
  static_assert(NX % BR == 0, "full tiles only");
};

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// `area`: >= XlAsm::END doubles of shared memory (16-byte aligned), `full`: two initialised mbarriers (count 1), `ph`: their
// phase parities (bits 0, 1; kept by the caller across calls).  All NT = BR * NY threads call it; ends with a CTA barrier.
template <int NX, int NY, int LS>
__device__ __forceinline__ void xline_assemble_blocks(double* area, unsigned long long* full, unsigned& ph, const double* f,
                                                      const double* yprev, const double* __restrict__ tx,
                                                      const double* __restrict__ ty, const double* __restrict__ cxy,
                                                      const double* __restrict__ U, const double* __restrict__ Ud, const Geometry& g,
                                                      double* l_s, double* e_s, double* wS_g, double* wN_g, double* rhs_g) {
  using A = XlAsm<NX, NY>;
  constexpr int BR = A::BR, NBLK = A::NBLK, VW = A::VW;
  const int tid = threadIdx.x;
  const int li = tid / NY, jc = tid - li * NY;
  double* V = area + A::oV;
  double* XK = area + A::oXK;
  double* XL = area + A::oXL;
  double* YK = area + A::oYK;
  double* YL = area + A::oYL;
  auto issue = [&](int b, int s) {   // one thread: arm the stage's barrier, launch the seven copies of block b
    const int r0 = b * BR;
    const int lo = r0 > 0 ? r0 - 1 : 0, hi = r0 + BR < NX ? r0 + BR : NX - 1;
    const int nrow = hi - lo + 1, drow = lo - (r0 - 1);
    double* st = area + s * A::STAGE;
    mbar_expect_tx(&full[s], (unsigned)((4 * nrow + 3 * BR) * NY * sizeof(double)));
    const unsigned hb = (unsigned)(nrow * NY * sizeof(double)), cb = (unsigned)(BR * NY * sizeof(double));
    bulk_g2s(st + A::oF + drow * NY, f + lo * NY, hb, &full[s]);
    bulk_g2s(st + A::oY + drow * NY, yprev + lo * NY, hb, &full[s]);
    bulk_g2s(st + A::oTX + drow * NY, tx + lo * NY, hb, &full[s]);
    bulk_g2s(st + A::oCX + drow * NY, cxy + lo * NY, hb, &full[s]);
    bulk_g2s(st + A::oTY, ty + r0 * NY, cb, &full[s]);
    bulk_g2s(st + A::oU, U + r0 * NY, cb, &full[s]);
    bulk_g2s(st + A::oUD, Ud + r0 * NY, cb, &full[s]);
  };
  if (tid == 0) {
    // the ring overwrites shared memory the generic proxy wrote (hat, p) and reads global memory the generic proxy
    // wrote (f, yprev of the last step): order both against the async proxy
    asm volatile("fence.proxy.async;" ::: "memory");
    issue(NBLK - 1, 0);
    if (NBLK > 1) issue(NBLK - 2, 1);
  }
#pragma unroll 1
  for (int n = 0; n < NBLK; ++n) {
    const int b = NBLK - 1 - n, s = n & 1;
    const int r0 = b * BR, i = r0 + li;
    const double* st = area + s * A::STAGE;
    // row i of a staged array with halo rows sits at slot i - (r0 - 1)
    const double* F = st + A::oF - (r0 - 1) * NY;
    const double* Y = st + A::oY - (r0 - 1) * NY;
    const double* TX = st + A::oTX - (r0 - 1) * NY;
    const double* CX = st + A::oCX - (r0 - 1) * NY;
    mbar_wait(&full[s], (ph >> s) & 1u);
    ph ^= 1u << s;
    // 1. vertices (r0 + lv, vj), lv = 0 .. BR, vj = 0 .. NY
    auto vtx = [&](int lv, int vj) {
      const int vi = r0 + lv;
      const int il = vi > 0 ? vi - 1 : 0, ih = vi < NX ? vi : NX - 1;
      const int jb = vj > 0 ? vj - 1 : 0, jh = vj < NY ? vj : NY - 1;
      return vertex_value(g, vi, vj, F[il * NY + jb], F[ih * NY + jb], F[il * NY + jh], F[ih * NY + jh]);
    };
    V[li * VW + jc] = vtx(li, jc);
    if (tid < NY) V[BR * VW + tid] = vtx(BR, tid);
    else if (tid < NY + BR + 1) V[(tid - NY) * VW + NY] = vtx(tid - NY, NY);
    __syncthreads();
    // 2. west and south face of the cell
    const double f00 = F[i * NY + jc], txP = TX[i * NY + jc], cP = CX[i * NY + jc], tyP = st[A::oTY + li * NY + jc];
    const double vSW = V[li * VW + jc], vSE = V[(li + 1) * VW + jc], vNW = V[li * VW + jc + 1], vNE = V[(li + 1) * VW + jc + 1];
    {
      if (li == 0 && n > 0) {   // the block above left its first row of west faces in row 0: they are this block's row BR
        XK[BR * NY + jc] = XK[jc];
        XL[BR * NY + jc] = XL[jc];
      }
      double AKw = 0.0, ALw = 0.0;
      if (i > 0) {   // K = (i, j), L = (i-1, j)
        const double t = TX[(i - 1) * NY + jc], c = CX[(i - 1) * NY + jc];
        const double aW_A = txP - cP, aW_B = txP + cP, lA = t - c, lB = t + c;
        face_pair(aW_A * vNW + aW_B * vSW, aW_A + aW_B, f00, lA * vSW + lB * vNW, lA + lB, F[(i - 1) * NY + jc], AKw, ALw);
      }
      double AKs = 0.0, ALs = 0.0;
      if (jc > 0) {   // K = (i, j), L = (i, j-1)
        const double t = st[A::oTY + li * NY + jc - 1], c = CX[i * NY + jc - 1];
        const double aS_A = tyP + cP, aS_B = tyP - cP, lA = t + c, lB = t - c;
        face_pair(aS_A * vSW + aS_B * vSE, aS_A + aS_B, f00, lA * vSE + lB * vSW, lA + lB, F[i * NY + jc - 1], AKs, ALs);
      }
      XK[li * NY + jc] = AKw; XL[li * NY + jc] = ALw;
      YK[li * NY + jc] = AKs; YL[li * NY + jc] = ALs;
    }
    __syncthreads();
    // 3. the row
    {
      Row row;
      double diag = 0.0, Rr = 0.0;
      row.oW = 0.0; row.oE = 0.0; row.oS = 0.0; row.oN = 0.0;
      if (i > 0) { diag += XK[li * NY + jc]; row.oW = -XL[li * NY + jc]; }
      else if (g.bc[0] == 0) diag += dirichlet_face((txP - cP) * vNW + (txP + cP) * vSW, (txP - cP) + (txP + cP), f00, Rr);
      if (i < NX - 1) { diag += XL[(li + 1) * NY + jc]; row.oE = -XK[(li + 1) * NY + jc]; }
      else if (g.bc[1] == 0) diag += dirichlet_face((txP - cP) * vSE + (txP + cP) * vNE, (txP - cP) + (txP + cP), f00, Rr);
      if (jc > 0) { diag += YK[li * NY + jc]; row.oS = -YL[li * NY + jc]; }
      else if (g.bc[2] == 0) diag += dirichlet_face((tyP + cP) * vSW + (tyP - cP) * vSE, (tyP + cP) + (tyP - cP), f00, Rr);
      if (jc < NY - 1) { diag += YL[li * NY + jc + 1]; row.oN = -YK[li * NY + jc + 1]; }
      else if (g.bc[3] == 0) diag += dirichlet_face((tyP + cP) * vNE + (tyP - cP) * vNW, (tyP + cP) + (tyP - cP), f00, Rr);
      diag += st[A::oUD + li * NY + jc];
      Rr += st[A::oU + li * NY + jc] * f00;
      row.diag = diag; row.R = Rr; row.f00 = f00;
      const int iW = i > 0 ? i - 1 : i, iE = i < NX - 1 ? i + 1 : i, jS = jc > 0 ? jc - 1 : jc, jN = jc < NY - 1 ? jc + 1 : jc;
      row.fW = F[iW * NY + jc]; row.fE = F[iE * NY + jc]; row.fS = F[i * NY + jS]; row.fN = F[i * NY + jN];
      Scaled sc;
      scale_row(row, Y[i * NY + jc], Y[iW * NY + jc], Y[iE * NY + jc], Y[i * NY + jS], Y[i * NY + jN], sc);
      l_s[i * LS + jc] = sc.wW;   // raw wW, wE until the factorisation
      e_s[i * LS + jc] = sc.wE;
      wS_g[i * NY + jc] = sc.wS; wN_g[i * NY + jc] = sc.wN; rhs_g[i * NY + jc] = sc.rhs;
    }
    __syncthreads();   // the stage, V and the face tiles are free
    if (tid == 0 && n + 2 < NBLK) issue(NBLK - 3 - n, s);
  }
}

// shared memory (doubles) of the full-tile instance: assembly area (>= hat + p) | wW/l' | wE/e | reduction buffers, work item, mbarriers
template <int R, int NT, int HS>
constexpr size_t xline_full_smem_doubles() {
  constexpr int NX = kXlineNCH * R, NY = NT / kXlineNCH, LS = NY + 1;
  constexpr size_t area = (size_t)(XlAsm<NX, NY>::END > NX * HS + R * NT ? XlAsm<NX, NY>::END : NX * HS + R * NT);
  return area + 2 * (size_t)NX * LS + 192 + 2 + 2 + (size_t)R * NT / 4;   // ... and the 16-bit shadow residual
}

// NTC > 0 fixes the CTA size at compile time (the 80-column production shape: 640 threads), so
// that every thread-private slot m*NT + tid is base + immediate and costs no address registers.
// HSC > 0 fixes the hat row stride too and promises a FULL tile (nx = NCH*R rows, ny = NT/NCH columns:
// every slot is a cell, no validity predicates): together the 80 x 80 instance.
template <int R, int MAXT, int NTC, int HSC>
__global__ void __launch_bounds__(MAXT, 1) k_problem_xline(XlineArgs xa) {
  constexpr int NCH = kXlineNCH;
  extern __shared__ double sm[];
  const ProblemArgs& a = xa.a;
  const int nx = a.g.nx, ny = a.g.ny, N = nx * ny;
  const int NT = NTC > 0 ? NTC : xa.NT;
  const int S = R * NT;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int k = lane % NCH, jj = lane / NCH;
  const int j = w * kXlineCPW + jj;
  constexpr bool FULL = HSC > 0;
  const bool col_ok = FULL || j < ny;
  const int i0 = k * R;
  const int hs = HSC > 0 ? HSC : xa.hs;   // hat row stride, chosen on the host so that the NCH lanes of a column hit distinct banks
  // Full tile: l', e in the natural (i, j) layout with row stride LS (written by the block assembly), the scaled wS', wN',
  // rhs' of the scratch in the natural layout with row stride ny; otherwise thread-private slots m * NT + tid.
  constexpr int NXC = NCH * R, NYC = (NTC > 0 ? NTC : NCH) / NCH, LS = NYC + 1;
  constexpr int AREA = FULL ? (XlAsm<NXC, NYC>::END > NXC * HSC + R * NTC ? XlAsm<NXC, NYC>::END : NXC * HSC + R * NTC) : 0;
  double* hat = sm;
  double* p_s = hat + nx * hs;   // the search direction p; during the assembly: 1/d of the factorisation
  double* l_s = FULL ? sm + AREA : p_s + S;
  double* e_s = l_s + (FULL ? NXC * LS : S);
  double* red = e_s + (FULL ? NXC * LS : S);
  int* s_item = reinterpret_cast<int*>(red + 192);
  unsigned long long* asm_bar = reinterpret_cast<unsigned long long*>(red + 194);   // full tile: the two stage barriers of the assembly ring
  unsigned asm_ph = 0;
  // Full tile: the shadow residual rhat of BiCGSTAB - any fixed vector with (rhat, r0) != 0 - is r0 TRUNCATED to the upper 16
  // bits of its doubles (sign, exponent, 4 bits of mantissa: rhat_i = r0_i (1 - eps_i), 0 <= eps_i < 1/16, so (rhat, r0) >=
  // 15/16 |r0|^2) and lives in 12.8 KB of shared memory instead of the L2 scratch: 16 B less L2 traffic per cell and iteration.
  // The arithmetic stays fp64; only the choice of the shadow vector differs from rhat = r0.
  constexpr bool RH16 = FULL;
  unsigned short* rh16 = reinterpret_cast<unsigned short*>(asm_bar + 2);
  const int lq0 = FULL ? i0 * LS + j : tid, lqs = FULL ? LS : NT;     // slot of owned row m in l_s / e_s: lq0 + m * lqs
  const int gq0 = FULL ? i0 * ny + j : tid, gqs = FULL ? ny : NT;     // ... in wS_g / wN_g / rhs_g
  // Full tile: wS', wN' and y of the owned rows live in TENSOR MEMORY (sy2d_tmem.cuh) instead of the L2 scratch: the thread's
  // TMEM lane, 10 R columns of the warp's range - (wS'_m, wN'_m) row by row in [0, 4R), y_m in [4R, 6R), the sweep factors l'_m in
  // [6R, 8R) and e_m in [8R, 10R) (TML: a line solve reads its factors four times per row; from tensor memory those 8 reads per
  // cell and iteration no longer go through the shared-memory pipe, the unit ncu shows the kernel is bound by).  The five warps
  // that share a lane quarter (w % 4) take column ranges 102 apart (5 x 102 <= 512; tcgen05.ld / st need no column alignment,
  // profiles/tmem_probe.cu).
  constexpr bool TMW = FULL;
  constexpr bool TML = TMW;
  // PTM (experiment, off): the search direction p (read three times and written twice per cell and iteration) takes y's place in
  // tensor memory and y (one read-modify-write) takes p's thread-private slots in shared memory - three shared-memory accesses
  // less, but every phase then waits for a tensor-memory round trip before it can start and the p batches cost registers
  // (spills 356 -> 512 bytes): 5.06 -> 5.73 ms per step at 4096 members.
  constexpr bool PTM = false;
  constexpr int kTmWarpCols = 102;
  static_assert(!TMW || (R == 10 && MAXT <= 640), "TMEM layout: 10 rows per thread, at most five warps per lane quarter");
  unsigned tcol = 0;
  if (FULL) {
    if (tid == 0) {
      mbar_init(&asm_bar[0], 1);
      mbar_init(&asm_bar[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (TMW) {
      if (w == 0) tmem_alloc_all(reinterpret_cast<unsigned*>(s_item + 1));
      tmem_fence_before_sync();
    }
    __syncthreads();
    if (TMW) {
      tmem_fence_after_sync();
      tcol = reinterpret_cast<volatile unsigned*>(s_item + 1)[0] + ((unsigned)(32 * (w & 3)) << 16) + (unsigned)((w >> 2) * kTmWarpCols);
    }
  }
  // rows m0 .. m0+3 / m0 .. m0+1 of (wS', wN') from tensor memory; fn(m, wS'_m, wN'_m) per row
  auto tm_rows = [&](auto fn) {
#pragma unroll
    for (int m0 = 0; m0 < 8; m0 += 4) {
      double t[8];
      tmem_ld<8>(t, tcol + 4 * m0);
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) fn(m0 + k4, t[2 * k4], t[2 * k4 + 1]);
    }
    double t[4];
    tmem_ld<4>(t, tcol + 32);
#pragma unroll
    for (int k2 = 0; k2 < 2; ++k2) fn(8 + k2, t[2 * k2], t[2 * k2 + 1]);
  };
  // the same with p_m (PTM): fn(m, wS'_m, wN'_m, p_m)
  auto tm_rows_p = [&](auto fn) {
#pragma unroll
    for (int m0 = 0; m0 < 8; m0 += 4) {
      double t[8], pp[4];
      tmem_ld<8>(t, tcol + 4 * m0);
      tmem_ld<4>(pp, tcol + 4 * R + 2 * m0);
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) fn(m0 + k4, t[2 * k4], t[2 * k4 + 1], pp[k4]);
    }
    double t[4], pp[2];
    tmem_ld<4>(t, tcol + 32);
    tmem_ld<2>(pp, tcol + 4 * R + 16);
#pragma unroll
    for (int k2 = 0; k2 < 2; ++k2) fn(8 + k2, t[2 * k2], t[2 * k2 + 1], pp[k2]);
  };
  double* scr = xa.scratch + (size_t)blockIdx.x * kXlineScratchArrays * S;
  double* wS_g = scr; double* wN_g = scr + S; double* v_g = scr + 2 * S; double* y_g = scr + 3 * S; double* rhs_g = scr + 4 * S;
  // full tile: the block assembly leaves wS, wN, rhs in the natural (i, j) layout (coalesced stores) in v, y and a sixth
  // array; the pivot-scaling pass moves them into the thread-private layout the iteration reads
  double* wS_n = FULL ? v_g : wS_g; double* wN_n = FULL ? y_g : wN_g; double* rhs_n = FULL ? scr + 5 * S : rhs_g;
  // RHREG (experiment, off): the ten 16-bit shadow-residual values of the owned rows packed into five registers instead of shared
  // memory (a warp's 2-byte reads cost a shared-memory wavefront each: 10 % of the kernel's wavefronts, ncu capture xu).  The
  // kernel sits at its 96-register limit: five more live registers raise the spills from 356 to 476 bytes, 4.75 -> 5.20 ms per step.
  constexpr bool RHREG = RH16 && SY2D_XLINE_RHREG;
  unsigned rhp[RHREG ? R / 2 : 1] = {};
  auto rhat_of = [&](int q, int m) {
    if (RHREG) return __hiloint2double((int)((m & 1) ? (rhp[RHREG ? m >> 1 : 0] & 0xffff0000u) : (rhp[RHREG ? m >> 1 : 0] << 16)), 0);
    return RH16 ? __hiloint2double((int)((unsigned)rh16[q] << 16), 0) : rhs_g[q];
  };
  const unsigned full = 0xffffffffu;

  double rs[R], z[R];
  // VREG: v stays in registers from the v phase to the end of the iteration, where the p array takes p - omega v (the next
  // search direction is then r + beta * that): no v array in the scratch, 16 B less L2 traffic per cell and iteration.
  constexpr bool VREG = FULL;
  double vr[VREG ? R : 1];

 for (;;) {   // work items: (problem, chunk of time steps)
  if (tid == 0) {
    const int ticket = atomicAdd(&xa.q->head, 1);
    int item = -1;
    while (ticket < ld_acquire_gpu(&xa.q->total)) {   // total shrinks when a problem fails: its remaining items never come
      item = ld_acquire_gpu(xa.slots + ticket);
      if (item >= 0) break;
      __nanosleep(200);
    }
    s_item[0] = item;
  }
  __syncthreads();
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
  const int step_begin = xa.steps_done[prob];
  const int step_end = min(step_begin + xa.chunk, a.nsteps);
  if (xa.hin && step_begin == 0) {   // first item of the problem: f from the host buffer (coalesced 16-byte loads over PCIe)
    const double* src = xa.hin + base;
    if ((N & 1) == 0) {
      for (int n = tid; n < N / 2; n += NT) reinterpret_cast<double2*>(f)[n] = reinterpret_cast<const double2*>(src)[n];
    } else {
      for (int n = tid; n < N; n += NT) f[n] = src[n];
    }
    __syncthreads();
  }
  int it_total = 0, it = 0, state = 1, steps_ok = step_begin;   // steps_ok: time steps of the call this problem has completed
  double rmax = 0.0, res_true = 0.0, res_rel = 0.0;

  for (int step = step_begin; step < step_end; ++step) {
    // ------------- assembly of the scaled rows owned by this thread -------------
    double acc[2] = {0.0, 0.0};
    if (FULL) {
      xline_assemble_blocks<NXC, NYC, LS>(sm, asm_bar, asm_ph, f, yprev, tx, ty, cxy, U, Ud, a.g, l_s, e_s, wS_n, wN_n, rhs_n);
    } else {
#pragma unroll
      for (int m = 0; m < R; ++m) {
        const int i = i0 + m, q = m * NT + tid;
        double wW = 0.0, wE = 0.0, wS = 0.0, wN = 0.0, rh = 0.0;
        if (col_ok && i < nx) {
          const Scaled sc = assemble_scaled_cell(f, yprev, tx, ty, cxy, U, Ud, a.g, i, j);
          wW = sc.wW; wE = sc.wE; wS = sc.wS; wN = sc.wN; rh = sc.rhs;
        }
        l_s[q] = wW;   // raw wW, wE until the factorisation below turns them into l and wE/d
        e_s[q] = wE;
        wS_g[q] = wS; wN_g[q] = wN; rhs_g[q] = rh;
      }
    }
    if (!FULL) __syncthreads();   // (the block assembly ends with a barrier)
    // LU of T down each column: chain over the NCH lanes of the column.  d_i = 1 - wW_i e_{i-1} (e = wE / d), so only
    // e is handed from row to row and from lane to lane; l' = wW / d, e = wE / d, 1/d parked in the p region.
    {
      double elast = 0.0;
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
        const double ein = __shfl_up_sync(full, elast, 1, NCH);
        if (k == c) {
          double eprev = k == 0 ? 0.0 : ein;   // wW of the first row of a column is 0
#pragma unroll
          for (int m = 0; m < R; ++m) {
            const int q = m * NT + tid, lq = lq0 + m * lqs;
            const double wW = l_s[lq], wE = e_s[lq];
            const double dinv = sy2d_div(1.0, 1.0 - wW * eprev);
            eprev = wE * dinv;
            l_s[lq] = wW * dinv; p_s[q] = dinv; e_s[lq] = eprev;
          }
          elast = eprev;
        }
      }
    }
    if (TML) {   // the thread's own factors (its own slots of l_s / e_s: no barrier needed) into tensor memory
      double t8[8], t2[2];
#pragma unroll
      for (int m = 0; m < 8; ++m) t8[m] = l_s[lq0 + m * lqs];
      t2[0] = l_s[lq0 + 8 * lqs]; t2[1] = l_s[lq0 + 9 * lqs];
      tmem_st<8>(tcol + 6 * R, t8);
      tmem_st<2>(tcol + 6 * R + 16, t2);
#pragma unroll
      for (int m = 0; m < 8; ++m) t8[m] = e_s[lq0 + m * lqs];
      t2[0] = e_s[lq0 + 8 * lqs]; t2[1] = e_s[lq0 + 9 * lqs];
      tmem_st<8>(tcol + 8 * R, t8);
      tmem_st<2>(tcol + 8 * R + 16, t2);
    }
    // pivot scaling of the rest of the row (thread-private slots: no barrier needed), r0 = rhs', rho0, max|r0|
#pragma unroll
    for (int m = 0; m < R; ++m) {
      const int q = m * NT + tid, gq = gq0 + m * gqs;
      const double dinv = p_s[q];
      if (TMW) { z[m] = wS_n[gq] * dinv; vr[VREG ? m : 0] = wN_n[gq] * dinv; }   // (z and v are free here) -> tensor memory below
      else { wS_g[q] = wS_n[gq] * dinv; wN_g[q] = wN_n[gq] * dinv; }
      const double rh = rhs_n[gq] * dinv;
      rhs_g[q] = rh;
      rs[m] = rh;
      if (RH16) {
        const unsigned h16 = (unsigned)__double2hiint(rh) >> 16;
        if (RHREG) rhp[RHREG ? m >> 1 : 0] = (m & 1) ? (rhp[RHREG ? m >> 1 : 0] | (h16 << 16)) : h16;
        else rh16[q] = (unsigned short)h16;
        acc[0] += rhat_of(q, m) * rh;
      } else {
        acc[0] += rh * rh;
      }
      acc[1] = nmax(acc[1], fabs(rh));
    }
    if (TMW) {
#pragma unroll
      for (int m0 = 0; m0 < 8; m0 += 4) {
        const double t[8] = {z[m0], vr[VREG ? m0 : 0], z[m0 + 1], vr[VREG ? m0 + 1 : 0], z[m0 + 2], vr[VREG ? m0 + 2 : 0], z[m0 + 3], vr[VREG ? m0 + 3 : 0]};
        tmem_st<8>(tcol + 4 * m0, t);
      }
      const double t[4] = {z[8], vr[VREG ? 8 : 0], z[9], vr[VREG ? 9 : 0]};
      tmem_st<4>(tcol + 32, t);
      tmem_wait_st();
    }
    cta_reduce_x<2>(acc, 1, red);
    double rho = acc[0];
    rmax = acc[1];
    double alpha = 1.0, omega = 1.0, beta = 0.0;
    bool first = true;
    it = 0;
    state = (rmax <= a.tol) ? 1 : 0;

    // One Thomas solve of the owned rows, z <- T^-1 b (b in registers), as a partitioned solve:
    // every lane runs the recurrence over its R rows with carry-in 0 and tracks the product of
    // the multipliers, i.e. its chunk as an affine map  carry_out = A + B * carry_in.  The maps
    // of the NCH lanes of a column are composed with a 3-step shuffle scan, and a second pass adds
    // (product up to row m) * carry_in.  All lanes work concurrently - no serial hand-off.
    auto tsolve = [&](auto bget, auto after_forward) {   // bget(m): right-hand side of owned row m (called once per row, in order); after_forward(): once all rows were fetched
      double A = 0.0, B = 1.0;
      double fa[TML ? 8 : 1], fb[TML ? 2 : 1];   // the factors of the sweep in flight (TML)
      if (TML) tmem_ld2(fa, tcol + 6 * R, fb, tcol + 6 * R + 16);
      auto lfac = [&](int m) { return TML ? (m < 8 ? fa[TML ? m : 0] : fb[TML ? m - 8 : 0]) : l_s[lq0 + m * lqs]; };
      auto efac = [&](int m) { return TML ? (m < 8 ? fa[TML ? m : 0] : fb[TML ? m - 8 : 0]) : e_s[lq0 + m * lqs]; };
#pragma unroll
      for (int m = 0; m < R; ++m) {
        const double lm = lfac(m);
        A = bget(m) - lm * A;
        z[m] = A;
        B = -lm * B;
      }
      after_forward();
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
        P = -lfac(m) * P;
        z[m] += P * cin;
      }
      // backward (unit diagonal): x_m = z_m - e_m x_{m+1}
      if (TML) tmem_ld2(fa, tcol + 8 * R, fb, tcol + 8 * R + 16);
      A = 0.0; B = 1.0;
#pragma unroll
      for (int m = R - 1; m >= 0; --m) {
        const double em = efac(m);
        A = z[m] - em * A;
        z[m] = A;
        B = -em * B;
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
        P = -efac(m) * P;
        z[m] += P * cin;
      }
    };
    // publish z into hat (natural layout) for the S/N neighbours
    auto publish = [&]() {
#pragma unroll
      for (int m = 0; m < R; ++m) {
        const int i = i0 + m;
        if (FULL || (col_ok && i < nx)) hat[i * hs + j] = z[m];
      }
      __syncthreads();
    };
    const int jS = j > 0 ? j - 1 : j, jN = j < ny - 1 ? j + 1 : j;

    while (state == 0) {
      // p = r + beta (p - omega v), formed row by row as the right-hand side of the first sweep
      if (PTM) {
        double p8[PTM ? 8 : 1], p2[PTM ? 2 : 1];
        if (!first) tmem_ld2(p8, tcol + 4 * R, p2, tcol + 4 * R + 16);
        tsolve([&](int m) {
          double& pm = m < 8 ? p8[PTM ? m : 0] : p2[PTM ? m - 8 : 0];
          pm = first ? rs[m] : rs[m] + beta * pm;
          return pm;
        }, [&]() {
          tmem_st_any(tcol + 4 * R, p8);
          tmem_st_any(tcol + 4 * R + 16, p2);
          tmem_wait_st();
        });
      } else {
        tsolve([&](int m) {
          const int q = m * NT + tid;
          const double pm = first ? rs[m] : (VREG ? rs[m] + beta * p_s[q] : rs[m] + beta * (p_s[q] - omega * v_g[q]));
          p_s[q] = pm;
          return pm;
        }, [] {});
      }
      publish();
      // v = p + wS phat_S + wN phat_N ; (rhat, v).  z is dead from here on (the thread re-reads its own
      // phat from hat), so v and later t reuse its registers: peak live arrays are rs, p and one more.
      double (&vv)[R] = z;
      double a1[1] = {0.0};
      if (PTM) {
        tm_rows_p([&](int m, double ws, double wn, double pm) {
          const int i = i0 + m, q = m * NT + tid;
          const double val = pm + (ws * hat[i * hs + jS] + wn * hat[i * hs + jN]);
          vr[VREG ? m : 0] = val;
          a1[0] += rhat_of(q, m) * val;
        });
      } else if (TMW) {
        tm_rows([&](int m, double ws, double wn) {
          const int i = i0 + m, q = m * NT + tid;
          const double val = p_s[q] + (ws * hat[i * hs + jS] + wn * hat[i * hs + jN]);
          vr[VREG ? m : 0] = val;
          a1[0] += rhat_of(q, m) * val;
        });
      } else {
#pragma unroll
        for (int m = 0; m < R; ++m) {
          const int i = i0 + m, q = m * NT + tid;
          double val = 0.0;
          if (FULL || (col_ok && i < nx)) val = p_s[q] + (wS_g[q] * hat[i * hs + jS] + wN_g[q] * hat[i * hs + jN]);
          if (VREG) vr[VREG ? m : 0] = val;
          else { vv[m] = val; v_g[q] = val; }
          a1[0] += rhat_of(q, m) * val;
        }
      }
      cta_reduce_x<1, false>(a1, 1, red);
      alpha = a1[0] != 0.0 ? rho / a1[0] : 0.0;
      // s = r - alpha v (in place)
#pragma unroll
      for (int m = 0; m < R; ++m) rs[m] -= alpha * (VREG ? vr[VREG ? m : 0] : vv[m]);
      tsolve([&](int m) { return rs[m]; }, [] {});
      publish();
      // t = s + wS shat_S + wN shat_N ; (t,s), (t,t)
      double a2[2] = {0.0, 0.0};
      if (TMW) {
        tm_rows([&](int m, double ws, double wn) {
          const int i = i0 + m;
          const double val = rs[m] + (ws * hat[i * hs + jS] + wn * hat[i * hs + jN]);
          vv[m] = val;  // t
          a2[0] += val * rs[m];
          a2[1] += val * val;
        });
      } else {
#pragma unroll
        for (int m = 0; m < R; ++m) {
          const int i = i0 + m, q = m * NT + tid;
          double val = 0.0;
          if (FULL || (col_ok && i < nx)) val = rs[m] + (wS_g[q] * hat[i * hs + jS] + wN_g[q] * hat[i * hs + jN]);
          vv[m] = val;  // t
          a2[0] += val * rs[m];
          a2[1] += val * val;
        }
      }
      cta_reduce_pair(a2[0], a2[1], red + 64);
      omega = a2[1] > 0.0 ? a2[0] / a2[1] : 0.0;
      // y += alpha p + omega s (x = T^-1 y is formed once, after the loop: x = sum alpha phat + omega shat and
      // T^-1 is linear) ; r = s - omega t ; (rhat, r), max|r|
      double a3[2] = {0.0, 0.0};
      auto yrow = [&](int m, double yold) {   // returns the new y_m
        const int q = m * NT + tid;
        const double pm = p_s[q];
        const double ynew = yold + (alpha * pm + omega * rs[m]);
        if (VREG) p_s[q] = pm - omega * vr[VREG ? m : 0];
        rs[m] -= omega * vv[m];
        a3[0] += rhat_of(q, m) * rs[m];
        a3[1] = nmax(a3[1], fabs(rs[m]));
        return ynew;
      };
      if (PTM) {   // p from / to tensor memory, y in the thread-private shared-memory slots
        double p8[PTM ? 8 : 1], p2[PTM ? 2 : 1];
        tmem_ld2(p8, tcol + 4 * R, p2, tcol + 4 * R + 16);
#pragma unroll
        for (int m = 0; m < R; ++m) {
          const int q = m * NT + tid;
          double& pm = m < 8 ? p8[PTM ? m : 0] : p2[PTM ? m - 8 : 0];
          p_s[q] = (first ? 0.0 : p_s[q]) + (alpha * pm + omega * rs[m]);   // y
          pm -= omega * vr[VREG ? m : 0];
          rs[m] -= omega * vv[m];
          a3[0] += rhat_of(q, m) * rs[m];
          a3[1] = nmax(a3[1], fabs(rs[m]));
        }
        tmem_st_any(tcol + 4 * R, p8);
        tmem_st_any(tcol + 4 * R + 16, p2);
        tmem_wait_st();
      } else if (TMW) {
        double y8[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, y2[2] = {0.0, 0.0};
        if (!first) { tmem_ld<8>(y8, tcol + 4 * R); tmem_ld<2>(y2, tcol + 4 * R + 16); }
#pragma unroll
        for (int m = 0; m < 8; ++m) y8[m] = yrow(m, y8[m]);
        tmem_st<8>(tcol + 4 * R, y8);
#pragma unroll
        for (int m = 0; m < 2; ++m) y2[m] = yrow(8 + m, y2[m]);
        tmem_st<2>(tcol + 4 * R + 16, y2);
        tmem_wait_st();
      } else {
#pragma unroll
        for (int m = 0; m < R; ++m) {
          const int q = m * NT + tid;
          y_g[q] = yrow(m, first ? 0.0 : y_g[q]);
        }
      }
      cta_reduce_sum_max(a3[0], a3[1], red + 128);
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
    // full tile: f^n and yprev of the owned cells are fetched now, into the registers of r and v (dead from here on), so
    // that their L2 latency hides behind the last line solve
    if (VREG) {
#pragma unroll
      for (int m = 0; m < R; ++m) {
        const int n = (i0 + m) * ny + j;
        rs[m] = f[n];
        vr[VREG ? m : 0] = yprev[n];
      }
    }
    // x = T^-1 y  (left in z)
    if (it > 0) {
      if (PTM) {
        tsolve([&](int m) { return p_s[m * NT + tid]; }, [] {});   // y sits in the p slots
      } else if (TMW) {   // y into z (the solve reads row m before it writes it)
        double y8[8], y2[2];
        tmem_ld<8>(y8, tcol + 4 * R);
        tmem_ld<2>(y2, tcol + 4 * R + 16);
#pragma unroll
        for (int m = 0; m < 8; ++m) z[m] = y8[m];
        z[8] = y2[0]; z[9] = y2[1];
        tsolve([&](int m) { return z[m]; }, [] {});
      } else {
        tsolve([&](int m) { return y_g[m * NT + tid]; }, [] {});
      }
    } else {
#pragma unroll
      for (int m = 0; m < R; ++m) z[m] = 0.0;
    }

    // ------------- f^{n+1} = c (1 + d) ; predictor ; true residual of the last step -------------
    const bool last = step == a.nsteps - 1;
    if (last) {
      // true residual with the FULL operator, row by row in the pivot-scaled form
      //   r'_i = rhs'_i - (l'_i x_W + (1/d_i) x_i + e_i x_E + wS'_i x_S + wN'_i x_N),   1/d_i = 1 + l'_i e_{i-1},
      // reported unscaled: |r_i| = d_i |r'_i| (the quantity the lockstep engine reports).  e_{i-1} lives in the
      // previous slot of this thread or in the last slot of lane k-1.
      __syncthreads();   // every thread is past its last read of shat
      publish();
      double mres = 0.0, mrel = 0.0;   // absolute, and componentwise-relative (k_true_residual) true residual
      auto resrow = [&](int m, double wS_m, double wN_m) {
        const int i = i0 + m, lq = lq0 + m * lqs, q = m * NT + tid;
        if (FULL || (col_ok && i < nx)) {
          const double lp = l_s[lq];
          double dinv_i = 1.0;
          if (i > 0) {
            const int kp = (i - 1) / R, mp = (i - 1) - kp * R;
            const int tp = tid - (k - kp);
            dinv_i = 1.0 + lp * e_s[FULL ? (i - 1) * LS + j : mp * NT + tp];
          }
          const double dW = i > 0 ? hat[(i - 1) * hs + j] : 0.0, dE = i < nx - 1 ? hat[(i + 1) * hs + j] : 0.0;
          const double tW = lp * dW, tE = e_s[lq] * dE, tS = wS_m * hat[i * hs + jS], tN = wN_m * hat[i * hs + jN];
          const double ax = dinv_i * z[m] + ((tW + tE) + (tS + tN));
          const double ra = fabs(rhs_g[q] - ax);
          mres = nmax(mres, ra / dinv_i);
          mrel = nmax(mrel, ra / (dinv_i * (1.0 + fabs(z[m])) + ((fabs(tW) + fabs(tE)) + (fabs(tS) + fabs(tN)))));
        }
      };
      if (TMW) {
        tm_rows(resrow);
      } else {
#pragma unroll
        for (int m = 0; m < R; ++m) resrow(m, wS_g[m * NT + tid], wN_g[m * NT + tid]);
      }
      double mm[2] = {mres, mrel};
      cta_reduce_x<2>(mm, 0, red);
      res_true = mm[0];
      res_rel = mm[1];
    }
    if (state >= 2) break;   // the solve stopped without converging (maxit, breakdown, NaN): f and yprev stay those of t^n
    double fneg = 0.0, fmin_neg = -1.0e300, fmin_l = 1.0e300;
    // full tile: the last ratios are fetched as one batch (inside the loop each load would wait for the stores of the row
    // before it - the compiler has to assume they alias)
    double yl[VREG ? R : 1];
    if (VREG && a.predictor >= 2) {
#pragma unroll
      for (int m = 0; m < R; ++m) yl[VREG ? m : 0] = a.ylast[base + (i0 + m) * ny + j];
    }
#pragma unroll
    for (int m = 0; m < R; ++m) {
      const int i = i0 + m;
      if (FULL || (col_ok && i < nx)) {
        const int n = i * ny + j;
        const double fold = VREG ? rs[m] : f[n];
        const double yp = VREG ? vr[VREG ? m : 0] : yprev[n];
        const double d = z[m];
        const double fnew = (fold * yp) * (1.0 + d);
        f[n] = fnew;
        if (a.predictor) {
          if (VREG) predictor_update(a.predictor, fnew, fold, yprev + n, a.ylast + base + n, yl[VREG ? m : 0]);
          else predictor_update(a.predictor, fnew, fold, yprev + n, a.ylast + base + n);
        }
        fneg += fnew < 0.0 ? 1.0 : 0.0;
        fmin_neg = nmax(fmin_neg, -fnew);
        fmin_l = ::fmin(fmin_l, fnew);
      }
    }
    if (last) {
      // sum of negatives; max of (-f) shifted to be non-negative for the zero-padded reduction
      double mm[2] = {fneg, fmin_neg + 1.0e300};   // (the shifted maximum only carries a NaN through; min f itself is reduced below)
      cta_reduce_x<2>(mm, 1, red);
      {   // min f of the problem: 1e300 + (-f) would round every |f| < 1e284 away
        const int lane_ = tid & 31, w_ = tid >> 5, nw_ = NT >> 5;
        double mn_ = warp_min(fmin_l);
        if (lane_ == 0) red[w_] = mn_;
        __syncthreads();
        fmin_l = warp_min(lane_ < nw_ ? red[lane_] : 1.0e300);
        __syncthreads();
      }
      if (tid == 0) {
        if (mm[0] > 0.0) atomicAdd(&a.stats->negatives, (unsigned long long)mm[0]);
        if (!(mm[1] == mm[1]) || !(res_rel == res_rel)) atomicAdd(&a.stats->n_bad, 1);   // non-finite f or residual
        const double mn = fmin_l;
        unsigned long long* addr = reinterpret_cast<unsigned long long*>(&a.stats->fmin);
        unsigned long long old = *addr;
        while (mn < __longlong_as_double((long long)old)) {
          const unsigned long long assumed = old;
          old = atomicCAS(addr, assumed, (unsigned long long)__double_as_longlong(mn));
          if (old == assumed) break;
        }
      }
    }
    if (FULL) __threadfence();   // the next assembly reads f, yprev through the async proxy (bulk copies)
    __syncthreads();  // f and yprev complete before the next step's assembly reads neighbours
    steps_ok = step + 1;
  }
  __syncthreads();   // (the failure path leaves the step loop without the barrier above)
  if (xa.hout && (state >= 2 || step_end == a.nsteps)) {   // the problem leaves the queue: its f goes to the host buffer
    double* dst = xa.hout + base;
    if ((N & 1) == 0) {
      for (int n = tid; n < N / 2; n += NT) reinterpret_cast<double2*>(dst)[n] = reinterpret_cast<const double2*>(f)[n];
    } else {
      for (int n = tid; n < N; n += NT) dst[n] = f[n];
    }
  }
  if (tid == 0) {
    Scal* sc = a.scal + prob;
    const int cost = (a.cost ? a.cost[prob] : 0) + it_total;
    if (a.cost) a.cost[prob] = cost;
    sc->it = it;
    sc->state = state;
    sc->rmax = rmax;
    if (state >= 2 || step_end == a.nsteps) atomicMax(&a.stats->it_max, it);   // iterations of the problem's LAST step of the call
    atomicAdd(&a.stats->it_sum_all, (unsigned long long)it_total);
    xa.steps_done[prob] = steps_ok;
    if (state >= 2) {   // failed: nothing of the failing step was committed, the problem leaves the queue
      atomicAdd(&a.stats->n_bad, 1);
      atomicMin(&a.stats->steps_min, steps_ok);
      atomicMax(&a.stats->it_total_max, cost);
      const int item_no = step_begin / xa.chunk;                     // chunks start at multiples of `chunk`
      atomicSub(&xa.q->total, xa.nchunks - 1 - item_no);              // its later items will never be pushed
    } else if (step_end < a.nsteps) {
      const int u = atomicAdd(&xa.q->tail, 1);
      st_release_gpu(xa.slots + u, prob);   // f, yprev, steps_done of this problem are visible to whoever pops it
    } else {
      atomicMin(&a.stats->steps_min, steps_ok);
      atomicMax(&a.stats->it_total_max, cost);
      atomicMax(reinterpret_cast<unsigned long long*>(&a.stats->resid_max), (unsigned long long)__double_as_longlong(res_true));
      atomicMax(reinterpret_cast<unsigned long long*>(&a.stats->resid_rel_max), (unsigned long long)__double_as_longlong(res_rel));
    }
  }
 }
  if (TMW) {
    tmem_fence_before_sync();
    __syncthreads();
    if (w == 0) tmem_dealloc_all(reinterpret_cast<volatile unsigned*>(s_item + 1)[0]);
  }
}

}  // namespace sy2d
