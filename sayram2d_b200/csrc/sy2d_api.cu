// C-ABI implementation (include/sayram2d.h) of the B200-native Sayram-2D engine.
// Host orchestration only: staging, launch order, convergence polling, CUDA
// graphs.  All arithmetic of the hot path lives in sy2d_kernels.cuh.
#include "../../include/sayram2d.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <string>
#include <vector>

#include "sy2d_geometry.h"
#include "sy2d_kernels.cuh"
#include "sy2d_problem_kernel.cuh"
#include "sy2d_xline_kernel.cuh"
#include "sy2d_xline_cluster.cuh"
#include "sy2d_xline_lockstep.cuh"
#include "sy2d_mg.cuh"
#include "sy2d_assemble_tma.cuh"
#include "sy2d_assemble_march.cuh"
#include "sy2d_assemble_col.cuh"
#include "sy2d_assemble_wide.cuh"
#include "sy2d_peaks.cuh"

using namespace sy2d;

namespace {
thread_local std::string g_create_error;

}  // namespace

namespace { struct SlabTransport; }

struct sy2d_ctx {
  int device = 0, nx = 0, ny = 0, nbatch = 0;
  size_t N = 0, total = 0;
  double dt = 0.0;
  long long istep = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  // grid caps of the grid-stride kernels (tuning: SY2D_CTAS_PER_SM, SY2D_ASM_CTAS_PER_SM).  6 = the CTAs of 256 threads x 40
  // registers that are resident on an SM at once: one full wave, no tail (profiles/sweep_caps_steps.py: 1024^2 4.98 -> 4.81 ms
  // per step against 16, 2048^2 15.2 -> 14.9, 4096^2 equal; 8 is worse everywhere - a partial second wave)
  int ctas_per_sm = 6, asm_ctas_per_sm = 4, march_ctas_per_sm = 5, wide_ctas_per_sm = 2;
  int mg_fuse = 1;            // p / s updates formed inside the first line solve of their V-cycle (SY2D_MG_FUSE=0: kernels of their own)
  int mg_fuse_pending = 0;    // 1 / 2: the next MODE-0 line launch forms p / s (set by launch_iteration_mg, consumed by mg_line_shape)
  int mg_line_pre = 1;   // line kernel with prefetched backward factors / old iterate and block-parallel scans (SY2D_MG_LINE_PRE=0: the three-phase kernel)
  int asm_kernel = 0;   // default fast assembly: 0 TMA-staged tiles (strided tile order), 1 warp-marching, 2 TMA-staged column runs (experiment, slower); SY2D_ASM_KERNEL = tma | march | col
  double col_edge_weight = 1.5;   // cost of a boundary tile relative to an interior one when the column runs are cut (SY2D_COL_EDGE_WEIGHT)
  int* d_col_runs = nullptr;      // k_assemble_col: first tile of every CTA's run (strip-major order)
  int col_runs_key[3] = {0, 0, 0};   // tiles_i, tiles_j, CTAs the device array was built for
  int pipe_max = 32, pipe_forced = 0;   // sy2d_step_host: at most pipe_max pipelined sub-batches (SY2D_PIPE_CHUNKS forces a count)
  int deterministic = 0;     // SY2D_DETERMINISTIC=1: cross-CTA sums in slot order instead of floating-point atomics (bitwise reproducible runs)
  double* part = nullptr;    // slots of the deterministic cross-CTA sums (sy2d_kernels.cuh, cta_totals)
  size_t part_stride = 0;
  std::string err;
  sy2d_options opt;
  bool have_coeffs = false, have_bc = false, have_f = false;

  // geometry
  std::vector<double> h_xe, h_ye;
  double *d_wxL = nullptr, *d_wxR = nullptr, *d_wyB = nullptr, *d_wyT = nullptr, *d_dx = nullptr, *d_dy = nullptr;
  double *d_bc[4] = {nullptr, nullptr, nullptr, nullptr};
  int bc[4] = {SY2D_ZEROFLUX, SY2D_ZEROFLUX, SY2D_ZEROFLUX, SY2D_ZEROFLUX};
  // coefficients
  double *tx = nullptr, *ty = nullptr, *cxy = nullptr, *U = nullptr, *Ud = nullptr;
  // state
  double *f = nullptr, *yprev = nullptr, *ylast = nullptr, *cs = nullptr;
  // operator + Krylov vectors
  double *wW = nullptr, *wE = nullptr, *wS = nullptr, *wN = nullptr, *rhs = nullptr;
  double *x = nullptr, *r = nullptr, *p = nullptr, *v = nullptr, *s = nullptr, *t = nullptr;
  double *xl_l = nullptr, *xl_dinv = nullptr, *xl_e = nullptr, *xl_hat = nullptr;  // engine 1 x-line: LU factors, hat vector
  // engine 1 multigrid preconditioner (sy2d_mg.cuh): level 0 reuses wW..wN, xl_l/xl_dinv/xl_e (line LU) and
  // xl_hat (phat); everything else lives in mg_bufs
  int mg_nlev = 0, mg_seg = 16;
  MgLevels mg;
  double* mg_rc[kMgMaxLevels] = {};   // writable right-hand sides of the coarse levels
  double *mg_om0 = nullptr, *mg_shat = nullptr;
  unsigned* mg_tail_ctr = nullptr;      // [nbatch] barrier counters of the fused coarse-tail kernel (zero between launches)
  int mg_cluster = 1;                   // SY2D_MG_CLUSTER=0: long columns in one CTA (fewer columns per CTA) instead of a CTA cluster
  int mg_tail_ny = 0;                   // SY2D_MG_TAIL_NY = n: levels with at most n columns run inside the fused k_mg_tail kernel.  Off by
                                        // default: at 1024^2 it halves the launches of a step (2878 -> 1558 in 5 steps) and changes nothing
                                        // (4.62 -> 4.68 ms per step) - a stage costs its dependent chain (loads, sweep, scan, sweep, scan,
                                        // store: ~7 us), not its launch, whether it is a kernel or a stage behind a barrier
  // row-slab mode, exact lines across ranks (spike correction, sy2d_mg.cuh): per level the damped spikes of the
  // own rows and the gathered spike tips [nranks][4][ny_l]; per solve the tips [2][ny] / gathered [nranks][2][ny] / coefficients
  double* mg_spW[kMgMaxLevels] = {};
  double* mg_spV[kMgMaxLevels] = {};
  double* mg_sp_all[kMgMaxLevels] = {};
  double *mg_tips = nullptr, *mg_tips_all = nullptr, *mg_coef = nullptr;
  bool mg_spike = false;   // spikes of the current operator are set up: line solves are corrected
  std::vector<double*> mg_bufs;
  // TMA-staged assembly (sy2d_assemble_tma.cuh): tensor maps of f, yprev, tx, ty, cxy, U, Ud
  bool have_tma = false;
  int last_asm_kernel = 0;   // which engine-1 assembly kernel the last launch_assembly used (sy2d_last_assembly_kernel; tests)
  bool have_tma_wide = true;   // cleared when the 68 x 10 / 64 x 8 boxes cannot be encoded (ny < 64)
  AsmMaps tma_maps;
  AsmMaps* d_tma_maps = nullptr;   // device copy read by the TMA unit
  // Asynchronous staging (sy2d_set_coeffs_async / sy2d_set_bc_async): a second set of coefficient arrays and boundary
  // lines is filled on `copy_stream` from pinned staging while the step in flight computes with the active set; the
  // sets are swapped at the start of the next time step (stage_mu guards the hand-over: the setters may be called
  // from another host thread while sy2d_step runs).
  std::mutex stage_mu;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t coeffs_ready = nullptr, bc_ready = nullptr;
  bool coeffs_pending = false, bc_pending = false;
  double *tx2 = nullptr, *ty2 = nullptr, *cxy2 = nullptr, *U2 = nullptr, *Ud2 = nullptr;   // the inactive set
  double* raw_dev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // G, Dxx, Dxy, Dyy, inv_tau as uploaded
  double* raw_pin[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // pinned host staging
  AsmMaps* d_tma_maps2 = nullptr;   // tensor maps of the inactive set
  double* d_bc2[4] = {nullptr, nullptr, nullptr, nullptr};
  double* bc_pin = nullptr;         // pinned: 2 (ny + 1) + 2 (nx + 1)
  int bc2[4] = {SY2D_ZEROFLUX, SY2D_ZEROFLUX, SY2D_ZEROFLUX, SY2D_ZEROFLUX};
  long long swaps = 0;              // buffer-set swaps done so far (tests)
  long long steps_begun = 0;        // sy2d_step / sy2d_step_host calls that have passed their swap-in point (guarded by stage_mu)
  Scal* scal = nullptr;
  int* d_nactive = nullptr;
  int* h_nactive = nullptr;  // pinned
  StepStats* d_stats = nullptr;
  StepStats* h_stats = nullptr;  // pinned
  // engine 2 / x-line: per-problem scratch in thread-private layout, allocated on first use
  // row-slab mode (sy2d_create_slab): this context holds rows [i_lo, i_hi) of an nx_glob x ny grid in
  // local arrays of nx = (i_hi - i_lo) + 2 rows (one halo row on each side)
  bool slab = false;
  int rank = 0, nranks = 1, nx_glob = 0, i_lo = 0, i_hi = 0;
  SlabTransport* tp = nullptr;   // NCCL (one process per GPU) or the in-process transport (tests, single-GPU boxes)
  double* d_gather = nullptr;    // [nranks][5]
  int* d_order = nullptr;       // engine 2 scheduling: problems sorted by last call's cost, longest first
  int* d_cost = nullptr;
  std::vector<int> h_cost, h_order;
  int order_age = 0, order_calls = 0, order_C = -1;   // calls since the issue order was last rebuilt from the measured costs; sub-batch count it was built for
  size_t xl_scratch_slots = 0;
  // engine 2 work queues (sy2d_xline_kernel.cuh): control words per sub-batch launch, ticket slots, steps done per problem
  XlineQueue* d_qctl = nullptr;
  int* d_slots = nullptr;
  size_t slots_cap = 0;
  int* d_steps_done = nullptr;
  int host_io_direct = 0;        // sy2d_step_host, SY2D_HOST_IO=direct: the x-line kernel reads / writes pinned host buffers itself instead of
                                 // the copy engines (measured SLOWER on B200 / PCIe: 9.9 against 8.5 ms per end-to-end step at 4096
                                 // members - the loads an SM issues to system memory run at ~1 GB/s per SM; kept as an experiment)
  int xl_cluster = 0;            // SY2D_XLINE_CLUSTER=1: 80 x 80 problems on pairs of CTAs (sy2d_xline_cluster.cuh), all iteration state on chip
  int xl_chunk = 1;              // time steps per work item (SY2D_XLINE_CHUNK; 0 = all steps of a call: one CTA per problem)
  std::vector<cudaStream_t> pipe_streams;   // sy2d_step_host: one stream per sub-batch
  std::vector<cudaEvent_t> pipe_events;
  cudaEvent_t pipe_start = nullptr;
  double* xl_scratch = nullptr;
  int xl_R = 0, xl_NT = 0, xl_S = 0;
  size_t xl_smem = 0;
  // iteration-chunk graph
  cudaGraphExec_t chunk_exec = nullptr;
  cudaGraphExec_t one_exec = nullptr;   // multigrid: a single iteration + the convergence poll
  cudaGraphExec_t slab_exec = nullptr;  // slab mode over NCCL: one multigrid-preconditioned iteration (kernels AND collectives) + the poll
  int slab_graph = 1;                   // SY2D_SLAB_GRAPH=0: issue the slab iteration call by call
  int mg_last_iters = 0;                // iterations of the previous time step (issue plan of the next one)
  bool mg_off = false;                  // multigrid failed on the current step: it is being redone with the x-line iteration
  int chunk_iters = 0;
  int chunk_variant = -1;
  // profiling
  bool profiling = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
  std::vector<int> ev_class;
  size_t ev_used = 0;
  sy2d_profile prof;
  std::vector<double> ev_cells;
  double cur_cells = 0.0;       // cells of active problems for the launches being recorded
  long long launches = 0;       // kernels launched since the start of the current sy2d_step
  cudaEvent_t ev_call0 = nullptr, ev_call1 = nullptr;
};

namespace {

int fail(sy2d_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf; else g_create_error = buf;
  return code;
}

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(c, SY2D_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

Geometry geometry(const sy2d_ctx* c) {
  Geometry g;
  g.wxL = c->d_wxL; g.wxR = c->d_wxR; g.wyB = c->d_wyB; g.wyT = c->d_wyT;
  g.bc_xmin = c->d_bc[0]; g.bc_xmax = c->d_bc[1]; g.bc_ymin = c->d_bc[2]; g.bc_ymax = c->d_bc[3];
  for (int k = 0; k < 4; ++k) g.bc[k] = c->bc[k];
  g.nx = c->slab ? c->nx_glob : c->nx; g.ny = c->ny;
  return g;
}

KrylovVecs krylov(const sy2d_ctx* c) {
  KrylovVecs k;
  k.wW = c->wW; k.wE = c->wE; k.wS = c->wS; k.wN = c->wN; k.rhs = c->rhs;
  k.x = c->x; k.r = c->r; k.p = c->p; k.v = c->v; k.s = c->s; k.t = c->t;
  k.scal = c->scal; k.part = c->deterministic ? c->part : nullptr; k.part_stride = c->part_stride; k.n_active = c->d_nactive; k.tol = c->opt.tol; k.maxit = c->opt.maxit;
  k.freeze_state = 0;
  k.n_begin = 0; k.n_end = c->N; k.defer = 0;
  if (c->slab) { k.n_begin = (size_t)c->ny; k.n_end = (size_t)(c->nx - 1) * c->ny; k.defer = 1; }
  return k;
}

dim3 grid_of(const sy2d_ctx* c) { return dim3((unsigned)((c->N + kBlock - 1) / kBlock), (unsigned)c->nbatch, 1); }

// Grid-stride kernels: one resident wave of CTAs over the whole batch (6 per SM), so that a problem
// costs a few hundred block-level atomics per reduction instead of one per 256 cells.
unsigned capped_blocks(const sy2d_ctx* c, size_t work_items_per_problem, int threads) {
  const size_t need = (work_items_per_problem + threads - 1) / threads;
  const size_t cap = std::max<size_t>(1, (size_t)c->sm_count * c->ctas_per_sm / (size_t)c->nbatch);
  return (unsigned)std::min(need, cap);
}

// Multigrid (sy2d_mg.cuh) needs whole columns inside one CTA of the line kernel (nx <= 64 segments of 16 or
// 32 rows), pairs of columns on every level (ny a multiple of 4 gives at least two levels) and a single GPU.
// rows the multigrid kernels work on: the owned rows of a slab context (the line kernel solves them locally, the spike
// correction couples them to the neighbour ranks' rows)
int mg_rows(const sy2d_ctx* c) { return c->slab ? c->nx - 2 : c->nx; }

int mg_level_count(const sy2d_ctx* c) {
  if (mg_rows(c) > 8192 || mg_rows(c) < 8 || c->ny % 4 != 0 || c->ny < 16) return 0;
  // default: coarsen until a level has at most 64 columns (measured optimum from 128^2 to 2048^2: fewer levels cost
  // iterations, more levels cost latency-bound launches), at least two levels; mg_levels > 0 caps the count instead.
  // A level is only halved while its ny is a multiple of 4, so the coarsest level keeps an even ny >= 8.
  const int cap = c->opt.mg_levels > 0 ? std::min(c->opt.mg_levels, kMgMaxLevels) : kMgMaxLevels;
  int nlev = 1, ny = c->ny;
  while (nlev < cap && ny % 4 == 0 && ny >= 16 && (c->opt.mg_levels > 0 || ny > 64 || nlev < 2)) { ny /= 2; ++nlev; }
  return nlev >= 2 ? nlev : 0;
}
int mg_coarse_sweeps(const sy2d_ctx* c) { return c->opt.mg_coarse_sweeps > 0 ? c->opt.mg_coarse_sweeps : kMgCoarseSweeps; }
bool lockstep_mg(const sy2d_ctx* c) {
  if (c->mg_off) return false;
  if (c->opt.precond != SY2D_PRECOND_AUTO && c->opt.precond != SY2D_PRECOND_MG) return false;
  return mg_level_count(c) > 0;
}
// Engine 1 falls back to the segmented x-line preconditioner (needs at least one full segment)
bool lockstep_xline(const sy2d_ctx* c) {
  const int rows = c->slab ? c->nx - 2 : c->nx;
  if (lockstep_mg(c)) return false;
  return c->opt.precond != SY2D_PRECOND_JACOBI && c->opt.precond != SY2D_PRECOND_MG && rows >= kSeg;
}

XlVecs xl_vecs(const sy2d_ctx* c);

// RAII-less event bracket used only in profiling mode
struct Prof {
  sy2d_ctx* c;
  int idx = -1;
  Prof(sy2d_ctx* c_, int klass) : c(c_) {
    if (!c->profiling) return;
    if (c->ev_used == c->ev_pool.size()) {
      cudaEvent_t a, b;
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      c->ev_pool.emplace_back(a, b);
      c->ev_class.push_back(klass);
      c->ev_cells.push_back(0.0);
    }
    idx = (int)c->ev_used++;
    c->ev_class[idx] = klass;
    c->ev_cells[idx] = c->cur_cells;
    cudaEventRecord(c->ev_pool[idx].first, c->stream);
  }
  ~Prof() {
    if (idx >= 0) cudaEventRecord(c->ev_pool[idx].second, c->stream);
  }
};

XlVecs xl_vecs(const sy2d_ctx* c) {
  XlVecs x;
  x.k = krylov(c);
  x.l = c->xl_l; x.dinv = c->xl_dinv; x.e = c->xl_e; x.hat = c->xl_hat;
  x.ny = c->ny;
  x.row0 = c->slab ? 1 : 0;
  x.nrows = c->slab ? c->nx - 2 : c->nx;
  return x;
}

int xl_alloc(sy2d_ctx* c) {
  if (c->xl_hat) return SY2D_OK;
  double** arrs[4] = {&c->xl_l, &c->xl_dinv, &c->xl_e, &c->xl_hat};
  for (double** a : arrs) {
    CU(cudaMalloc(reinterpret_cast<void**>(a), c->total * sizeof(double)));
    CU(cudaMemsetAsync(*a, 0, c->total * sizeof(double), c->stream));
  }
  return SY2D_OK;
}

void launch_iteration_xline(sy2d_ctx* c) {
  const XlVecs x = xl_vecs(c);
  const int nseg = (x.nrows + kSeg - 1) / kSeg;
  const dim3 gs(capped_blocks(c, (size_t)nseg * c->ny, kSweepThreads), (unsigned)c->nbatch, 1);
  const bool v2 = c->ny % 2 == 0;
  const dim3 gc(capped_blocks(c, (x.k.n_end - x.k.n_begin) / (v2 ? 2 : 1), kBlock), (unsigned)c->nbatch, 1);
  { Prof p(c, SY2D_K_P_UPDATE); k_xl_sweep<0><<<gs, kSweepThreads, 0, c->stream>>>(x, c->N); }
  { Prof p(c, SY2D_K_SPMV_V); if (v2) k_xl_spmv_v<2><<<gc, kBlock, 0, c->stream>>>(x, c->N); else k_xl_spmv_v<1><<<gc, kBlock, 0, c->stream>>>(x, c->N); }
  { Prof p(c, SY2D_K_S_UPDATE); k_xl_sweep<1><<<gs, kSweepThreads, 0, c->stream>>>(x, c->N); }
  { Prof p(c, SY2D_K_SPMV_T); if (v2) k_xl_spmv_t<2><<<gc, kBlock, 0, c->stream>>>(x, c->N); else k_xl_spmv_t<1><<<gc, kBlock, 0, c->stream>>>(x, c->N); }
  { Prof p(c, SY2D_K_XR_UPDATE); if (v2) k_xl_xr<2><<<gc, kBlock, 0, c->stream>>>(x, c->N); else k_xl_xr<1><<<gc, kBlock, 0, c->stream>>>(x, c->N); }
}


// Tensor maps for the TMA-staged assembly: [nbatch][local rows][ny] fp64 arrays seen as 3-D tensors (ny fastest),
// halo boxes of 36 x 10 cells and interior boxes of 32 x 8 cells, out-of-bounds elements zero-filled.
// cuTensorMapEncodeTiled is a driver entry point; it is resolved through the runtime, so the library
// keeps linking against libcudart only.
bool tma_encode_maps(sy2d_ctx* c, AsmMaps* host_maps, double* tx, double* ty, double* cxy, double* U, double* Ud);

bool tma_build_maps(sy2d_ctx* c) {
  if (c->ny % 2 != 0 || mg_rows(c) < 2 * kTI || c->ny < kTJ) return false;
  if (!tma_encode_maps(c, &c->tma_maps, c->tx, c->ty, c->cxy, c->U, c->Ud)) return false;
  if (cudaFuncSetAttribute(k_assemble_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTmaSmemBytes) != cudaSuccess ||
      cudaFuncSetAttribute(k_assemble_col, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kColSmemBytes) != cudaSuccess ||
      cudaFuncSetAttribute(k_assemble_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWideSmemBytes) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&c->d_tma_maps), sizeof(AsmMaps)) != cudaSuccess ||
      cudaMemcpy(c->d_tma_maps, &c->tma_maps, sizeof(AsmMaps), cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return true;
}

bool tma_encode_maps(sy2d_ctx* c, AsmMaps* host_maps, double* tx, double* ty, double* cxy, double* U, double* Ud) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
      qres != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    return false;
  }
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  EncodeFn encode = reinterpret_cast<EncodeFn>(fn);
  double* arrs[7] = {c->f, c->yprev, tx, ty, cxy, U, Ud};
  const cuuint64_t dims[3] = {(cuuint64_t)c->ny, (cuuint64_t)c->nx, (cuuint64_t)c->nbatch};
  const cuuint64_t strides[2] = {(cuuint64_t)c->ny * sizeof(double), (cuuint64_t)c->N * sizeof(double)};
  const cuuint32_t estr[3] = {1, 1, 1};
  if (c->ny < kWJ) c->have_tma_wide = false;
  for (int k = 0; k < (c->have_tma_wide ? 14 : 7); ++k) {
    const bool wide = k >= 7, halo = k % 7 < 5;
    const cuuint32_t box[3] = {(cuuint32_t)(halo ? (wide ? kWHaloJ : kTmaHaloJ) : (wide ? kWJ : kTJ)), (cuuint32_t)(halo ? kTmaHaloI : kTI), 1};
    const CUresult r = encode(&host_maps->m[k], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, arrs[k % 7], dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,  // (L2 promotion raises an illegal-instruction fault with these boxes on B200 / driver 580: profiles/tma_probe.cu)
                             
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      if (!wide) return false;
      c->have_tma_wide = false;   // the wide boxes are only needed by the two-cells-per-thread experiment: the narrow maps stay usable
      break;
    }
  }
  return true;
}

// Warp-marching assembly of local rows [li_begin, li_end) (sy2d_assemble_march.cuh): one warp per 16 x 30 strip.
void launch_march(sy2d_ctx* c, const Geometry& geo, const AssembleOut& o, int gi0, int li_begin, int li_end, int defer) {
  const int strips_i = (li_end - li_begin + kMarchRows - 1) / kMarchRows, strips_j = (c->ny + kMarchCols - 1) / kMarchCols;
  const int nstrips = strips_i * strips_j;
  const size_t need = ((size_t)nstrips + kMarchWarps - 1) / kMarchWarps;
  const size_t cap = std::max<size_t>(1, (size_t)c->sm_count * c->march_ctas_per_sm / (size_t)c->nbatch);
  k_assemble_march<<<dim3((unsigned)std::min(need, cap), (unsigned)c->nbatch, 1), kMarchWarps * 32, kMarchSmemBytes, c->stream>>>(
      c->f, c->yprev, c->tx, c->ty, c->cxy, c->U, c->Ud, geo, o, strips_j, nstrips, gi0, li_begin, li_end, defer);
}

// Column-run assembly (sy2d_assemble_col.cuh): the tiles in strip-major order are cut into one contiguous run per CTA so that
// every run has the same COST; a tile on the boundary of the grid (predicated path, Dirichlet faces) counts col_edge_weight.
bool col_runs(sy2d_ctx* c, int tiles_i, int tiles_j, int ctas) {
  if (c->d_col_runs && c->col_runs_key[0] == tiles_i && c->col_runs_key[1] == tiles_j && c->col_runs_key[2] == ctas) return true;
  const int ntiles = tiles_i * tiles_j;
  std::vector<double> cost((size_t)ntiles + 1, 0.0);
  for (int t = 0; t < ntiles; ++t) {
    const int tj = t / tiles_i, ti = t - tj * tiles_i;
    const bool edge = ti == 0 || ti == tiles_i - 1 || tj == 0 || tj == tiles_j - 1;
    cost[(size_t)t + 1] = cost[t] + (edge ? c->col_edge_weight : 1.0);
  }
  std::vector<int> start((size_t)ctas + 1, ntiles);
  start[0] = 0;
  int t = 0;
  for (int k = 1; k < ctas; ++k) {
    const double target = cost[ntiles] * k / ctas;
    while (t < ntiles && cost[(size_t)t + 1] <= target + 1e-9) ++t;   // the tile that crosses the target goes to the later run
    if (t < ntiles && cost[(size_t)t + 1] - target < target - cost[t]) ++t;   // ... unless most of it lies before the target
    start[k] = std::max(t, start[(size_t)k - 1]);
  }
  if (c->d_col_runs) { cudaFree(c->d_col_runs); c->d_col_runs = nullptr; }
  if (cudaMalloc(reinterpret_cast<void**>(&c->d_col_runs), start.size() * sizeof(int)) != cudaSuccess ||
      cudaMemcpy(c->d_col_runs, start.data(), start.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaGetLastError();
    c->d_col_runs = nullptr;
    return false;
  }
  c->col_runs_key[0] = tiles_i; c->col_runs_key[1] = tiles_j; c->col_runs_key[2] = ctas;
  return true;
}

bool launch_col(sy2d_ctx* c, const Geometry& geo, const AssembleOut& o, int tiles_i, int tiles_j, int gi0, int li_begin, int li_end, int defer) {
  const int ntiles = tiles_i * tiles_j;   // per problem; every problem of a batch is cut into the same runs (blockIdx.y = problem)
  const int ctas = (int)std::min<size_t>((size_t)ntiles, std::max<size_t>(1, (size_t)c->sm_count * c->asm_ctas_per_sm / (size_t)c->nbatch));
  if (!col_runs(c, tiles_i, tiles_j, ctas)) return false;
  k_assemble_col<<<dim3((unsigned)ctas, (unsigned)c->nbatch, 1), kColThreads, kColSmemBytes, c->stream>>>(c->d_tma_maps, geo, o, c->d_col_runs, tiles_i, gi0, li_begin, li_end, defer);
  return true;
}

// Two cells per thread on 8 x 64 tiles (sy2d_assemble_wide.cuh); two CTAs per SM.
bool launch_wide(sy2d_ctx* c, const Geometry& geo, const AssembleOut& o, int tiles_i, int gi0, int li_begin, int li_end, int defer) {
  if (c->ny < kWJ || !c->have_tma_wide) return false;
  const int tiles_j = (c->ny + kWJ - 1) / kWJ, ntiles = tiles_i * tiles_j;
  const unsigned ctas = (unsigned)std::min<size_t>((size_t)ntiles, std::max<size_t>(1, (size_t)c->sm_count * c->wide_ctas_per_sm / (size_t)c->nbatch));
  k_assemble_wide<<<dim3(ctas, (unsigned)c->nbatch, 1), kTI * kTJ, kWideSmemBytes, c->stream>>>(c->d_tma_maps, geo, o, tiles_j, ntiles, gi0, li_begin, li_end, defer);
  return true;
}

// Engine-1 assembly of the whole batch.  variant: 0 auto (SY2D_ASM_KERNEL = march | tma picks the default fast kernel),
// 1 per-cell kernel, 2 tiled kernel without TMA, 3 warp-marching kernel, 4 TMA-staged tiles (tests compare them).
void launch_assembly(sy2d_ctx* c, const Geometry& geo, const AssembleOut& o, int variant) {
  const int tiles_i = (c->nx + kTI - 1) / kTI, tiles_j = (c->ny + kTJ - 1) / kTJ;
  const bool tiled = c->nx >= 2 * kTI && c->ny >= kTJ && variant != 1;
  // default: TMA-staged tiles (the fastest of the three at every size measured: 34 / 101 / 364 us at 1024^2 / 2048^2 /
  // 4096^2 against 37 / 104 / 410 us for the marching kernel and 55 / 196 / 767 us for the plain-load tiles); where no
  // tensor map exists (odd ny) the marching kernel, which has no alignment requirement
  if (tiled && (variant == 3 || (variant == 0 && (c->asm_kernel == 1 || !c->have_tma)))) {
    launch_march(c, geo, o, 0, 0, c->nx, 0);
    c->last_asm_kernel = 3;
  } else if (tiled && c->have_tma && (variant == 6 || (variant == 0 && c->asm_kernel == 3)) && launch_wide(c, geo, o, tiles_i, 0, 0, c->nx, 0)) {
    c->last_asm_kernel = 6;
  } else if (tiled && c->have_tma && (variant == 5 || (variant == 0 && c->asm_kernel == 2)) && launch_col(c, geo, o, tiles_i, tiles_j, 0, 0, c->nx, 0)) {
    c->last_asm_kernel = 5;
  } else if (tiled && c->have_tma && (variant == 0 || variant == 4 || variant == 5 || variant == 6)) {
    c->last_asm_kernel = 4;
    const int ntiles = tiles_i * tiles_j;
    const unsigned ctas = (unsigned)std::min<size_t>((size_t)ntiles, std::max<size_t>(1, (size_t)c->sm_count * c->asm_ctas_per_sm / (size_t)c->nbatch));
    k_assemble_tma<<<dim3(ctas, (unsigned)c->nbatch, 1), kTI * kTJ, kTmaSmemBytes, c->stream>>>(c->d_tma_maps, geo, o, tiles_j, ntiles, 0, 0, c->nx, 0);
  } else if (tiled) {
    c->last_asm_kernel = 2;
    k_assemble_tiled<<<dim3(capped_blocks(c, (size_t)tiles_i * tiles_j, 1), (unsigned)c->nbatch, 1), kTI * kTJ, 0, c->stream>>>(
        c->f, c->yprev, c->tx, c->ty, c->cxy, c->U, c->Ud, geo, o, tiles_j, 0, 0, c->nx, 0);
  } else {
    c->last_asm_kernel = 1;
    k_assemble<0><<<grid_of(c), kBlock, 0, c->stream>>>(c->f, c->yprev, c->tx, c->ty, c->cxy, c->U, c->Ud, geo, o);
  }
}

// ---- multigrid preconditioner: allocation, per-step setup, V-cycle, iteration ----
int mg_alloc(sy2d_ctx* c) {
  const int nlev = mg_level_count(c);
  if (c->mg_nlev == nlev && !c->mg_bufs.empty()) return SY2D_OK;
  for (double* b : c->mg_bufs) cudaFree(b);
  c->mg_bufs.clear();
  int rc = xl_alloc(c);
  if (rc) return rc;
  auto grab = [&](size_t n, double** out) -> int {
    CU(cudaMalloc(reinterpret_cast<void**>(out), n * sizeof(double)));
    CU(cudaMemsetAsync(*out, 0, n * sizeof(double), c->stream));
    c->mg_bufs.push_back(*out);
    return SY2D_OK;
  };
  // Slab contexts: every level array carries one halo row on each side (like the fine-grid arrays); the level
  // pointers address the first OWNED row, so the kernels index rows 0 .. rows-1 and the residual kernels reach
  // the halo rows at -1 and rows.
  const int rows = mg_rows(c), halo_rows = c->slab ? 2 : 0;
  c->mg_seg = rows <= 1024 ? 8 : 16;
  std::memset(&c->mg, 0, sizeof c->mg);
  c->mg.nlev = nlev;
  double* t0 = nullptr;
  if ((rc = grab(c->total, &c->mg_om0)) || (rc = grab(c->total, &c->mg_shat)) || (rc = grab(c->total, &t0))) return rc;
  const size_t off0 = c->slab ? (size_t)c->ny : 0;
  MgLevel& l0 = c->mg.lv[0];
  l0.wW = c->wW + off0; l0.wE = c->wE + off0; l0.wS = c->wS + off0; l0.wN = c->wN + off0; l0.om = c->mg_om0 + off0;
  l0.l = c->xl_l + off0; l0.dinv = c->xl_dinv + off0; l0.e = c->xl_e + off0;
  l0.r = nullptr; l0.z = nullptr; l0.t = t0 + off0;
  l0.ny = c->ny; l0.N = (size_t)rows * c->ny;
  int ny = c->ny;
  for (int k = 1; k < nlev; ++k) {
    ny /= 2;
    const size_t n = (size_t)(rows + halo_rows) * ny * c->nbatch;
    const size_t off = c->slab ? (size_t)ny : 0;
    double* a[11];
    for (double*& q : a) if ((rc = grab(n, &q))) return rc;
    MgLevel& lv = c->mg.lv[k];
    lv.wW = a[0] + off; lv.wE = a[1] + off; lv.wS = a[2] + off; lv.wN = a[3] + off; lv.om = a[4] + off;
    lv.l = a[5] + off; lv.dinv = a[6] + off; lv.e = a[7] + off;
    lv.r = a[8] + off; c->mg_rc[k] = a[8] + off;
    lv.z = a[9] + off; lv.t = a[10] + off;
    lv.ny = ny; lv.N = (size_t)rows * ny;
  }
  if (c->slab) {
    if (c->nranks > kMgMaxRanks) return fail(c, SY2D_ERR_INVALID, "multigrid in slab mode: at most %d ranks", kMgMaxRanks);
    int nyl = c->ny;
    for (int k = 0; k < nlev; ++k, nyl /= 2) {
      const size_t n = (size_t)(rows + halo_rows) * nyl;
      double *w = nullptr, *v = nullptr;
      if ((rc = grab(n, &w)) || (rc = grab(n, &v)) || (rc = grab((size_t)c->nranks * 4 * nyl, &c->mg_sp_all[k]))) return rc;
      c->mg_spW[k] = w + nyl; c->mg_spV[k] = v + nyl;   // first owned row, like the level arrays
    }
    if ((rc = grab((size_t)6 * c->ny, &c->mg_tips)) || (rc = grab((size_t)c->nranks * 2 * c->ny, &c->mg_tips_all)) ||
        (rc = grab((size_t)2 * c->ny, &c->mg_coef)))
      return rc;
  }
  if (!c->mg_tail_ctr) {
    CU(cudaMalloc(reinterpret_cast<void**>(&c->mg_tail_ctr), c->nbatch * sizeof(unsigned)));
    CU(cudaMemsetAsync(c->mg_tail_ctr, 0, c->nbatch * sizeof(unsigned), c->stream));
  }
  c->mg_nlev = nlev;
  return SY2D_OK;
}

// First level of the V-cycle that runs inside the fused coarse-tail kernel (sy2d_mg.cuh, k_mg_tail), or mg_nlev when the
// cycle is not fused: single-GPU contexts only (slab ranks exchange halos and spike tips between the stages), all fused
// levels must have the predicate-free shape, and the CTAs of all problems must be resident at once (one per SM).
int mg_tail_cols(const sy2d_ctx* c) {   // columns per CTA inside the tail: 4 while the CTA stays within its thread limit
  const int rows = mg_rows(c), seg = c->mg_seg, nseg = (rows + seg - 1) / seg;
  return nseg * 4 <= 512 ? 4 : 2;
}
int mg_tail_first_level(const sy2d_ctx* c) {
  const int L = c->mg_nlev;
  if (c->slab || c->mg_tail_ny <= 0 || c->profiling || c->nbatch > c->sm_count) return L;
  const int rows = mg_rows(c), seg = c->mg_seg, nseg = (rows + seg - 1) / seg, cols = mg_tail_cols(c);
  if (rows % seg != 0 || (nseg * cols) % 32 != 0 || nseg * cols > 512) return L;
  int k0 = L;
  for (int k = L - 1; k >= 0; --k) {
    const int ny = c->mg.lv[k].ny;
    if (ny > c->mg_tail_ny || ny % cols != 0 || ny % 2 != 0) break;
    k0 = k;
  }
  return k0;
}

// coarse operators (level by level) and the line LU of every level (one launch) for this step's operator
void mg_setup(sy2d_ctx* c) {
  Prof p(c, SY2D_K_MG_SETUP);
  for (int k = 0; k + 1 < c->mg_nlev; ++k) {
    const MgLevel& f = c->mg.lv[k];
    const MgLevel& g = c->mg.lv[k + 1];
    k_mg_coarsen<<<dim3(capped_blocks(c, g.N, kBlock), (unsigned)c->nbatch, 1), kBlock, 0, c->stream>>>(
        f, const_cast<double*>(g.wW), const_cast<double*>(g.wE), const_cast<double*>(g.wS), const_cast<double*>(g.wN),
        const_cast<double*>(g.om), mg_rows(c));
  }
  k_mg_factor<<<dim3((unsigned)((c->ny + 63) / 64), (unsigned)c->nbatch, (unsigned)c->mg_nlev), 64, 0, c->stream>>>(c->mg, mg_rows(c));
  c->launches += c->mg_nlev;
}

// Shape of the line kernel: rows per thread (SEG) x columns per CTA (COLS); a CTA holds all nx / SEG
// segments of its columns.  SEG = 8 up to nx = 1024 (at most 1024 threads of ~60 registers), else 16 (512
// threads).  COLS is chosen PER LEVEL: as many columns as the thread limit allows (8 at most: 64-byte row
// chunks) on levels with enough columns to fill the GPU, fewer on the coarse levels, so that a level is always
// spread over ~128 CTAs - a CTA streams its columns' factors through one SM's L1 (131 KB per phase with 8
// columns at nx = 1024), and that per-SM streaming time, not DRAM, is what a coarse-level solve waits for.
template <int SEG, int COLS, int MODE>
void mg_line_shape(sy2d_ctx* c, const MgLevel& lv, const double* zc) {
  const int rows = mg_rows(c);
  MgArgs a{c->scal, rows, c->slab ? 1 : 0, c->slab ? c->mg_tips : nullptr};
  if (MODE == 0 && c->mg_fuse_pending) {   // the Krylov update that feeds this V-cycle is formed inside its first line solve
    a.fuse = c->mg_fuse_pending;
    a.f_rhs = c->rhs; a.f_r = c->r; a.f_v = c->v;
    a.f_dst = const_cast<double*>(lv.r);
    c->mg_fuse_pending = 0;
  }
  const int nseg = (rows + SEG - 1) / SEG;
  const int threads = (nseg * COLS + 31) / 32 * 32;
  const size_t smem = (size_t)3 * COLS * (nseg + 1) * sizeof(double);
  const dim3 g((unsigned)((lv.ny + COLS - 1) / COLS), (unsigned)c->nbatch, 1);
  const bool full = rows % SEG == 0 && lv.ny % COLS == 0 && nseg * COLS == threads;
  if (full && nseg % 32 == 0 && c->mg_line_pre) {   // one exposed memory latency per solve, scans on all warps (sy2d_mg.cuh, PRE)
    const size_t smem_pre = ((size_t)4 * COLS * (nseg + 1) + (size_t)2 * COLS * (nseg / 32) + (size_t)2 * SEG * threads) * sizeof(double);
    static bool attr_set = false;   // per instantiation
    if (!attr_set) {
      attr_set = cudaFuncSetAttribute(k_mg_line<SEG, COLS, MODE, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) == cudaSuccess;
      cudaGetLastError();
    }
    if (attr_set && smem_pre <= 200 * 1024) {
      k_mg_line<SEG, COLS, MODE, true, true><<<g, threads, smem_pre, c->stream>>>(lv, zc, a);
      return;
    }
  }
  if (full) k_mg_line<SEG, COLS, MODE, true><<<g, threads, smem, c->stream>>>(lv, zc, a);
  else k_mg_line<SEG, COLS, MODE, false><<<g, threads, smem, c->stream>>>(lv, zc, a);
}

template <int SEG>
int mg_line_cols(const sy2d_ctx* c, const MgLevel& lv) {
  const int nseg = (mg_rows(c) + SEG - 1) / SEG;
  const int max_threads = SEG <= 8 ? 1024 : 512;
  int cols = 8;
  while (cols > 1 && nseg * cols > max_threads) cols /= 2;
  // fewer columns per CTA while that brings the level closer to one CTA per SM (and a CTA keeps >= 2 warps)
  while (cols > 1 && lv.ny * c->nbatch / cols < 128 && nseg * (cols / 2) >= 64) cols /= 2;
  return cols;
}

// Long columns (16-row segments, i.e. more than 1024 rows): one group of 8 columns per thread-block cluster whose CTAs split
// the segments (k_mg_line_cluster); returns false when the shape does not fit (the stand-alone kernel runs then).
int mg_cluster_size(const sy2d_ctx* c, const MgLevel& lv) {
  if (!c->mg_cluster || c->mg_seg != 16 || lv.ny % 8 != 0) return 0;
  const int rows = mg_rows(c);
  if (rows % 16 != 0) return 0;
  const int nseg = rows / 16;
  for (int cl : {2, 4, 8}) {
    if (nseg % cl) continue;
    const int threads = nseg / cl * 8;
    if (threads <= 512 && threads % 32 == 0) return cl;
  }
  return 0;
}

template <int MODE, int CL>
cudaError_t mg_line_cluster_launch(sy2d_ctx* c, const MgLevel& lv, const double* zc) {
  const int rows = mg_rows(c), nseg_c = rows / (16 * CL);
  const MgArgs a{c->scal, rows, c->slab ? 1 : 0, c->slab ? c->mg_tips : nullptr};
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(lv.ny / 8 * CL), (unsigned)c->nbatch, 1);
  cfg.blockDim = dim3((unsigned)(nseg_c * 8), 1, 1);
  cfg.dynamicSmemBytes = (size_t)(4 * 8 * (nseg_c + 1) + 4 * 8) * sizeof(double);
  cfg.stream = c->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, k_mg_line_cluster<16, 8, MODE, CL>, lv, zc, a);
}

template <int MODE>
bool mg_line_cluster(sy2d_ctx* c, const MgLevel& lv, const double* zc) {
  cudaError_t e = cudaErrorInvalidValue;
  switch (mg_cluster_size(c, lv)) {
    case 2: e = mg_line_cluster_launch<MODE, 2>(c, lv, zc); break;
    case 4: e = mg_line_cluster_launch<MODE, 4>(c, lv, zc); break;
    case 8: e = mg_line_cluster_launch<MODE, 8>(c, lv, zc); break;
    default: return false;
  }
  if (e != cudaSuccess) { cudaGetLastError(); return false; }   // e.g. the cluster cannot be scheduled: stand-alone kernel
  return true;
}

template <int SEG, int MODE>
void mg_line_seg(sy2d_ctx* c, const MgLevel& lv, const double* zc) {
  if (SEG == 16 && mg_line_cluster<MODE>(c, lv, zc)) return;
  switch (mg_line_cols<SEG>(c, lv)) {
    case 8: mg_line_shape<SEG, 8, MODE>(c, lv, zc); break;
    case 4: mg_line_shape<SEG, 4, MODE>(c, lv, zc); break;
    case 2: mg_line_shape<SEG, 2, MODE>(c, lv, zc); break;
    default: mg_line_shape<SEG, 1, MODE>(c, lv, zc); break;
  }
}

int slab_gather(sy2d_ctx* c, const double* src, double* dst, size_t count);

// Row-slab mode: turns the local line solutions just written to lv.z into the solutions of the global lines
// (sy2d_mg.cuh, "exact x-lines across ranks"): all-gather of the tips, reduced system per column, one correction pass.
int mg_spike_fix(sy2d_ctx* c, const MgLevel& lv, int k) {
  int rc = slab_gather(c, c->mg_tips, c->mg_tips_all, (size_t)2 * lv.ny);
  if (rc) return rc;
  k_mg_spike_reduced<<<(unsigned)((lv.ny + 127) / 128), 128, 0, c->stream>>>(c->mg_tips_all, c->mg_sp_all[k], c->mg_coef, c->scal, c->rank, c->nranks, lv.ny);
  k_mg_spike_apply<<<capped_blocks(c, lv.N / 2, kBlock), kBlock, 0, c->stream>>>(lv.z, c->mg_spW[k], c->mg_spV[k], c->mg_coef, c->scal, mg_rows(c), lv.ny);
  CU(cudaGetLastError());
  return SY2D_OK;
}

template <int MODE>
int mg_line(sy2d_ctx* c, const MgLevel& lv, const double* zc, int k) {
  Prof p(c, SY2D_K_MG_LINE);
  if (c->mg_seg == 8) mg_line_seg<8, MODE>(c, lv, zc);
  else mg_line_seg<16, MODE>(c, lv, zc);
  if (c->slab && c->mg_spike) return mg_spike_fix(c, lv, k);
  return SY2D_OK;
}

// Row-slab mode, once per time step after mg_setup: the two spikes of every level (damped like the smoother's
// output, so that the correction pass needs no extra factor) and the all-gathered spike tips.
int mg_spike_setup(sy2d_ctx* c) {
  c->mg_spike = false;   // the spike solves themselves are plain local solves
  for (int k = 0; k < c->mg_nlev; ++k) {
    MgLevel lv = c->mg.lv[k];
    const int rows = mg_rows(c);
    double* rhs = lv.t;   // free at setup time
    for (int which = 0; which < 2; ++which) {
      k_mg_spike_rhs<<<capped_blocks(c, lv.N, kBlock), kBlock, 0, c->stream>>>(rhs, which == 0 ? lv.wW : lv.wE, which, rows, lv.ny);
      MgLevel sv = lv;
      sv.r = rhs; sv.z = which == 0 ? c->mg_spW[k] : c->mg_spV[k];
      int rc = mg_line<0>(c, sv, nullptr, k);
      if (rc) return rc;
      // tips of the undamped spike: [2][ny_l] -> slots (2 which, 2 which + 1) of the local [4][ny_l] block
      CU(cudaMemcpyAsync(c->mg_tips + (size_t)(2 + 2 * which) * lv.ny, c->mg_tips, (size_t)2 * lv.ny * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    }
    // local block [4][ny_l] sits at mg_tips + 2 ny_l (behind the [2][ny_l] the line kernel writes)
    int rc = slab_gather(c, c->mg_tips + (size_t)2 * lv.ny, c->mg_sp_all[k], (size_t)4 * lv.ny);
    if (rc) return rc;
  }
  CU(cudaGetLastError());
  c->mg_spike = true;
  return SY2D_OK;
}

template <int KIND>
void mg_resid(sy2d_ctx* c, const MgLevel& lv, const double* zc, const double* omc, double* rc) {
  Prof p(c, SY2D_K_MG_RESID);
  const MgArgs a{c->scal, mg_rows(c), c->slab ? 1 : 0, nullptr};
  k_mg_resid<KIND><<<dim3(capped_blocks(c, lv.N / 2, kBlock), (unsigned)c->nbatch, 1), kBlock, 0, c->stream>>>(lv, zc, omc, rc, a);
}

int slab_halo_exchange_n(sy2d_ctx* c, double* a, int ny_l);

// z = V(r): one V(1,1) cycle on the fine grid; r and z are [nbatch][nx][ny] device vectors (local arrays incl. the
// halo rows on a slab context).  Slab contexts exchange one halo row of the iterate with both neighbours before
// every residual that reads it (10 small ncclSend/Recv pairs per cycle); the caller exchanges the halo of z.
int mg_vcycle(sy2d_ctx* c, const double* r, double* z) {
  const int L = c->mg_nlev;
  const size_t off0 = c->slab ? (size_t)c->ny : 0;
  MgLevel lv0 = c->mg.lv[0];
  lv0.r = r + off0; lv0.z = z + off0;
  auto level = [&](int k) -> const MgLevel& { return k == 0 ? lv0 : c->mg.lv[k]; };
  auto halo = [&](int k) -> int {   // the iterate of level k
    if (!c->slab) return SY2D_OK;
    const MgLevel& lv = level(k);
    return slab_halo_exchange_n(c, lv.z - lv.ny, lv.ny);
  };
  int rc = SY2D_OK;
  const double scale0 = c->cur_cells;
  const int k0 = mg_tail_first_level(c);
  if (k0 < L) {
    // levels 0 .. k0-1 as separate launches, levels k0 .. L-1 (down, coarsest sweeps, up) inside ONE kernel
    for (int k = 0; k < k0; ++k) {
      if ((rc = mg_line<0>(c, level(k), nullptr, k))) return rc;
      mg_resid<1>(c, level(k), nullptr, c->mg.lv[k + 1].om, c->mg_rc[k + 1]);
    }
    {
      Prof p(c, SY2D_K_MG_LINE);
      MgTailArgs t;
      t.L = c->mg;
      t.L.lv[0] = lv0;
      t.a = MgArgs{c->scal, mg_rows(c), 0, nullptr};
      t.barrier = c->mg_tail_ctr;
      t.k0 = k0;
      t.coarse_sweeps = mg_coarse_sweeps(c);
      const int cols = mg_tail_cols(c), seg = c->mg_seg, nseg = mg_rows(c) / seg;
      const int groups = level(k0).ny / cols;
      const unsigned ctas = (unsigned)std::max(1, std::min(groups, c->sm_count / c->nbatch));
      const dim3 g(ctas, (unsigned)c->nbatch, 1);
      const int threads = nseg * cols;
      const size_t smem = (size_t)3 * cols * (nseg + 1) * sizeof(double);
      if (seg == 8 && cols == 4) k_mg_tail<8, 4><<<g, threads, smem, c->stream>>>(t);
      else if (seg == 8) k_mg_tail<8, 2><<<g, threads, smem, c->stream>>>(t);
      else if (cols == 4) k_mg_tail<16, 4><<<g, threads, smem, c->stream>>>(t);
      else k_mg_tail<16, 2><<<g, threads, smem, c->stream>>>(t);
    }
    for (int k = k0 - 1; k >= 0; --k) {
      mg_resid<2>(c, level(k), c->mg.lv[k + 1].z, nullptr, nullptr);
      if ((rc = mg_line<1>(c, level(k), c->mg.lv[k + 1].z, k))) return rc;
    }
    c->cur_cells = scale0;
    return SY2D_OK;
  }
  for (int k = 0; k + 1 < L; ++k) {
    c->cur_cells = scale0 / (double)(1 << k);
    if ((rc = mg_line<0>(c, level(k), nullptr, k))) return rc;
    if ((rc = halo(k))) return rc;
    mg_resid<1>(c, level(k), nullptr, c->mg.lv[k + 1].om, c->mg_rc[k + 1]);
  }
  c->cur_cells = scale0 / (double)(1 << (L - 1));
  if ((rc = mg_line<0>(c, level(L - 1), nullptr, L - 1))) return rc;
  for (int sweep = 1; sweep < mg_coarse_sweeps(c); ++sweep) {
    if ((rc = halo(L - 1))) return rc;
    mg_resid<0>(c, level(L - 1), nullptr, nullptr, nullptr);
    if ((rc = mg_line<2>(c, level(L - 1), nullptr, L - 1))) return rc;
  }
  for (int k = L - 2; k >= 0; --k) {
    c->cur_cells = scale0 / (double)(1 << k);
    if ((rc = halo(k + 1))) return rc;   // the correction; the halo of z_k is still the one exchanged on the way down
    mg_resid<2>(c, level(k), c->mg.lv[k + 1].z, nullptr, nullptr);
    if ((rc = mg_line<1>(c, level(k), c->mg.lv[k + 1].z, k))) return rc;
  }
  c->cur_cells = scale0;
  return SY2D_OK;
}
int mg_kernels_per_vcycle(const sy2d_ctx* c) {
  const int k0 = mg_tail_first_level(c);
  if (k0 < c->mg_nlev) return 4 * k0 + 1;
  return 4 * (c->mg_nlev - 1) + 1 + 2 * (mg_coarse_sweeps(c) - 1);
}

// The fused p / s updates need the first stage of the V-cycle to be a k_mg_line launch on level 0 (not the cluster kernel of long
// columns, not the fused coarse tail) on a single-GPU context.
bool mg_fuse_updates(sy2d_ctx* c) {
  return c->mg_fuse && !c->slab && c->mg_nlev >= 2 && mg_cluster_size(c, c->mg.lv[0]) == 0 && mg_tail_first_level(c) > 0;
}

int launch_iteration_mg(sy2d_ctx* c) {
  int rc = SY2D_OK;
  KrylovVecs k = krylov(c);
  const dim3 g2(capped_blocks(c, c->N / 2, kBlock), (unsigned)c->nbatch, 1);
  double* phat = c->xl_hat;
  double* shat = c->mg_shat;
  // single GPU, level 0 solved by the stand-alone line kernel: the p and s updates are formed inside the first line solve of
  // their V-cycle (one launch and one pass over p / s less per half-iteration)
  const bool fuse = mg_fuse_updates(c);
  if (fuse) c->mg_fuse_pending = 1;
  else { Prof p(c, SY2D_K_P_UPDATE); k_p_update2<<<g2, kBlock, 0, c->stream>>>(k, c->N); }
  if ((rc = mg_vcycle(c, c->p, phat))) return rc;
  {
    Prof p(c, SY2D_K_SPMV_V);
    KrylovVecs kv = k;
    kv.p = phat;   // v = A phat, (rhat, v)
    k_spmv_v2<<<g2, kBlock, 0, c->stream>>>(kv, c->N, c->ny);
  }
  if (fuse) c->mg_fuse_pending = 2;
  else { Prof p(c, SY2D_K_S_UPDATE); k_s_update2<<<g2, kBlock, 0, c->stream>>>(k, c->N); }
  if ((rc = mg_vcycle(c, c->s, shat))) return rc;
  { Prof p(c, SY2D_K_SPMV_T); k_mg_spmv_t<<<g2, kBlock, 0, c->stream>>>(k, shat, c->N, c->ny); }
  { Prof p(c, SY2D_K_XR_UPDATE); k_mg_xr<<<g2, kBlock, 0, c->stream>>>(k, phat, shat, c->N); }
  return SY2D_OK;
}

int launch_iteration(sy2d_ctx* c) {
  if (lockstep_mg(c)) return launch_iteration_mg(c);
  if (lockstep_xline(c)) { launch_iteration_xline(c); return SY2D_OK; }
  const KrylovVecs k = krylov(c);
  if (c->ny % 2 == 0) {  // two cells per thread, 16-byte accesses
    const dim3 g2(capped_blocks(c, c->N / 2, kBlock), (unsigned)c->nbatch, 1);
    { Prof p(c, SY2D_K_P_UPDATE); k_p_update2<<<g2, kBlock, 0, c->stream>>>(k, c->N); }
    { Prof p(c, SY2D_K_SPMV_V); k_spmv_v2<<<g2, kBlock, 0, c->stream>>>(k, c->N, c->ny); }
    { Prof p(c, SY2D_K_S_UPDATE); k_s_update2<<<g2, kBlock, 0, c->stream>>>(k, c->N); }
    { Prof p(c, SY2D_K_SPMV_T); k_spmv_t2<<<g2, kBlock, 0, c->stream>>>(k, c->N, c->ny); }
    { Prof p(c, SY2D_K_XR_UPDATE); k_xr_update2<<<g2, kBlock, 0, c->stream>>>(k, c->N); }
    return SY2D_OK;
  }
  const dim3 g = grid_of(c);
  { Prof p(c, SY2D_K_P_UPDATE); k_p_update<<<g, kBlock, 0, c->stream>>>(k, c->N); }
  { Prof p(c, SY2D_K_SPMV_V); k_spmv_v<<<g, kBlock, 0, c->stream>>>(k, c->N, c->ny); }
  { Prof p(c, SY2D_K_S_UPDATE); k_s_update<<<g, kBlock, 0, c->stream>>>(k, c->N); }
  { Prof p(c, SY2D_K_SPMV_T); k_spmv_t<<<g, kBlock, 0, c->stream>>>(k, c->N, c->ny); }
  { Prof p(c, SY2D_K_XR_UPDATE); k_xr_update<<<g, kBlock, 0, c->stream>>>(k, c->N); }
  return SY2D_OK;
}
int kernels_per_iteration(const sy2d_ctx* c) {
  if (!lockstep_mg(c)) return 5;
  return (mg_fuse_updates(const_cast<sy2d_ctx*>(c)) ? 3 : 5) + 2 * mg_kernels_per_vcycle(c);
}
// with multigrid an iteration is ~50 launches and a solve ~15 iterations: poll more often
int effective_check_every(const sy2d_ctx* c) { return lockstep_mg(c) ? std::min(c->opt.check_every, 4) : c->opt.check_every; }

int build_chunk_graph(sy2d_ctx* c) {
  const int variant = lockstep_mg(c) ? 2 + c->mg_nlev + 16 * mg_coarse_sweeps(c) : (lockstep_xline(c) ? 1 : 0);
  const int check_every = effective_check_every(c);
  if (c->chunk_exec && c->chunk_iters == check_every && c->chunk_variant == variant) return SY2D_OK;
  c->chunk_variant = variant;
  if (c->chunk_exec) { cudaGraphExecDestroy(c->chunk_exec); c->chunk_exec = nullptr; }
  if (c->one_exec) { cudaGraphExecDestroy(c->one_exec); c->one_exec = nullptr; }
  cudaGraph_t graph = nullptr;
  if (lockstep_mg(c)) {
    CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    int rci = launch_iteration(c);
    CU(cudaMemcpyAsync(c->h_nactive, c->d_nactive, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamEndCapture(c->stream, &graph));
    if (rci) { cudaGraphDestroy(graph); return rci; }
    CU(cudaGraphInstantiate(&c->one_exec, graph, 0));
    cudaGraphDestroy(graph);
    graph = nullptr;
  }
  CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
  int rcc = SY2D_OK;
  for (int it = 0; it < check_every && !rcc; ++it) rcc = launch_iteration(c);
  CU(cudaMemcpyAsync(c->h_nactive, c->d_nactive, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamEndCapture(c->stream, &graph));
  if (rcc) { cudaGraphDestroy(graph); return rcc; }
  CU(cudaGraphInstantiate(&c->chunk_exec, graph, 0));
  cudaGraphDestroy(graph);
  c->chunk_iters = check_every;
  return SY2D_OK;
}

int collect_profile(sy2d_ctx* c) {
  if (!c->profiling) return SY2D_OK;
  for (size_t k = 0; k < c->ev_used; ++k) {
    float t = 0.f;
    CU(cudaEventElapsedTime(&t, c->ev_pool[k].first, c->ev_pool[k].second));
    c->prof.ms[c->ev_class[k]] += t;
    c->prof.launches[c->ev_class[k]] += 1;
    c->prof.cells[c->ev_class[k]] += c->ev_cells[k];
  }
  c->ev_used = 0;
  return SY2D_OK;
}

int step_slab(sy2d_ctx* c, int nsteps, sy2d_stats* stats);
int slab_halo_exchange(sy2d_ctx* c, double* a);
void slab_comm_destroy(sy2d_ctx* c);
double* own_ptr(const sy2d_ctx* c, double* a);
size_t own_elems(const sy2d_ctx* c);

template <class T>
int dalloc(sy2d_ctx* c, T** p, size_t n) {
  CU(cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T)));
  return SY2D_OK;
}

}  // namespace


// ---------------------------------------------------------------------------------------------
// Row-slab mode: NCCL through dlopen (the library does not link libnccl; a process that never
// creates a slab context never needs it, and a process that already loaded NCCL - e.g. through
// torch - reuses that copy).
// ---------------------------------------------------------------------------------------------
#include <dlfcn.h>
#include <nccl.h>

namespace {
struct NcclApi {
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclCommAbort) CommAbort = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  bool ok = false;
  std::string why;
};

NcclApi& nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { api.why = std::string("cannot load libnccl.so.2: ") + dlerror(); return api; }
#define SY2D_NCCL_SYM(name) api.name = reinterpret_cast<decltype(api.name)>(dlsym(h, "nccl" #name)); if (!api.name) { api.why = "missing symbol nccl" #name; return api; }
  SY2D_NCCL_SYM(GetUniqueId) SY2D_NCCL_SYM(CommInitRank) SY2D_NCCL_SYM(CommDestroy) SY2D_NCCL_SYM(AllGather)
  SY2D_NCCL_SYM(Send) SY2D_NCCL_SYM(Recv) SY2D_NCCL_SYM(GroupStart) SY2D_NCCL_SYM(GroupEnd) SY2D_NCCL_SYM(GetErrorString)
#undef SY2D_NCCL_SYM
  api.CommAbort = reinterpret_cast<decltype(api.CommAbort)>(dlsym(h, "ncclCommAbort"));   // optional
  api.ok = true;
  return api;
}

#define NC(call)                                                                                             \
  do {                                                                                                       \
    ncclResult_t r_ = (call);                                                                                \
    if (r_ != ncclSuccess) return fail(c, SY2D_ERR_CUDA, "NCCL: %s failed: %s", #call, nccl().GetErrorString(r_)); \
  } while (0)

double* own_ptr(const sy2d_ctx* c, double* a) { return c->slab ? a + c->ny : a; }
size_t own_elems(const sy2d_ctx* c) { return c->slab ? (size_t)(c->nx - 2) * c->ny : c->total; }

// ---- transports: how the slabs of one grid exchange halo lines and small vectors ----
// halo(a, ny_l): one line of ny_l doubles with each neighbour - my first owned row -> the bottom halo of rank-1, my last
//   owned row -> the top halo of rank+1 (a = local array incl. the halo rows; contiguous 8 ny_l bytes).
// gather(src, dst, count): count doubles of every rank -> dst[nranks][count] on every rank.
// abort(): this rank gives up (CUDA / transport failure); the peers must not wait for it for ever.
// Convergence decisions are taken from all-gathered scalars reduced in rank order, i.e. they are IDENTICAL on every rank:
// SY2D_ERR_NOT_CONVERGED is returned by all ranks of a step together, only hard failures can separate the ranks.
struct SlabTransport {
  virtual ~SlabTransport() {}
  virtual int halo(sy2d_ctx* c, double* a, int ny_l) = 0;
  virtual int gather(sy2d_ctx* c, const double* src, double* dst, size_t count) = 0;
  virtual void abort(sy2d_ctx* c) = 0;
  virtual const char* name() const = 0;
  virtual bool capturable() const { return false; }   // the calls are plain stream work that a CUDA graph can record
};

// NCCL over NVLink / NVSwitch: one process (or thread) per GPU.
struct NcclTransport : SlabTransport {
  ncclComm_t comm = nullptr;
  ~NcclTransport() override { if (comm && nccl().ok) nccl().CommDestroy(comm); }
  int halo(sy2d_ctx* c, double* a, int ny_l) override {
    const size_t ny = (size_t)ny_l;
    const int rows = c->nx - 2;
    NC(nccl().GroupStart());
    if (c->rank > 0) {
      NC(nccl().Send(a + ny, ny, ncclDouble, c->rank - 1, comm, c->stream));
      NC(nccl().Recv(a, ny, ncclDouble, c->rank - 1, comm, c->stream));
    }
    if (c->rank < c->nranks - 1) {
      NC(nccl().Send(a + (size_t)rows * ny, ny, ncclDouble, c->rank + 1, comm, c->stream));
      NC(nccl().Recv(a + (size_t)(rows + 1) * ny, ny, ncclDouble, c->rank + 1, comm, c->stream));
    }
    NC(nccl().GroupEnd());
    return SY2D_OK;
  }
  int gather(sy2d_ctx* c, const double* src, double* dst, size_t count) override {
    NC(nccl().AllGather(src, dst, count, ncclDouble, comm, c->stream));
    return SY2D_OK;
  }
  void abort(sy2d_ctx*) override {
    if (comm && nccl().CommAbort) { nccl().CommAbort(comm); comm = nullptr; }   // frees this rank's resources without a collective teardown
  }
  const char* name() const override { return "nccl"; }
  bool capturable() const override { return true; }
};
}  // namespace

// In-process transport: the P slab contexts live in ONE process (one host thread per context, on one device or on
// several), lines and small vectors move with cudaMemcpyAsync between the contexts' buffers, ordered by CUDA events
// and a host barrier.  This is what lets a single-GPU box run - and test - the slab decomposition at P = 2, 4, 8:
// the kernels, the spike-coupled multigrid and the reduction order are exactly those of the NCCL runs.
struct sy2d_local_group {
  int nranks = 0;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  unsigned long long gen = 0;
  bool aborted = false;
  int attached = 0;
  std::vector<const double*> ptr;
  std::vector<int> rows;
  std::vector<cudaEvent_t> ready, done;
  // false once the group was aborted: the caller must fail instead of waiting for a rank that gave up
  bool barrier() {
    std::unique_lock<std::mutex> lk(mu);
    if (aborted) return false;
    const unsigned long long g0 = gen;
    if (++arrived == nranks) {
      arrived = 0;
      ++gen;
      cv.notify_all();
      return true;
    }
    cv.wait(lk, [&] { return gen != g0 || aborted; });
    return gen != g0;
  }
  void abort() {
    std::lock_guard<std::mutex> lk(mu);
    aborted = true;
    cv.notify_all();
  }
};

namespace {
struct LocalTransport : SlabTransport {
  sy2d_local_group* g = nullptr;
  int rank = 0;
  ~LocalTransport() override {
    if (!g) return;
    if (g->ready[rank]) cudaEventDestroy(g->ready[rank]);
    if (g->done[rank]) cudaEventDestroy(g->done[rank]);
    g->ready[rank] = g->done[rank] = nullptr;
    std::lock_guard<std::mutex> lk(g->mu);
    g->attached -= 1;
  }
  int halo(sy2d_ctx* c, double* a, int ny_l) override {
    const size_t ny = (size_t)ny_l;
    const int r = rank, P = g->nranks, rows = c->nx - 2;
    g->ptr[r] = a;
    g->rows[r] = rows;
    CU(cudaEventRecord(g->ready[r], c->stream));
    if (!g->barrier()) return fail(c, SY2D_ERR_CUDA, "slab group aborted by another rank");
    if (r > 0) {
      CU(cudaStreamWaitEvent(c->stream, g->ready[r - 1], 0));
      CU(cudaMemcpyAsync(a, g->ptr[r - 1] + (size_t)g->rows[r - 1] * ny, ny * sizeof(double), cudaMemcpyDefault, c->stream));
    }
    if (r < P - 1) {
      CU(cudaStreamWaitEvent(c->stream, g->ready[r + 1], 0));
      CU(cudaMemcpyAsync(a + (size_t)(rows + 1) * ny, g->ptr[r + 1] + ny, ny * sizeof(double), cudaMemcpyDefault, c->stream));
    }
    CU(cudaEventRecord(g->done[r], c->stream));
    if (!g->barrier()) return fail(c, SY2D_ERR_CUDA, "slab group aborted by another rank");
    // my owned rows may only be overwritten once the neighbours have copied them
    if (r > 0) CU(cudaStreamWaitEvent(c->stream, g->done[r - 1], 0));
    if (r < P - 1) CU(cudaStreamWaitEvent(c->stream, g->done[r + 1], 0));
    return SY2D_OK;
  }
  int gather(sy2d_ctx* c, const double* src, double* dst, size_t count) override {
    const int r = rank, P = g->nranks;
    g->ptr[r] = src;
    CU(cudaEventRecord(g->ready[r], c->stream));
    if (!g->barrier()) return fail(c, SY2D_ERR_CUDA, "slab group aborted by another rank");
    for (int q = 0; q < P; ++q) {
      if (q != r) CU(cudaStreamWaitEvent(c->stream, g->ready[q], 0));
      CU(cudaMemcpyAsync(dst + (size_t)q * count, g->ptr[q], count * sizeof(double), cudaMemcpyDefault, c->stream));
    }
    CU(cudaEventRecord(g->done[r], c->stream));
    if (!g->barrier()) return fail(c, SY2D_ERR_CUDA, "slab group aborted by another rank");
    for (int q = 0; q < P; ++q)
      if (q != r) CU(cudaStreamWaitEvent(c->stream, g->done[q], 0));
    return SY2D_OK;
  }
  void abort(sy2d_ctx*) override { g->abort(); }
  const char* name() const override { return "local"; }
};

void slab_comm_destroy(sy2d_ctx* c) {
  delete c->tp;
  c->tp = nullptr;
}

int slab_halo_exchange_n(sy2d_ctx* c, double* a, int ny_l) { return c->tp->halo(c, a, ny_l); }
int slab_halo_exchange(sy2d_ctx* c, double* a) { return c->tp->halo(c, a, c->ny); }
int slab_gather(sy2d_ctx* c, const double* src, double* dst, size_t count) { return c->tp->gather(c, src, dst, count); }

// accumulators of every rank -> scalars on every rank (identical summation order everywhere)
int slab_reduce(sy2d_ctx* c, int phase, const KrylovVecs& k) {
  Prof p(c, SY2D_K_OTHER);
  int rc = slab_gather(c, &c->scal->acc_rv, c->d_gather, 5);
  if (rc) return rc;
  k_slab_scalars<<<1, 32, 0, c->stream>>>(phase, c->scal, c->d_gather, c->nranks, k);
  CU(cudaGetLastError());
  return SY2D_OK;
}

int step_slab_impl(sy2d_ctx* c, int nsteps, sy2d_stats* stats);

// A hard failure (CUDA, transport) on this rank must not leave the peers waiting in the next collective: the transport
// is aborted (in-process group: the peers' calls fail at once; NCCL: the communicator is torn down without a collective).
// SY2D_ERR_NOT_CONVERGED is not such a failure - every rank returns it for the same step.
int step_slab(sy2d_ctx* c, int nsteps, sy2d_stats* stats) {
  const int rc = step_slab_impl(c, nsteps, stats);
  if (rc == SY2D_ERR_CUDA && c->tp) c->tp->abort(c);
  return rc;
}

int step_slab_impl(sy2d_ctx* c, int nsteps, sy2d_stats* stats) {
  if (c->ny % 2) return fail(c, SY2D_ERR_INVALID, "slab mode needs an even ny");
  const bool budget = c->opt.reserved[1] != 0;   // fixed iteration budget without a convergence error (bench)
  const Geometry geo = geometry(c);
  const KrylovVecs k = krylov(c);
  const int rows = c->nx - 2;
  const size_t own = (size_t)rows * c->ny;
  const dim3 g2(capped_blocks(c, own / 2, kBlock), 1, 1);
  const int tiles_i = (rows + kTI - 1) / kTI, tiles_j = (c->ny + kTJ - 1) / kTJ;
  sy2d_stats st;
  std::memset(&st, 0, sizeof st);
  st.engine = 1;
  const bool mg = lockstep_mg(c);
  const bool xl = lockstep_xline(c);
  if (c->opt.precond == SY2D_PRECOND_MG && !mg)
    return fail(c, SY2D_ERR_INVALID, "sy2d_step (slab): the multigrid preconditioner needs 8 <= rows per rank <= 8192 and ny a multiple of 4, >= 16");
  st.precond = mg ? SY2D_PRECOND_MG : (xl ? SY2D_PRECOND_XLINE : SY2D_PRECOND_JACOBI);
  if (xl) { int rc0 = xl_alloc(c); if (rc0) return rc0; }
  if (mg) { int rc0 = mg_alloc(c); if (rc0) return rc0; }
  const int check_every = effective_check_every(c);
  const XlVecs xv = xl_vecs(c);
  const int nseg = (rows + kSeg - 1) / kSeg;
  const dim3 gs(capped_blocks(c, (size_t)nseg * c->ny, kSweepThreads), 1, 1);
  const dim3 gc(capped_blocks(c, own / 2, kBlock), 1, 1);
  CU(cudaEventRecord(c->ev_call0, c->stream));
  for (int step = 0; step < nsteps; ++step) {
    int rc = slab_halo_exchange(c, c->f);
    if (!rc) rc = slab_halo_exchange(c, c->yprev);
    if (rc) return rc;
    CU(cudaMemsetAsync(c->d_nactive, 0, 2 * sizeof(int), c->stream));
    AssembleOut o;
    std::memset(&o, 0, sizeof o);
    o.wW = c->wW; o.wE = c->wE; o.wS = c->wS; o.wN = c->wN; o.rhs = c->rhs; o.cs = c->cs;
    o.scal = c->scal; o.part = c->deterministic ? c->part : nullptr; o.part_stride = c->part_stride; o.n_active = c->d_nactive; o.tol = c->opt.tol; o.local_rows = c->nx;
    if (mg) o.om = c->mg_om0;
    if (c->opt.reserved[0] == 3 || (c->opt.reserved[0] == 0 && (c->asm_kernel == 1 || !c->have_tma))) {
      launch_march(c, geo, o, c->i_lo - 1, 1, rows + 1, 1);
    } else if (c->have_tma && (c->opt.reserved[0] == 6 || (c->opt.reserved[0] == 0 && c->asm_kernel == 3)) &&
               launch_wide(c, geo, o, tiles_i, c->i_lo - 1, 1, rows + 1, 1)) {
    } else if (c->have_tma && (c->opt.reserved[0] == 5 || (c->opt.reserved[0] == 0 && c->asm_kernel == 2)) &&
               launch_col(c, geo, o, tiles_i, tiles_j, c->i_lo - 1, 1, rows + 1, 1)) {
    } else if (c->have_tma && (c->opt.reserved[0] == 0 || c->opt.reserved[0] == 4 || c->opt.reserved[0] == 5 || c->opt.reserved[0] == 6)) {
      const int ntiles = tiles_i * tiles_j;
      const unsigned ctas = (unsigned)std::min<size_t>((size_t)ntiles, (size_t)c->sm_count * c->asm_ctas_per_sm);
      k_assemble_tma<<<dim3(ctas, 1, 1), kTI * kTJ, kTmaSmemBytes, c->stream>>>(c->d_tma_maps, geo, o, tiles_j, ntiles, c->i_lo - 1, 1, rows + 1, 1);
    } else {
      k_assemble_tiled<<<dim3(capped_blocks(c, (size_t)tiles_i * tiles_j, 1), 1, 1), kTI * kTJ, 0, c->stream>>>(
          c->f, c->yprev, c->tx, c->ty, c->cxy, c->U, c->Ud, geo, o, tiles_j, c->i_lo - 1, 1, rows + 1, 1);
    }
    CU(cudaGetLastError());
    rc = slab_reduce(c, 0, k);
    if (rc) return rc;
    if (xl) k_xl_factor<<<gs, kBlock, 0, c->stream>>>(xv, c->N);
    if (mg) {
      mg_setup(c);   // coarse operators and line LU of the owned rows (local) ...
      if (c->opt.reserved[2] != 2 && (rc = mg_spike_setup(c))) return rc;   // ... and the spikes that couple the ranks' lines
      if (c->opt.reserved[2] == 2) c->mg_spike = false;   // reserved[2] = 2: lines end at the slab (block Jacobi across ranks; tests / bench comparison)
    }
    c->launches += 3;
    CU(cudaMemcpyAsync(c->h_nactive, c->d_nactive, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    int done_its = 0;
    // right-preconditioned BiCGSTAB, one V-cycle per preconditioner application; the smoother's line solves are made
    // exact across the ranks by mg_spike_fix, every residual uses the neighbours' rows (halo exchange)
    auto mg_iteration = [&]() -> int {
      double* phat = c->xl_hat;
      double* shat = c->mg_shat;
      int r = SY2D_OK;
      k_p_update2<<<g2, kBlock, 0, c->stream>>>(k, c->N);
      if ((r = mg_vcycle(c, c->p, phat))) return r;
      if ((r = slab_halo_exchange(c, phat))) return r;
      KrylovVecs kv = k;
      kv.p = phat;
      k_spmv_v2<<<g2, kBlock, 0, c->stream>>>(kv, c->N, c->ny);
      if ((r = slab_reduce(c, 1, k))) return r;
      k_s_update2<<<g2, kBlock, 0, c->stream>>>(k, c->N);
      if ((r = mg_vcycle(c, c->s, shat))) return r;
      if ((r = slab_halo_exchange(c, shat))) return r;
      k_mg_spmv_t<<<g2, kBlock, 0, c->stream>>>(k, shat, c->N, c->ny);
      if ((r = slab_reduce(c, 2, k))) return r;
      k_mg_xr<<<g2, kBlock, 0, c->stream>>>(k, phat, shat, c->N);
      return slab_reduce(c, 3, k);
    };
    if (mg && c->slab_graph && !c->profiling && c->tp->capturable() && *c->h_nactive > 0) {
      // Over NCCL the iteration - ~90 kernels and ~45 collectives (halo lines, spike tips, scalar gathers) - is recorded once
      // into a CUDA graph (NCCL's calls are capturable stream work) and replayed: without it the step is bound by the
      // host-side cost of enqueueing the collectives (4096^2 on 2 GPUs: 11 ms per iteration against 2.3 ms of kernels).
      // Every rank replays the same number of times: the polled counters derive from all-gathered scalars.
      if (!c->slab_exec) {
        cudaGraph_t graph = nullptr;
        CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
        const int rci = mg_iteration();
        cudaMemcpyAsync(c->h_nactive, c->d_nactive, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream);
        const cudaError_t ec = cudaStreamEndCapture(c->stream, &graph);
        if (rci) { if (graph) cudaGraphDestroy(graph); return rci; }
        if (ec != cudaSuccess) return fail(c, SY2D_ERR_CUDA, "capturing the slab iteration failed: %s (SY2D_SLAB_GRAPH=0 issues it call by call)", cudaGetErrorString(ec));
        CU(cudaGraphInstantiate(&c->slab_exec, graph, 0));
        cudaGraphDestroy(graph);
      }
      // iteration counts barely change from step to step: last step's count minus one without polling, then one by one
      int planned = std::min(c->mg_last_iters - 1, c->opt.maxit);
      for (; planned > 0; --planned, ++done_its) CU(cudaGraphLaunch(c->slab_exec, c->stream));
      if (done_its > 0) CU(cudaStreamSynchronize(c->stream));
      while (*c->h_nactive > 0 && done_its < c->opt.maxit + 1) {
        CU(cudaGraphLaunch(c->slab_exec, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        done_its += 1;
      }
      c->launches += (long long)done_its * kernels_per_iteration(c);
    }
    while (*c->h_nactive > 0 && done_its < c->opt.maxit + check_every) {
      for (int it = 0; it < check_every && mg; ++it)
        if ((rc = mg_iteration())) return rc;
      for (int it = 0; it < check_every && xl; ++it) {
        k_xl_sweep<0><<<gs, kSweepThreads, 0, c->stream>>>(xv, c->N);
        if ((rc = slab_halo_exchange(c, c->xl_hat))) return rc;
        k_xl_spmv_v<2><<<gc, kBlock, 0, c->stream>>>(xv, c->N);
        if ((rc = slab_reduce(c, 1, k))) return rc;
        k_xl_sweep<1><<<gs, kSweepThreads, 0, c->stream>>>(xv, c->N);
        if ((rc = slab_halo_exchange(c, c->xl_hat))) return rc;
        k_xl_spmv_t<2><<<gc, kBlock, 0, c->stream>>>(xv, c->N);
        if ((rc = slab_reduce(c, 2, k))) return rc;
        k_xl_xr<2><<<gc, kBlock, 0, c->stream>>>(xv, c->N);
        if ((rc = slab_reduce(c, 3, k))) return rc;
      }
      for (int it = 0; it < check_every && !xl && !mg; ++it) {
        k_p_update2<<<g2, kBlock, 0, c->stream>>>(k, c->N);
        if ((rc = slab_halo_exchange(c, c->p))) return rc;
        k_spmv_v2<<<g2, kBlock, 0, c->stream>>>(k, c->N, c->ny);
        if ((rc = slab_reduce(c, 1, k))) return rc;
        k_s_update2<<<g2, kBlock, 0, c->stream>>>(k, c->N);
        if ((rc = slab_halo_exchange(c, c->s))) return rc;
        k_spmv_t2<<<g2, kBlock, 0, c->stream>>>(k, c->N, c->ny);
        if ((rc = slab_reduce(c, 2, k))) return rc;
        k_xr_update2<<<g2, kBlock, 0, c->stream>>>(k, c->N);
        if ((rc = slab_reduce(c, 3, k))) return rc;
      }
      CU(cudaGetLastError());
      CU(cudaMemcpyAsync(c->h_nactive, c->d_nactive, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      CU(cudaStreamSynchronize(c->stream));
      done_its += check_every;
      c->launches += (long long)check_every * (mg ? kernels_per_iteration(c) : 8);
    }
    StepStats init;
    std::memset(&init, 0, sizeof init);
    init.fmin = 1.0e300;
    *c->h_stats = init;
    CU(cudaMemcpyAsync(c->d_stats, c->h_stats, sizeof(StepStats), cudaMemcpyHostToDevice, c->stream));
    // h_nactive comes from scalars that are identical on every rank, so all ranks take this branch together
    const bool solve_failed = (c->h_nactive[0] > 0 || c->h_nactive[1] > 0) && !budget;
    if (solve_failed) {
      k_fail_stats<<<1, 32, 0, c->stream>>>(c->scal, 1, c->d_stats);
    } else {
      // true residual max|rhs - A d| over the whole grid: own rows with the neighbours' halo rows of d, then the max over
      // the ranks (all-gathered, so the commit decision of k_finish is the same everywhere)
      if ((rc = slab_halo_exchange(c, c->x))) return rc;
      k_true_residual<<<grid_of(c), kBlock, 0, c->stream>>>(k, c->N, c->ny, &c->d_stats->resid_max);
      if ((rc = slab_gather(c, &c->d_stats->resid_max, c->d_gather, 2))) return rc;   // resid_max, resid_rel_max are adjacent
      k_slab_max<<<1, 32, 0, c->stream>>>(c->d_gather, c->nranks, &c->d_stats->resid_max);
      const dim3 gf((unsigned)((own + kBlock - 1) / kBlock), 1, 1);
      k_finish<<<gf, kBlock, 0, c->stream>>>(c->x + c->ny, c->cs + c->ny, c->f + c->ny, c->yprev + c->ny, c->ylast + c->ny, c->scal, own, c->opt.predictor, c->d_stats,
                                             budget ? 1.0e300 : 1000.0 * c->opt.tol);
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(c->h_stats, c->d_stats, sizeof(StepStats), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    const bool bad = solve_failed || (c->h_stats->n_bad > 0 && !budget);
    if (!bad) {
      c->istep += 1;
      st.steps += 1;
    }
    st.iters_total += c->h_stats->it_max;
    st.iters_sum_all += c->h_stats->it_max;
    st.iters_last = c->h_stats->it_max;
    if (!bad) c->mg_last_iters = c->h_stats->it_max;
    st.fmin = c->h_stats->fmin;                      // of this rank's rows
    st.negatives = (long long)c->h_stats->negatives;  // of this rank's rows
    st.kernel_launches = c->launches;
    st.resid_last = c->h_stats->resid_max;           // true residual, max over all ranks
    if (bad) {
      if (stats) *stats = st;
      return fail(c, SY2D_ERR_NOT_CONVERGED, "sy2d_step (slab): step %lld not committed (%s; %d iterations, true residual %.3e)",
                  c->istep + 1, solve_failed ? "BiCGSTAB did not converge" : "componentwise backward error of the true residual above 1000 x tol", c->h_stats->it_max, c->h_stats->resid_max);
    }
  }
  CU(cudaEventRecord(c->ev_call1, c->stream));
  CU(cudaEventSynchronize(c->ev_call1));
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, c->ev_call0, c->ev_call1));
  st.seconds_device = ms * 1e-3;
  if (stats) *stats = st;
  return collect_profile(c);
}
}  // namespace

static int step_per_problem(sy2d_ctx* c, int nsteps, sy2d_stats* stats, const double* h_in, double* h_out);
static int engine_of(const sy2d_ctx* c);

extern "C" {

int sy2d_last_assembly_kernel(const sy2d_ctx* c) { return c ? c->last_asm_kernel : 0; }

const char* sy2d_build_info(void) {
  return "sayram2d_b200;arch=sm_100a;cuda="
#define SY2D_STR2(x) #x
#define SY2D_STR(x) SY2D_STR2(x)
      SY2D_STR(CUDART_VERSION) ";fp64;engines=lockstep-bicgstab,cta-per-problem;precond=jacobi,xline,multigrid;assembly=tma";
}

int sy2d_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int sy2d_default_options(sy2d_options* o) {
  if (!o) return SY2D_ERR_INVALID;
  std::memset(o, 0, sizeof *o);
  o->tol = 1e-14;
  o->maxit = 20000;
  o->precond = SY2D_PRECOND_AUTO;
  o->predictor = 2;
  o->check_every = 16;
  o->use_graph = 1;
  o->engine = 0;
  return SY2D_OK;
}

const char* sy2d_last_error(const sy2d_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

}  // extern "C"

// nx: rows of the global grid.  Slab contexts (nranks > 1) store rows [i_lo, i_hi) plus two halo rows.
static int create_impl(sy2d_ctx** out, int device, int nx, int ny, int nbatch, const double* xe, const double* ye, double dt,
                       int rank, int nranks) {
  sy2d_ctx* c = nullptr;
  if (!out) return fail(c, SY2D_ERR_INVALID, "sy2d_create: out is NULL");
  *out = nullptr;
  if (nx < 1 || ny < 1) return fail(c, SY2D_ERR_INVALID, "Grid2D: edges must have size >= 2.");
  if (nbatch < 1 || nbatch > 65535) return fail(c, SY2D_ERR_INVALID, "sy2d_create: nbatch must be in [1, 65535]");
  if (!xe || !ye) return fail(c, SY2D_ERR_INVALID, "sy2d_create: NULL edge array");
  if (!(dt > 0.0)) return fail(c, SY2D_ERR_INVALID, "sy2d_create: dt must be positive");
  for (int i = 0; i < nx; ++i)
    if (!(xe[i + 1] > xe[i])) return fail(c, SY2D_ERR_INVALID, "Grid2D: x_edges must be strictly increasing at i=%d", i);
  for (int j = 0; j < ny; ++j)
    if (!(ye[j + 1] > ye[j])) return fail(c, SY2D_ERR_INVALID, "Grid2D: y_edges must be strictly increasing at j=%d", j);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(c, SY2D_ERR_CUDA, "sy2d_create: no CUDA device available (this library has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(c, SY2D_ERR_INVALID, "sy2d_create: device %d out of range (%d devices)", device, ndev);

  sy2d_ctx* ctx = new sy2d_ctx;
  ctx->device = device; ctx->nx = nx; ctx->ny = ny; ctx->nbatch = nbatch; ctx->dt = dt;
  if (nranks > 1) {
    const int base_rows = nx / nranks, extra = nx % nranks;
    ctx->slab = true; ctx->rank = rank; ctx->nranks = nranks; ctx->nx_glob = nx;
    ctx->i_lo = rank * base_rows + std::min(rank, extra);
    ctx->i_hi = ctx->i_lo + base_rows + (rank < extra ? 1 : 0);
    ctx->nx = ctx->i_hi - ctx->i_lo + 2;  // local rows incl. the two halo rows
  }
  ctx->N = (size_t)ctx->nx * ny; ctx->total = ctx->N * nbatch;
  sy2d_default_options(&ctx->opt);
  std::memset(&ctx->prof, 0, sizeof ctx->prof);
  ctx->h_xe.assign(xe, xe + nx + 1);
  ctx->h_ye.assign(ye, ye + ny + 1);
  c = ctx;
  auto bail = [&](int code) { g_create_error = ctx->err; sy2d_destroy(ctx); return code; };
#define CUB(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fail(c, SY2D_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); return bail(SY2D_ERR_CUDA); } } while (0)
  CUB(cudaSetDevice(device));
  CUB(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  CUB(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device));
  if (const char* e = std::getenv("SY2D_CTAS_PER_SM")) ctx->ctas_per_sm = std::max(1, std::min(32, std::atoi(e)));
  if (const char* e = std::getenv("SY2D_PIPE_CHUNKS")) ctx->pipe_forced = std::max(1, std::min(64, std::atoi(e)));
  if (const char* e = std::getenv("SY2D_ASM_CTAS_PER_SM")) ctx->asm_ctas_per_sm = std::max(1, std::min(4, std::atoi(e)));
  if (const char* e = std::getenv("SY2D_XLINE_CHUNK")) ctx->xl_chunk = std::max(0, std::atoi(e));
  if (const char* e = std::getenv("SY2D_XLINE_CLUSTER")) ctx->xl_cluster = std::max(0, std::min(2, std::atoi(e)));
  if (const char* e = std::getenv("SY2D_MG_CLUSTER")) ctx->mg_cluster = std::atoi(e) != 0;
  if (const char* e = std::getenv("SY2D_MG_TAIL_NY")) ctx->mg_tail_ny = std::max(0, std::atoi(e));
  if (const char* e = std::getenv("SY2D_SLAB_GRAPH")) ctx->slab_graph = std::atoi(e) != 0;
  if (const char* e = std::getenv("SY2D_HOST_IO")) ctx->host_io_direct = std::string(e) == "direct" ? 1 : 0;
  if (const char* e = std::getenv("SY2D_ASM_KERNEL")) ctx->asm_kernel = std::string(e) == "march" ? 1 : std::string(e) == "col" ? 2 : std::string(e) == "wide" ? 3 : 0;
  if (const char* e = std::getenv("SY2D_WIDE_CTAS_PER_SM")) ctx->wide_ctas_per_sm = std::max(1, std::min(2, std::atoi(e)));
  if (const char* e = std::getenv("SY2D_DETERMINISTIC")) ctx->deterministic = std::atoi(e) != 0;
  if (const char* e = std::getenv("SY2D_MG_FUSE")) ctx->mg_fuse = std::atoi(e) != 0;
  if (const char* e = std::getenv("SY2D_MG_LINE_PRE")) ctx->mg_line_pre = std::atoi(e) != 0;
  if (const char* e = std::getenv("SY2D_COL_EDGE_WEIGHT")) ctx->col_edge_weight = std::max(0.25, std::min(8.0, std::atof(e)));
  if (const char* e = std::getenv("SY2D_MARCH_CTAS_PER_SM")) ctx->march_ctas_per_sm = std::max(1, std::min(16, std::atoi(e)));
  const HostGeometry hg = make_host_geometry(nx, ny, xe, ye);
  const std::vector<double>&wxL = hg.wxL, &wxR = hg.wxR, &wyB = hg.wyB, &wyT = hg.wyT, &dx = hg.dx, &dy = hg.dy;
  struct Up { double** dst; const std::vector<double>* src; } ups[] = {
      {&ctx->d_wxL, &wxL}, {&ctx->d_wxR, &wxR}, {&ctx->d_wyB, &wyB}, {&ctx->d_wyT, &wyT}, {&ctx->d_dx, &dx}, {&ctx->d_dy, &dy}};
  for (auto& u : ups) {
    CUB(cudaMalloc(reinterpret_cast<void**>(u.dst), u.src->size() * sizeof(double)));
    CUB(cudaMemcpy(*u.dst, u.src->data(), u.src->size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  for (int k = 0; k < 4; ++k) {
    const size_t n = (k < 2 ? ny : nx) + 1;
    CUB(cudaMalloc(reinterpret_cast<void**>(&ctx->d_bc[k]), n * sizeof(double)));
    CUB(cudaMemset(ctx->d_bc[k], 0, n * sizeof(double)));
  }
  double** fields[] = {&ctx->tx, &ctx->ty, &ctx->cxy, &ctx->U, &ctx->Ud, &ctx->f, &ctx->yprev, &ctx->ylast, &ctx->cs, &ctx->wW, &ctx->wE,
                       &ctx->wS, &ctx->wN, &ctx->rhs, &ctx->x, &ctx->r, &ctx->p, &ctx->v, &ctx->s, &ctx->t};
  for (double** fp : fields) {
    CUB(cudaMalloc(reinterpret_cast<void**>(fp), ctx->total * sizeof(double)));
    if (ctx->slab) CUB(cudaMemset(*fp, 0, ctx->total * sizeof(double)));  // halo rows outside the domain stay finite
  }
  CUB(cudaMalloc(reinterpret_cast<void**>(&ctx->scal), nbatch * sizeof(Scal)));
  CUB(cudaMemset(ctx->scal, 0, nbatch * sizeof(Scal)));
  // slots of the deterministic cross-CTA sums: one per CTA a kernel can launch for a problem - ceil(N / 256) for the
  // one-thread-per-cell kernels, the per-problem share of the capped grids (at most 32 CTAs per SM) for the others
  ctx->part_stride = (size_t)kPartSlot * (std::max<size_t>((ctx->N + kBlock - 1) / kBlock, ((size_t)ctx->sm_count * 32 + nbatch - 1) / nbatch) + 8);
  CUB(cudaMalloc(reinterpret_cast<void**>(&ctx->part), (size_t)nbatch * ctx->part_stride * sizeof(double)));
  CUB(cudaMemset(ctx->part, 0, (size_t)nbatch * ctx->part_stride * sizeof(double)));
  CUB(cudaMalloc(reinterpret_cast<void**>(&ctx->d_nactive), 2 * sizeof(int)));
  CUB(cudaMallocHost(reinterpret_cast<void**>(&ctx->h_nactive), 2 * sizeof(int)));
  CUB(cudaMalloc(reinterpret_cast<void**>(&ctx->d_stats), sizeof(StepStats)));
  CUB(cudaMallocHost(reinterpret_cast<void**>(&ctx->h_stats), sizeof(StepStats)));
  CUB(cudaEventCreate(&ctx->ev_call0));
  CUB(cudaEventCreate(&ctx->ev_call1));
  ctx->have_tma = tma_build_maps(ctx);
#undef CUB
  *out = ctx;
  return SY2D_OK;
}

extern "C" {

int sy2d_create(sy2d_ctx** out, int device, int nx, int ny, int nbatch, const double* xe, const double* ye, double dt) {
  return create_impl(out, device, nx, ny, nbatch, xe, ye, dt, 0, 1);
}

void sy2d_destroy(sy2d_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->chunk_exec) cudaGraphExecDestroy(c->chunk_exec);
  if (c->one_exec) cudaGraphExecDestroy(c->one_exec);
  if (c->slab_exec) cudaGraphExecDestroy(c->slab_exec);
  double* bufs[] = {c->d_wxL, c->d_wxR, c->d_wyB, c->d_wyT, c->d_dx, c->d_dy, c->d_bc[0], c->d_bc[1], c->d_bc[2], c->d_bc[3],
                    c->tx, c->ty, c->cxy, c->U, c->Ud, c->f, c->yprev, c->ylast, c->cs, c->wW, c->wE, c->wS, c->wN, c->rhs,
                    c->x, c->r, c->p, c->v, c->s, c->t, c->xl_scratch, c->xl_l, c->xl_dinv, c->xl_e, c->xl_hat};
  for (double* b : bufs) if (b) cudaFree(b);
  if (c->scal) cudaFree(c->scal);
  if (c->part) cudaFree(c->part);
  for (double* b : c->mg_bufs) cudaFree(b);
  if (c->mg_tail_ctr) cudaFree(c->mg_tail_ctr);
  if (c->d_tma_maps) cudaFree(c->d_tma_maps);
  if (c->d_col_runs) cudaFree(c->d_col_runs);
  if (c->d_tma_maps2) cudaFree(c->d_tma_maps2);
  if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
  if (c->coeffs_ready) cudaEventDestroy(c->coeffs_ready);
  if (c->bc_ready) cudaEventDestroy(c->bc_ready);
  {
    double* more[] = {c->tx2, c->ty2, c->cxy2, c->U2, c->Ud2, c->raw_dev[0], c->raw_dev[1], c->raw_dev[2], c->raw_dev[3], c->raw_dev[4],
                      c->d_bc2[0], c->d_bc2[1], c->d_bc2[2], c->d_bc2[3]};
    for (double* b : more) if (b) cudaFree(b);
    for (double* b : c->raw_pin) if (b) cudaFreeHost(b);
    if (c->bc_pin) cudaFreeHost(c->bc_pin);
  }
  for (cudaStream_t sk : c->pipe_streams) if (sk) { cudaStreamSynchronize(sk); cudaStreamDestroy(sk); }
  for (cudaEvent_t ek : c->pipe_events) if (ek) cudaEventDestroy(ek);
  if (c->pipe_start) cudaEventDestroy(c->pipe_start);
  if (c->d_order) cudaFree(c->d_order);
  if (c->d_cost) cudaFree(c->d_cost);
  if (c->d_qctl) cudaFree(c->d_qctl);
  if (c->d_slots) cudaFree(c->d_slots);
  if (c->d_steps_done) cudaFree(c->d_steps_done);
  if (c->d_gather) cudaFree(c->d_gather);
  slab_comm_destroy(c);
  if (c->d_nactive) cudaFree(c->d_nactive);
  if (c->h_nactive) cudaFreeHost(c->h_nactive);
  if (c->d_stats) cudaFree(c->d_stats);
  if (c->h_stats) cudaFreeHost(c->h_stats);
  for (auto& e : c->ev_pool) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
  if (c->ev_call0) cudaEventDestroy(c->ev_call0);
  if (c->ev_call1) cudaEventDestroy(c->ev_call1);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int sy2d_set_options(sy2d_ctx* c, const sy2d_options* o) {
  if (!c || !o) return SY2D_ERR_INVALID;
  if (!(o->tol > 0.0) || o->maxit < 1 || o->check_every < 1)
    return fail(c, SY2D_ERR_INVALID, "sy2d_set_options: tol, maxit and check_every must be positive");
  if (o->precond < SY2D_PRECOND_AUTO || o->precond > SY2D_PRECOND_MG)
    return fail(c, SY2D_ERR_INVALID, "sy2d_set_options: unknown preconditioner %d", o->precond);
  // the captured iteration graphs hold tol and maxit by value (KrylovVecs is a kernel argument): rebuild them on change
  if (o->tol != c->opt.tol || o->maxit != c->opt.maxit) {
    if (c->chunk_exec) { cudaGraphExecDestroy(c->chunk_exec); c->chunk_exec = nullptr; }
    if (c->one_exec) { cudaGraphExecDestroy(c->one_exec); c->one_exec = nullptr; }
    if (c->slab_exec) { cudaGraphExecDestroy(c->slab_exec); c->slab_exec = nullptr; }
  }
  if (o->precond != c->opt.precond || o->mg_levels != c->opt.mg_levels || o->mg_coarse_sweeps != c->opt.mg_coarse_sweeps ||
      o->reserved[2] != c->opt.reserved[2]) {
    if (c->slab_exec) { cudaGraphExecDestroy(c->slab_exec); c->slab_exec = nullptr; }
  }
  c->opt = *o;
  if (c->slab) c->opt.use_graph = 0;
  return SY2D_OK;
}

int sy2d_set_coeffs_dev(sy2d_ctx* c, const double* G, const double* Dxx, const double* Dxy, const double* Dyy, const double* inv_tau) {
  if (!c) return SY2D_ERR_INVALID;
  if (!G || !Dxx || !Dxy || !Dyy) return fail(c, SY2D_ERR_INVALID, "sy2d_set_coeffs: NULL field");
  CU(cudaSetDevice(c->device));
  {
    std::lock_guard<std::mutex> lk(c->stage_mu);   // a synchronous set supersedes fields staged asynchronously before it
    c->coeffs_pending = false;
  }
  if (c->slab) {  // inputs are the owned rows [i_lo, i_hi); outputs go to local rows 1..; dx is indexed globally
    const int rows = c->nx - 2, off = c->ny;
    const dim3 g((unsigned)(((size_t)rows * c->ny + kBlock - 1) / kBlock), 1, 1);
    k_prepare_coeffs<<<g, kBlock, 0, c->stream>>>(G, Dxx, Dxy, Dyy, inv_tau, c->d_dx + c->i_lo, c->d_dy, c->dt, rows, c->ny,
                                                  c->tx + off, c->ty + off, c->cxy + off, c->U + off, c->Ud + off);
    CU(cudaGetLastError());
    int rc = slab_halo_exchange(c, c->tx);
    if (!rc) rc = slab_halo_exchange(c, c->cxy);
    if (rc) return rc;
  } else {
    k_prepare_coeffs<<<grid_of(c), kBlock, 0, c->stream>>>(G, Dxx, Dxy, Dyy, inv_tau, c->d_dx, c->d_dy, c->dt, c->nx, c->ny,
                                                         c->tx, c->ty, c->cxy, c->U, c->Ud);
    CU(cudaGetLastError());
  }
  CU(cudaStreamSynchronize(c->stream));
  c->have_coeffs = true;
  return SY2D_OK;
}

int sy2d_set_coeffs(sy2d_ctx* c, const double* G, const double* Dxx, const double* Dxy, const double* Dyy, const double* inv_tau) {
  if (!c) return SY2D_ERR_INVALID;
  if (!G || !Dxx || !Dxy || !Dyy) return fail(c, SY2D_ERR_INVALID, "sy2d_set_coeffs: NULL field");
  CU(cudaSetDevice(c->device));
  // stage through the (not yet needed) Krylov vectors: x, r, p, v, s hold the five raw fields
  double* dst[5] = {c->x, c->r, c->p, c->v, c->s};
  const double* src[5] = {G, Dxx, Dxy, Dyy, inv_tau};
  for (int k = 0; k < 5; ++k)
    if (src[k]) CU(cudaMemcpyAsync(dst[k], src[k], own_elems(c) * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  return sy2d_set_coeffs_dev(c, dst[0], dst[1], dst[2], dst[3], inv_tau ? dst[4] : nullptr);
}

}  // extern "C"

// ---- asynchronous staging of the next step's fields (Solver.cc:286-289 without stalling the step in flight) ----
static int stage_alloc(sy2d_ctx* c) {
  if (c->copy_stream) return SY2D_OK;
  CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&c->coeffs_ready, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&c->bc_ready, cudaEventDisableTiming));
  double** dev[] = {&c->tx2, &c->ty2, &c->cxy2, &c->U2, &c->Ud2, &c->raw_dev[0], &c->raw_dev[1], &c->raw_dev[2], &c->raw_dev[3], &c->raw_dev[4]};
  for (double** d : dev) CU(cudaMalloc(reinterpret_cast<void**>(d), c->total * sizeof(double)));
  for (double*& h : c->raw_pin) CU(cudaMallocHost(reinterpret_cast<void**>(&h), c->total * sizeof(double)));
  for (int k = 0; k < 4; ++k) {
    const size_t n = (k < 2 ? c->ny : c->nx) + 1;
    CU(cudaMalloc(reinterpret_cast<void**>(&c->d_bc2[k]), n * sizeof(double)));
    CU(cudaMemset(c->d_bc2[k], 0, n * sizeof(double)));
  }
  CU(cudaMallocHost(reinterpret_cast<void**>(&c->bc_pin), (size_t)(2 * (c->ny + 1) + 2 * (c->nx + 1)) * sizeof(double)));
  if (c->have_tma) {
    AsmMaps m2;
    if (!tma_encode_maps(c, &m2, c->tx2, c->ty2, c->cxy2, c->U2, c->Ud2)) return fail(c, SY2D_ERR_CUDA, "cuTensorMapEncodeTiled failed for the second coefficient set");
    CU(cudaMalloc(reinterpret_cast<void**>(&c->d_tma_maps2), sizeof(AsmMaps)));
    CU(cudaMemcpy(c->d_tma_maps2, &m2, sizeof(AsmMaps), cudaMemcpyHostToDevice));
  }
  return SY2D_OK;
}

// Start of a time step: fields staged asynchronously since the last step become the active set (the compute stream
// waits for their upload; the host does not).
static int stage_swap_in(sy2d_ctx* c) {
  std::lock_guard<std::mutex> lk(c->stage_mu);
  if (c->coeffs_pending) {
    CU(cudaStreamWaitEvent(c->stream, c->coeffs_ready, 0));
    std::swap(c->tx, c->tx2); std::swap(c->ty, c->ty2); std::swap(c->cxy, c->cxy2); std::swap(c->U, c->U2); std::swap(c->Ud, c->Ud2);
    std::swap(c->d_tma_maps, c->d_tma_maps2);
    c->coeffs_pending = false;
    c->have_coeffs = true;
    c->swaps += 1;
  }
  if (c->bc_pending) {
    CU(cudaStreamWaitEvent(c->stream, c->bc_ready, 0));
    for (int k = 0; k < 4; ++k) { std::swap(c->d_bc[k], c->d_bc2[k]); std::swap(c->bc[k], c->bc2[k]); }
    c->bc_pending = false;
    c->have_bc = true;
    c->swaps += 1;
  }
  c->steps_begun += 1;
  return SY2D_OK;
}

extern "C" {

int sy2d_set_coeffs_async(sy2d_ctx* c, const double* G, const double* Dxx, const double* Dxy, const double* Dyy, const double* inv_tau) {
  if (!c) return SY2D_ERR_INVALID;
  if (!G || !Dxx || !Dxy || !Dyy) return fail(c, SY2D_ERR_INVALID, "sy2d_set_coeffs_async: NULL field");
  if (c->slab) return fail(c, SY2D_ERR_INVALID, "sy2d_set_coeffs_async: not available on a slab context (use sy2d_set_coeffs)");
  CU(cudaSetDevice(c->device));
  std::lock_guard<std::mutex> lk(c->stage_mu);
  int rc = stage_alloc(c);
  if (rc) return rc;
  // a set staged earlier and not yet swapped in is overwritten (the last call before a step wins); its upload must be over
  // before the pinned staging is reused
  CU(cudaStreamSynchronize(c->copy_stream));
  const double* src[5] = {G, Dxx, Dxy, Dyy, inv_tau};
  for (int k = 0; k < 5; ++k) {
    if (!src[k]) continue;
    std::memcpy(c->raw_pin[k], src[k], c->total * sizeof(double));   // the caller's arrays are free again on return
    CU(cudaMemcpyAsync(c->raw_dev[k], c->raw_pin[k], c->total * sizeof(double), cudaMemcpyHostToDevice, c->copy_stream));
  }
  k_prepare_coeffs<<<grid_of(c), kBlock, 0, c->copy_stream>>>(c->raw_dev[0], c->raw_dev[1], c->raw_dev[2], c->raw_dev[3], inv_tau ? c->raw_dev[4] : nullptr,
                                                             c->d_dx, c->d_dy, c->dt, c->nx, c->ny, c->tx2, c->ty2, c->cxy2, c->U2, c->Ud2);
  CU(cudaGetLastError());
  CU(cudaEventRecord(c->coeffs_ready, c->copy_stream));
  c->coeffs_pending = true;
  return SY2D_OK;
}

int sy2d_set_bc_async(sy2d_ctx* c, const int bc_type[4], const double* xmin, const double* xmax, const double* ymin, const double* ymax) {
  if (!c || !bc_type) return SY2D_ERR_INVALID;
  if (c->slab) return fail(c, SY2D_ERR_INVALID, "sy2d_set_bc_async: not available on a slab context (use sy2d_set_bc)");
  const double* lines[4] = {xmin, xmax, ymin, ymax};
  for (int k = 0; k < 4; ++k) {
    if (bc_type[k] != SY2D_DIRICHLET && bc_type[k] != SY2D_ZEROFLUX) return fail(c, SY2D_ERR_INVALID, "sy2d_set_bc_async: unknown BCType %d", bc_type[k]);
    if (bc_type[k] == SY2D_DIRICHLET && !lines[k]) return fail(c, SY2D_ERR_BC, "Dirichlet BC: missing value.");
  }
  CU(cudaSetDevice(c->device));
  std::lock_guard<std::mutex> lk(c->stage_mu);
  int rc = stage_alloc(c);
  if (rc) return rc;
  CU(cudaStreamSynchronize(c->copy_stream));
  double* pin = c->bc_pin;
  for (int k = 0; k < 4; ++k) {
    const size_t n = (k < 2 ? c->ny : c->nx) + 1;
    c->bc2[k] = bc_type[k];
    if (lines[k]) {
      std::memcpy(pin, lines[k], n * sizeof(double));
      CU(cudaMemcpyAsync(c->d_bc2[k], pin, n * sizeof(double), cudaMemcpyHostToDevice, c->copy_stream));
    }
    pin += n;
  }
  CU(cudaEventRecord(c->bc_ready, c->copy_stream));
  c->bc_pending = true;
  return SY2D_OK;
}

long long sy2d_stage_swaps(const sy2d_ctx* c) { return c ? c->swaps : 0; }

long long sy2d_steps_begun(sy2d_ctx* c) {
  if (!c) return 0;
  std::lock_guard<std::mutex> lk(c->stage_mu);
  return c->steps_begun;
}

int sy2d_set_bc(sy2d_ctx* c, const int bc_type[4], const double* xmin, const double* xmax, const double* ymin, const double* ymax) {
  if (!c || !bc_type) return SY2D_ERR_INVALID;
  const double* lines[4] = {xmin, xmax, ymin, ymax};
  for (int k = 0; k < 4; ++k) {
    if (bc_type[k] != SY2D_DIRICHLET && bc_type[k] != SY2D_ZEROFLUX) return fail(c, SY2D_ERR_INVALID, "sy2d_set_bc: unknown BCType %d", bc_type[k]);
    if (bc_type[k] == SY2D_DIRICHLET && !lines[k]) return fail(c, SY2D_ERR_BC, "Dirichlet BC: missing value.");
  }
  CU(cudaSetDevice(c->device));
  {
    std::lock_guard<std::mutex> lk(c->stage_mu);
    c->bc_pending = false;
  }
  for (int k = 0; k < 4; ++k) {
    c->bc[k] = bc_type[k];
    const size_t n = (k < 2 ? c->ny : (c->slab ? c->nx_glob : c->nx)) + 1;
    if (lines[k]) CU(cudaMemcpyAsync(c->d_bc[k], lines[k], n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  }
  CU(cudaStreamSynchronize(c->stream));
  c->have_bc = true;
  return SY2D_OK;
}

// f must be finite and > 0 in every cell (k_check_f); on failure the context has no usable f
static int check_f_input(sy2d_ctx* c, const char* who) {
  unsigned long long* d_bad = &c->d_stats->negatives;
  CU(cudaMemsetAsync(d_bad, 0, sizeof(unsigned long long), c->stream));
  const size_t n = own_elems(c);
  k_check_f<<<(unsigned)std::min<size_t>((n + kBlock - 1) / kBlock, (size_t)c->sm_count * 8), kBlock, 0, c->stream>>>(own_ptr(c, c->f), n, d_bad);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(&c->h_stats->negatives, d_bad, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (c->h_stats->negatives > 0) {
    c->have_f = false;
    return fail(c, SY2D_ERR_INVALID, "%s: f must be finite and > 0 in every cell (%llu cells are not): the engine solves for the per-cell ratio "
                "f^{n+1}/f^n; the reference's cases add gEPS to f0 for the same reason (Albert_Young.h:39)", who, c->h_stats->negatives);
  }
  return SY2D_OK;
}

static int reset_state(sy2d_ctx* c) {
  // yprev = 1: the first step column-scales by f^n alone
  k_fill<<<grid_of(c), kBlock, 0, c->stream>>>(c->yprev, c->N, 1.0);
  k_fill<<<grid_of(c), kBlock, 0, c->stream>>>(c->ylast, c->N, 0.0);   // no ratio history yet (predictor 2 extrapolates from the second step on)
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(c->stream));
  c->istep = 0;
  return SY2D_OK;
}

int sy2d_set_f_dev(sy2d_ctx* c, const double* f) {
  if (!c || !f) return SY2D_ERR_INVALID;
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpyAsync(own_ptr(c, c->f), f, own_elems(c) * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  int rc = check_f_input(c, "sy2d_set_f_dev");
  if (rc) return rc;
  c->have_f = true;
  return reset_state(c);
}

int sy2d_set_f(sy2d_ctx* c, const double* f) {
  if (!c || !f) return SY2D_ERR_INVALID;
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpyAsync(own_ptr(c, c->f), f, own_elems(c) * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  int rc = check_f_input(c, "sy2d_set_f");
  if (rc) return rc;
  c->have_f = true;
  return reset_state(c);
}

int sy2d_put_f(sy2d_ctx* c, const double* f) {
  if (!c || !f) return SY2D_ERR_INVALID;
  if (!c->have_f) return sy2d_set_f(c, f);
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpyAsync(own_ptr(c, c->f), f, own_elems(c) * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  return check_f_input(c, "sy2d_put_f");   // synchronises: the caller's buffer is free again on return
}

int sy2d_get_f(sy2d_ctx* c, double* out) {
  if (!c || !out) return SY2D_ERR_INVALID;
  if (!c->have_f) return fail(c, SY2D_ERR_STATE, "sy2d_get_f: f not set");
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpyAsync(out, own_ptr(c, c->f), own_elems(c) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return SY2D_OK;
}

int sy2d_get_f_dev(sy2d_ctx* c, double* out) {
  if (!c || !out) return SY2D_ERR_INVALID;
  if (!c->have_f) return fail(c, SY2D_ERR_STATE, "sy2d_get_f: f not set");
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpyAsync(out, own_ptr(c, c->f), own_elems(c) * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return SY2D_OK;
}

double sy2d_time(const sy2d_ctx* c) { return c ? (double)c->istep * c->dt : 0.0; }
long long sy2d_step_count(const sy2d_ctx* c) { return c ? c->istep : 0; }

static int ready(sy2d_ctx* c, const char* who) {
  if (!c->have_coeffs) return fail(c, SY2D_ERR_STATE, "%s: coefficients not set (sy2d_set_coeffs)", who);
  if (!c->have_bc) return fail(c, SY2D_ERR_STATE, "%s: boundary conditions not set (sy2d_set_bc)", who);
  if (!c->have_f) return fail(c, SY2D_ERR_STATE, "%s: f not set (sy2d_set_f)", who);
  return SY2D_OK;
}

}  // extern "C"

static int engine_of(const sy2d_ctx* c) {
  if (c->slab) return 1;
  if (c->opt.engine == 1 || c->opt.engine == 2) return c->opt.engine;
  return c->N <= 16384 ? 2 : 1;
}

// Shape of the x-line kernel for this grid: R rows per thread (template instance), NT threads.
// Returns false when the grid does not fit (too many rows per lane, too wide, or too much shared memory).
static bool xline_shape(const sy2d_ctx* c, int* R, int* NT, int* S, int* HS, size_t* smem) {
  const int need = (c->nx + kXlineNCH - 1) / kXlineNCH;
  const int choices[4] = {4, 6, 8, 10};
  int r = 0;
  for (int q : choices) if (q >= need) { r = q; break; }
  if (r == 0) return false;
  const int ny_pad = (c->ny + kXlineCPW - 1) / kXlineCPW * kXlineCPW;
  const int nt = ny_pad * kXlineNCH;
  if (nt > 1024) return false;
  // Row stride of hat.  A 64-bit shared access is served per HALF-warp (16 lanes, 16 eight-byte slots): lanes
  // 0-15 are (jj, k) = (0..1, 0..7) and hit slots (k*r*hs + jj) mod 16, so r*hs = 2 (mod 16) makes the 16 distinct
  // (ncu: every hat access was a 2-way conflict with the earlier rule r*hs = 4, which only spreads a full warp).
  // Not solvable for r = 4 and r = 8 (r*hs is then a multiple of 4): fall back to 4 (mod 16), then to ny + 1.
  int hs = 0;
  for (int want : {2, 4}) {
    for (int t = c->ny; t < c->ny + 16 && !hs; ++t)
      if ((r * t) % 16 == want) hs = t;
    if (hs) break;
  }
  if (!hs) hs = c->ny + 1;
  const size_t bytes = ((size_t)c->nx * hs + 3 * (size_t)r * nt + 194) * sizeof(double)   /* + three 64-double reduction buffers + the work-item word */;
  if (bytes > 232448) return false;
  *R = r; *NT = nt; *S = r * nt; *HS = hs; *smem = bytes;
  return true;
}

template <int R, int MAXT, int NTC, int HSC = 0>
static cudaError_t launch_xline(const XlineArgs& xa, int nctas, size_t smem, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(k_problem_xline<R, MAXT, NTC, HSC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_problem_xline<R, MAXT, NTC, HSC><<<nctas, xa.NT, smem, stream>>>(xa);
  return cudaGetLastError();
}

// The pair kernel: two CTAs (one cluster) per problem; the cluster shape is a compile-time attribute of the kernel.
template <int R, int NCH, int NT, int HS>
static cudaError_t launch_xline_cl(const XlineArgs& xa, int nclusters, cudaStream_t stream) {
  constexpr size_t smem = ((size_t)NCH * R * HS + 6 * (size_t)R * NT + 384 + 8) * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(k_problem_xline_cl<R, NCH, NT, HS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_problem_xline_cl<R, NCH, NT, HS><<<2 * nclusters, NT, smem, stream>>>(xa);
  return cudaGetLastError();
}

// slots[0 .. nprob) = the problems of the launch in issue order, the rest "not yet pushed"; control words
__global__ void k_xline_queue_init(XlineQueue* q, int* __restrict__ slots, const int* __restrict__ order, int nprob, int nchunks) {
  const int total = nprob * nchunks;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) slots[t] = t < nprob ? order[t] : -1;
  if (blockIdx.x == 0 && threadIdx.x == 0) { q->head = 0; q->tail = nprob; q->total = total; q->pad = 0; }
}

static cudaError_t dispatch_xline(const XlineArgs& xa, int R, int nbatch, size_t smem, cudaStream_t stream) {
  if (R == 10 && xa.NT == 640 && xa.hs == 85 && xa.a.g.nx == 80 && xa.a.g.ny == 80)
    return launch_xline<10, 640, 640, 85>(xa, nbatch, xline_full_smem_doubles<10, 640, 85>() * sizeof(double), stream);   // 80 x 80: the production shape, everything compile-time
  if (R == 10 && xa.NT == 640) return launch_xline<10, 640, 640>(xa, nbatch, smem, stream);
  const bool small = xa.NT <= 640;
  switch (R) {
    case 4: return small ? launch_xline<4, 640, 0>(xa, nbatch, smem, stream) : launch_xline<4, 1024, 0>(xa, nbatch, smem, stream);
    case 6: return small ? launch_xline<6, 640, 0>(xa, nbatch, smem, stream) : launch_xline<6, 1024, 0>(xa, nbatch, smem, stream);
    case 8: return small ? launch_xline<8, 640, 0>(xa, nbatch, smem, stream) : launch_xline<8, 1024, 0>(xa, nbatch, smem, stream);
    default: return small ? launch_xline<10, 640, 0>(xa, nbatch, smem, stream) : launch_xline<10, 1024, 0>(xa, nbatch, smem, stream);
  }
}

// Engine 2: the whole call (nsteps time steps of every problem) is ONE kernel launch.
// Number of sub-batches a host-buffer call is pipelined over (and within which problems are sorted
// by cost): one wave of CTAs per sub-batch (half a wave while that gives at most 6; the tail of a sub-batch is
// filled by the next one's CTAs, the streams run concurrently), at most 32 - what is exposed of the copies is
// the upload of the first sub-batch and the download of the last one (measured, 4096 members: 8 sub-batches
// 9.07 ms per end-to-end step, 16: 8.55, 27: 8.41, 55: 8.38 against 8.2 ms device-resident; 512 members: 3
// sub-batches 1.50 ms, 6: 1.46 against 1.13).
static int pipe_chunks(const sy2d_ctx* c) {
  if (c->pipe_forced > 0) return std::max(1, std::min(c->pipe_forced, c->nbatch));
  const int waves = c->nbatch / c->sm_count, half_waves = 2 * c->nbatch / c->sm_count;
  return std::max(1, std::min(c->pipe_max, std::max(waves, std::min(6, half_waves))));
}

// Engine 2: nsteps time steps of every problem in ONE kernel launch per sub-batch.  With host
// buffers (h_in / h_out non-NULL) the batch is cut into pipe_chunks() contiguous sub-batches on
// separate streams, so that the H2D copy of sub-batch k+1, the kernel of sub-batch k and the D2H
// copy of sub-batch k-1 overlap (and the tail wave of one kernel is filled by the next one).
// x-line kernel: persistent CTAs (one per SM) pull (problem, chunk of steps) items from a device-side queue.
static int build_issue_order(sy2d_ctx* c, int C) {
  // inside every sub-batch, most expensive problems (iterations of the last measured call) first
  for (int b = 0; b < c->nbatch; ++b) c->h_order[b] = b;
  for (int k = 0; k < C; ++k) {
    const int b0 = (int)((long long)c->nbatch * k / C), b1 = (int)((long long)c->nbatch * (k + 1) / C);
    std::stable_sort(c->h_order.begin() + b0, c->h_order.begin() + b1, [&](int x, int y) { return c->h_cost[x] > c->h_cost[y]; });
  }
  CU(cudaMemcpyAsync(c->d_order, c->h_order.data(), c->nbatch * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  c->order_C = C;
  c->order_age = 0;
  return SY2D_OK;
}

static int step_per_problem(sy2d_ctx* c, int nsteps, sy2d_stats* stats, const double* h_in, double* h_out) {
  sy2d_stats st;
  std::memset(&st, 0, sizeof st);
  st.engine = 2;
  if (nsteps == 0) { if (stats) *stats = st; return SY2D_OK; }
  // Host buffers the device can address (pinned / registered memory under UVA) are read and written by the x-line kernel
  // itself; anything else goes through the copy engines, pipelined over sub-batches.
  auto device_view = [&](const void* h) -> double* {
    if (!h) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, h) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
    return static_cast<double*>(at.devicePointer);
  };
  int R = 0, NT = 0, S = 0, HS = 0;
  size_t smem = 0;
  const bool xline = c->opt.precond != SY2D_PRECOND_JACOBI && xline_shape(c, &R, &NT, &S, &HS, &smem);
  double* d_hin = c->host_io_direct && xline ? device_view(h_in) : nullptr;
  double* d_hout = c->host_io_direct && xline ? device_view(h_out) : nullptr;
  const bool direct = (h_in || h_out) && (!h_in || d_hin) && (!h_out || d_hout) && !c->profiling;
  if (!direct) d_hin = d_hout = nullptr;
  const bool piped = (h_in || h_out) && !c->profiling && !direct;
  const int C = piped ? pipe_chunks(c) : 1;
  StepStats init;
  std::memset(&init, 0, sizeof init);
  init.fmin = 1.0e300;
  init.steps_min = nsteps;
  *c->h_stats = init;
  CU(cudaEventRecord(c->ev_call0, c->stream));
  CU(cudaMemcpyAsync(c->d_stats, c->h_stats, sizeof(StepStats), cudaMemcpyHostToDevice, c->stream));
  ProblemArgs a;
  a.tx = c->tx; a.ty = c->ty; a.cxy = c->cxy; a.U = c->U; a.Ud = c->Ud;
  a.f = c->f; a.yprev = c->yprev; a.ylast = c->ylast; a.cs = c->cs;
  a.wW = c->wW; a.wE = c->wE; a.wS = c->wS; a.wN = c->wN; a.rhs = c->rhs;
  a.x = c->x; a.r = c->r; a.p = c->p; a.v = c->v; a.s = c->s; a.t = c->t;
  a.scal = c->scal; a.stats = c->d_stats; a.g = geometry(c);
  a.tol = c->opt.tol; a.maxit = c->opt.maxit; a.predictor = c->opt.predictor; a.nsteps = nsteps;
  if (!c->d_order) {
    CU(cudaMalloc(reinterpret_cast<void**>(&c->d_order), c->nbatch * sizeof(int)));
    CU(cudaMalloc(reinterpret_cast<void**>(&c->d_cost), c->nbatch * sizeof(int)));
    CU(cudaMalloc(reinterpret_cast<void**>(&c->d_steps_done), c->nbatch * sizeof(int)));
    CU(cudaMalloc(reinterpret_cast<void**>(&c->d_qctl), 64 * sizeof(XlineQueue)));
    c->h_cost.assign(c->nbatch, 0);
    c->h_order.resize(c->nbatch);
    c->order_C = -1;
  }
  // The issue order is rebuilt from the measured costs on the first calls, when the sub-batch partition changes (the
  // order must keep every problem inside its own sub-batch: the copies are per sub-batch) and every 16th call after that.
  if (c->order_C != C || c->order_age >= 16 || c->order_calls < 2) { int rc0 = build_issue_order(c, C); if (rc0) return rc0; }
  c->order_age += 1;
  c->order_calls += 1;
  a.cost = c->d_cost;
  CU(cudaMemsetAsync(c->d_cost, 0, c->nbatch * sizeof(int), c->stream));
  c->cur_cells = (double)c->total * nsteps;
  if (c->opt.precond == SY2D_PRECOND_XLINE && !xline)
    return fail(c, SY2D_ERR_INVALID, "sy2d_step: the x-line preconditioner needs nx <= 80, ny <= 128 and engine 2");
  st.precond = xline ? SY2D_PRECOND_XLINE : SY2D_PRECOND_JACOBI;
  // x-line kernel: work items of `chunk` time steps (default 1; SY2D_XLINE_CHUNK=0: the whole call, i.e. one CTA per problem)
  const int chunk = c->xl_chunk > 0 ? std::min(c->xl_chunk, nsteps) : nsteps;
  const int nchunks = (nsteps + chunk - 1) / chunk;
  // co-resident CTAs of a launch: one per SM when the CTA needs more than half of an SM's shared memory (always for the
  // production shape), else what fits; never more than the problems of the launch
  const int ctas_per_sm = smem > 113 * 1024 ? 1 : (int)std::min<size_t>(2, (227 * 1024) / std::max<size_t>(smem, 1));
  if (xline) {
    const size_t scratch_slots = (size_t)std::min(c->nbatch, C * c->sm_count * ctas_per_sm);
    if (!c->xl_scratch || c->xl_S != S || c->xl_scratch_slots < scratch_slots) {
      if (c->xl_scratch) cudaFree(c->xl_scratch);
      c->xl_scratch = nullptr;
      CU(cudaMalloc(reinterpret_cast<void**>(&c->xl_scratch), scratch_slots * kXlineScratchArrays * S * sizeof(double)));
      c->xl_S = S;
      c->xl_scratch_slots = scratch_slots;
    }
    const size_t need = (size_t)c->nbatch * nchunks;
    if (c->slots_cap < need) {
      if (c->d_slots) cudaFree(c->d_slots);
      c->d_slots = nullptr;
      CU(cudaMalloc(reinterpret_cast<void**>(&c->d_slots), need * sizeof(int)));
      c->slots_cap = need;
    }
    CU(cudaMemsetAsync(c->d_steps_done, 0, c->nbatch * sizeof(int), c->stream));
  }
  int scratch_next = 0;   // first free scratch slot (launches of one call use disjoint slots)
  auto launch = [&](int k, int b0, int b1, cudaStream_t stream) -> cudaError_t {
    ProblemArgs aa = a;
    if (xline) {
      const int nprob = b1 - b0;
      const int nctas = std::min(nprob, c->sm_count * ctas_per_sm);
      XlineArgs xa;
      xa.a = aa; xa.NT = NT; xa.S = S; xa.hs = HS;
      xa.scratch = c->xl_scratch + (size_t)scratch_next * kXlineScratchArrays * S;
      scratch_next += nctas;
      xa.q = c->d_qctl + k;
      xa.slots = c->d_slots + (size_t)b0 * nchunks;
      xa.steps_done = c->d_steps_done;
      xa.chunk = chunk; xa.nchunks = nchunks;
      xa.hin = d_hin; xa.hout = d_hout;
      k_xline_queue_init<<<std::max(1, std::min(64, (nprob * nchunks + 255) / 256)), 256, 0, stream>>>(xa.q, xa.slots, c->d_order + b0, nprob, nchunks);
      if (c->xl_cluster && c->nx == 80 && c->ny == 80)
        return c->xl_cluster == 2 ? launch_xline_cl<5, 16, 640, 43>(xa, std::min(nprob, c->sm_count / 2), stream)
                                  : launch_xline_cl<10, 8, 320, 45>(xa, std::min(nprob, c->sm_count / 2), stream);
      return dispatch_xline(xa, R, nctas, smem, stream);
    }
    aa.order = c->d_order + b0;   // CTA b of this launch works on problem order[b0 + b]
    k_problem_steps<<<b1 - b0, kProblemThreads, 0, stream>>>(aa);
    return cudaGetLastError();
  };
  if (!piped) {
    if (h_in && !direct) CU(cudaMemcpyAsync(c->f, h_in, c->total * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    {
      Prof p(c, SY2D_K_PROBLEM_STEPS);
      CU(launch(0, 0, c->nbatch, c->stream));
    }
    if (h_out && !direct) CU(cudaMemcpyAsync(h_out, c->f, c->total * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    st.kernel_launches = xline ? 2 : 1;
  } else {
    if ((int)c->pipe_streams.size() < C) {
      const size_t have = c->pipe_streams.size();
      c->pipe_streams.resize(C);
      c->pipe_events.resize(C);
      for (size_t k = have; k < (size_t)C; ++k) {
        CU(cudaStreamCreateWithFlags(&c->pipe_streams[k], cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&c->pipe_events[k], cudaEventDisableTiming));
      }
      if (!c->pipe_start) CU(cudaEventCreateWithFlags(&c->pipe_start, cudaEventDisableTiming));
    }
    CU(cudaEventRecord(c->pipe_start, c->stream));  // stats initialised, order uploaded
    for (int k = 0; k < C; ++k) {
      const int b0 = (int)((long long)c->nbatch * k / C), b1 = (int)((long long)c->nbatch * (k + 1) / C);
      const size_t off = (size_t)b0 * c->N, bytes = (size_t)(b1 - b0) * c->N * sizeof(double);
      cudaStream_t sk = c->pipe_streams[k];
      CU(cudaStreamWaitEvent(sk, c->pipe_start, 0));
      if (h_in) CU(cudaMemcpyAsync(c->f + off, h_in + off, bytes, cudaMemcpyHostToDevice, sk));
      CU(launch(k, b0, b1, sk));
      if (h_out) CU(cudaMemcpyAsync(h_out + off, c->f + off, bytes, cudaMemcpyDeviceToHost, sk));
      CU(cudaEventRecord(c->pipe_events[k], sk));
    }
    for (int k = 0; k < C; ++k) CU(cudaStreamWaitEvent(c->stream, c->pipe_events[k], 0));
    st.kernel_launches = (xline ? 2 : 1) * C;
  }
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(c->h_stats, c->d_stats, sizeof(StepStats), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaEventRecord(c->ev_call1, c->stream));
  if (c->nbatch > 1) CU(cudaMemcpyAsync(c->h_cost.data(), c->d_cost, c->nbatch * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, c->ev_call0, c->ev_call1));
  int rc = collect_profile(c);
  if (rc) return rc;
  const StepStats& h = *c->h_stats;
  // a failed problem keeps the f of the last time step it completed (nothing of the failing step is committed) while the
  // others finish the call: the context's clock advances by what EVERY problem completed
  const int steps_all = h.n_bad > 0 ? std::min(h.steps_min, nsteps) : nsteps;
  c->istep += steps_all;
  st.steps = steps_all;
  st.iters_total = h.it_total_max;
  st.iters_last = h.it_max;
  st.iters_sum_all = (long long)h.it_sum_all;
  st.resid_last = h.resid_max;
  st.fmin = h.fmin;
  st.negatives = (long long)h.negatives;
  st.seconds_device = ms * 1e-3;
  if (stats) *stats = st;
  if (h.n_bad > 0)
    return fail(c, SY2D_ERR_NOT_CONVERGED, "sy2d_step: BiCGSTAB did not converge (%d problems, up to %d iterations in a step; every problem "
                "completed %d of %d time steps, a failed problem keeps the f of its last completed step)", h.n_bad, h.it_max, steps_all, nsteps);
  if (!(h.resid_rel_max <= 1000.0 * c->opt.tol))
    return fail(c, SY2D_ERR_NOT_CONVERGED, "sy2d_step: true residual after the last step (%.3e, componentwise backward error %.3e) exceeds 1000 x tol",
                h.resid_max, h.resid_rel_max);
  return SY2D_OK;
}

extern "C" {

int sy2d_step(sy2d_ctx* c, int nsteps, sy2d_stats* stats) {
  if (!c || nsteps < 0) return SY2D_ERR_INVALID;
  CU(cudaSetDevice(c->device));
  int rc = stage_swap_in(c);   // fields staged with sy2d_set_coeffs_async / sy2d_set_bc_async since the last call
  if (rc) return rc;
  rc = ready(c, "sy2d_step");
  if (rc) return rc;
  if (c->slab) return step_slab(c, nsteps, stats);
  if (engine_of(c) == 2) return step_per_problem(c, nsteps, stats, nullptr, nullptr);
  const dim3 g = grid_of(c);
  const Geometry geo = geometry(c);
  const bool graph = c->opt.use_graph && !c->profiling;
  c->launches = 0;
  sy2d_stats st;
  std::memset(&st, 0, sizeof st);
  st.engine = 1;
  bool mg = lockstep_mg(c);
  bool xl = lockstep_xline(c);
  if (c->opt.precond == SY2D_PRECOND_XLINE && !xl) return fail(c, SY2D_ERR_INVALID, "sy2d_step: the x-line preconditioner needs nx >= %d", kSeg);
  if (c->opt.precond == SY2D_PRECOND_MG && !mg)
    return fail(c, SY2D_ERR_INVALID, "sy2d_step: the multigrid preconditioner needs engine 1, 8 <= nx <= 8192 (rows per rank) and ny a multiple of 4, >= 16");
  auto prepare = [&]() -> int {   // buffers and graphs of the preconditioner in force
    mg = lockstep_mg(c);
    xl = lockstep_xline(c);
    st.precond = mg ? SY2D_PRECOND_MG : (xl ? SY2D_PRECOND_XLINE : SY2D_PRECOND_JACOBI);
    int r = SY2D_OK;
    if (xl) r = xl_alloc(c);
    if (!r && mg) r = mg_alloc(c);
    if (!r && graph) r = build_chunk_graph(c);
    return r;
  };
  rc = prepare();
  if (rc) return rc;
  CU(cudaEventRecord(c->ev_call0, c->stream));
  for (int step = 0; step < nsteps; ++step) {
   // One attempt with the preconditioner in force; when the multigrid-preconditioned solve of an AUTO context stops
   // without converging (maxit, breakdown), the step is redone from the same f with the segmented x-line iteration
   // (f is only updated by k_finish below, after a successful solve).
   for (int attempt = 0; attempt < 2; ++attempt) {
    const int check_every = c->profiling ? 1 : effective_check_every(c);  // profiling: no zero-work launches
    CU(cudaMemsetAsync(c->d_nactive, 0, 2 * sizeof(int), c->stream));
    c->cur_cells = (double)c->total;
    {
      Prof p(c, SY2D_K_ASSEMBLY);
      AssembleOut o;
      std::memset(&o, 0, sizeof o);
      o.wW = c->wW; o.wE = c->wE; o.wS = c->wS; o.wN = c->wN; o.rhs = c->rhs; o.cs = c->cs;
      o.scal = c->scal; o.part = c->deterministic ? c->part : nullptr; o.part_stride = c->part_stride; o.n_active = c->d_nactive; o.tol = c->opt.tol; o.local_rows = c->nx;
      if (mg) o.om = c->mg_om0;
      launch_assembly(c, geo, o, c->opt.reserved[0]);  // reserved[0]: 1 forces the per-cell kernel, 2 the tiled kernel without TMA (tests)
    }
    if (xl) {  // LU of the x-line segments for this step's operator
      Prof p(c, SY2D_K_OTHER);
      const XlVecs x = xl_vecs(c);
      const int nseg = (x.nrows + kSeg - 1) / kSeg;
      k_xl_factor<<<dim3(capped_blocks(c, (size_t)nseg * c->ny, kBlock), (unsigned)c->nbatch, 1), kBlock, 0, c->stream>>>(x, c->N);
      c->launches += 1;
    }
    if (mg) mg_setup(c);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(c->h_nactive, c->d_nactive, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->launches += 1;
    int launched = 0;
    if (mg && graph && *c->h_nactive > 0) {
      // Multigrid iteration counts barely change from one time step to the next: issue last step's
      // count minus one without polling, then single iterations with a poll after each.
      int planned = std::min(c->mg_last_iters - 1, c->opt.maxit);
      c->cur_cells = (double)c->total;
      if (planned > 0) {
        for (; planned >= check_every; planned -= check_every, launched += check_every) CU(cudaGraphLaunch(c->chunk_exec, c->stream));
        for (; planned > 0; --planned, ++launched) CU(cudaGraphLaunch(c->one_exec, c->stream));
        CU(cudaStreamSynchronize(c->stream));
      }
      while (*c->h_nactive > 0 && launched < c->opt.maxit + 1) {
        CU(cudaGraphLaunch(c->one_exec, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        launched += 1;
      }
      c->launches += (long long)launched * kernels_per_iteration(c);
    }
    while (*c->h_nactive > 0 && launched < c->opt.maxit + check_every) {
      c->cur_cells = (double)*c->h_nactive * (double)c->N;
      if (graph) {
        CU(cudaGraphLaunch(c->chunk_exec, c->stream));
      } else {
        for (int it = 0; it < check_every; ++it) { rc = launch_iteration(c); if (rc) return rc; }
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(c->h_nactive, c->d_nactive, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      }
      CU(cudaStreamSynchronize(c->stream));
      launched += check_every;
      c->launches += (long long)check_every * kernels_per_iteration(c);
    }
    const bool failed = c->h_nactive[0] > 0 || c->h_nactive[1] > 0 || (attempt == 0 && c->opt.reserved[2] == 1);  // reserved[2] = 1: test hook
    if (!(failed && mg && attempt == 0 && c->opt.precond == SY2D_PRECOND_AUTO)) break;
    c->mg_off = true;
    st.restarts_total += 1;
    rc = prepare();
    if (rc) { c->mg_off = false; return rc; }
   }
   const int solved_by = st.precond;   // what solved this step
   const bool fell_back = c->mg_off;
   if (fell_back) {   // back to multigrid for the next step
     c->mg_off = false;
     rc = prepare();
     if (rc) return rc;
   }
    c->cur_cells = (double)c->total;
    // verification + finish
    StepStats init;
    std::memset(&init, 0, sizeof init);
    init.fmin = 1.0e300;
    *c->h_stats = init;
    CU(cudaMemcpyAsync(c->d_stats, c->h_stats, sizeof(StepStats), cudaMemcpyHostToDevice, c->stream));
    // A solve that stopped without converging (maxit, breakdown, NaN) commits NOTHING: f, yprev and the step counter stay
    // those of t^n, so the caller may retry with other options.  A converged solve is committed by k_finish only when its
    // TRUE residual (k_true_residual) is finite and <= 1000 tol - the recursive BiCGSTAB residual alone can drift.
    const bool solve_failed = c->h_nactive[0] > 0 || c->h_nactive[1] > 0;
    if (solve_failed) {
      k_fail_stats<<<1, 256, 0, c->stream>>>(c->scal, c->nbatch, c->d_stats);
      c->launches += 1;
    } else {
      {
        Prof p(c, SY2D_K_OTHER);
        k_true_residual<<<g, kBlock, 0, c->stream>>>(krylov(c), c->N, c->ny, &c->d_stats->resid_max);
      }
      {
        Prof p(c, SY2D_K_FINISH);
        k_finish<<<g, kBlock, 0, c->stream>>>(c->x, c->cs, c->f, c->yprev, c->ylast, c->scal, c->N, c->opt.predictor, c->d_stats, 1000.0 * c->opt.tol);
      }
      c->launches += 2;
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(c->h_stats, c->d_stats, sizeof(StepStats), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    rc = collect_profile(c);
    if (rc) return rc;
    const bool bad = solve_failed || c->h_stats->n_bad > 0;
    if (!bad) {
      c->istep += 1;
      st.steps += 1;
    }
    st.iters_total += c->h_stats->it_max;
    st.iters_sum_all += (long long)c->h_stats->it_max * c->nbatch;
    st.iters_last = c->h_stats->it_max;
    if (!fell_back && !bad) c->mg_last_iters = c->h_stats->it_max;
    st.precond = solved_by;
    st.resid_last = c->h_stats->resid_max;
    st.fmin = c->h_stats->fmin;
    st.negatives = (long long)c->h_stats->negatives;
    st.kernel_launches = c->launches;
    if (bad) {
      if (stats) *stats = st;
      if (solve_failed)
        return fail(c, SY2D_ERR_NOT_CONVERGED, "sy2d_step: BiCGSTAB did not converge at step %lld (%d problems, %d iterations); the step was not committed",
                    c->istep + 1, c->h_stats->n_bad, c->h_stats->it_max);
      return fail(c, SY2D_ERR_NOT_CONVERGED, "sy2d_step: step %lld not committed: true residual %.3e (componentwise backward error %.3e) exceeds 1000 x tol, or f is not finite",
                  c->istep + 1, c->h_stats->resid_max, c->h_stats->resid_rel_max);
    }
  }
  CU(cudaEventRecord(c->ev_call1, c->stream));
  CU(cudaEventSynchronize(c->ev_call1));
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, c->ev_call0, c->ev_call1));
  st.seconds_device = ms * 1e-3;
  rc = collect_profile(c);
  if (stats) *stats = st;
  return rc;
}

int sy2d_step_host(sy2d_ctx* c, const double* f_in, double* f_out, int nsteps, sy2d_stats* stats) {
  if (!c || nsteps < 0) return SY2D_ERR_INVALID;
  if (f_in && !c->have_f) { int rc0 = sy2d_set_f(c, f_in); if (rc0) return rc0; }
  CU(cudaSetDevice(c->device));
  int rc = stage_swap_in(c);
  if (rc) return rc;
  rc = ready(c, "sy2d_step_host");
  if (rc) return rc;
  if (!c->slab && engine_of(c) == 2) return step_per_problem(c, nsteps, stats, f_in, f_out);
  if (f_in) { rc = sy2d_put_f(c, f_in); if (rc) return rc; }
  rc = sy2d_step(c, nsteps, stats);
  if (rc) return rc;
  return f_out ? sy2d_get_f(c, f_out) : SY2D_OK;
}

int sy2d_dump_operator(sy2d_ctx* c, double* diags, double* rhs) {
  if (!c || !diags || !rhs) return SY2D_ERR_INVALID;
  int rc = ready(c, "sy2d_dump_operator");
  if (rc) return rc;
  if (c->slab) return fail(c, SY2D_ERR_INVALID, "sy2d_dump_operator: not available on a slab context");
  CU(cudaSetDevice(c->device));
  AssembleOut o;
  std::memset(&o, 0, sizeof o);
  // the Krylov vectors are free between steps: x,r,p,v,s,t receive diag,W,E,S,N,R
  o.diag = c->x; o.oW = c->r; o.oE = c->p; o.oS = c->v; o.oN = c->s; o.R = c->t;
  k_assemble<1><<<grid_of(c), kBlock, 0, c->stream>>>(c->f, c->yprev, c->tx, c->ty, c->cxy, c->U, c->Ud, geometry(c), o);
  CU(cudaGetLastError());
  double* src[5] = {c->x, c->r, c->p, c->v, c->s};
  for (int k = 0; k < 5; ++k)
    CU(cudaMemcpyAsync(diags + (size_t)k * c->total, src[k], c->total * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(rhs, c->t, c->total * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return SY2D_OK;
}

// Test hook: the SCALED unit-diagonal system A d = rhs of the current f exactly as the next sy2d_step of the lockstep
// engine would assemble it, with the assembly kernel selected by options.reserved[0]:
// w4 = [4][nbatch][nx][ny] (wW, wE, wS, wN), rhs and cs = [nbatch][nx][ny] (cs may be NULL).  Does not advance time.
int sy2d_dump_scaled_operator(sy2d_ctx* c, double* w4, double* rhs, double* cs) {
  if (!c || !w4 || !rhs) return SY2D_ERR_INVALID;
  CU(cudaSetDevice(c->device));
  int rc = ready(c, "sy2d_dump_scaled_operator");
  if (rc) return rc;
  if (c->slab) return fail(c, SY2D_ERR_INVALID, "sy2d_dump_scaled_operator: not available on a slab context");
  CU(cudaMemsetAsync(c->d_nactive, 0, 2 * sizeof(int), c->stream));
  CU(cudaMemsetAsync(c->scal, 0, c->nbatch * sizeof(Scal), c->stream));
  AssembleOut o;
  std::memset(&o, 0, sizeof o);
  o.wW = c->wW; o.wE = c->wE; o.wS = c->wS; o.wN = c->wN; o.rhs = c->rhs; o.cs = c->cs;
  o.scal = c->scal; o.part = c->deterministic ? c->part : nullptr; o.part_stride = c->part_stride; o.n_active = c->d_nactive; o.tol = -1.0;
  o.local_rows = c->nx;
  launch_assembly(c, geometry(c), o, c->opt.reserved[0]);
  CU(cudaGetLastError());
  double* src[4] = {c->wW, c->wE, c->wS, c->wN};
  for (int k = 0; k < 4; ++k)
    CU(cudaMemcpyAsync(w4 + (size_t)k * c->total, src[k], c->total * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(rhs, c->rhs, c->total * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (cs) CU(cudaMemcpyAsync(cs, c->cs, c->total * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaMemsetAsync(c->scal, 0, c->nbatch * sizeof(Scal), c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return SY2D_OK;
}

int sy2d_dump_vertex_f(sy2d_ctx* c, double* vf) {
  if (!c || !vf) return SY2D_ERR_INVALID;
  int rc = ready(c, "sy2d_dump_vertex_f");
  if (rc) return rc;
  if (c->slab) return fail(c, SY2D_ERR_INVALID, "sy2d_dump_vertex_f: not available on a slab context");
  CU(cudaSetDevice(c->device));
  const size_t nv = (size_t)(c->nx + 1) * (c->ny + 1) * c->nbatch;
  double* d_vf = nullptr;
  CU(cudaMalloc(reinterpret_cast<void**>(&d_vf), nv * sizeof(double)));
  AssembleOut o;
  std::memset(&o, 0, sizeof o);
  o.vf = d_vf;
  k_assemble<2><<<grid_of(c), kBlock, 0, c->stream>>>(c->f, c->yprev, c->tx, c->ty, c->cxy, c->U, c->Ud, geometry(c), o);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(vf, d_vf, nv * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(d_vf);
  if (e != cudaSuccess) return fail(c, SY2D_ERR_CUDA, "sy2d_dump_vertex_f: %s", cudaGetErrorString(e));
  return SY2D_OK;
}

// Test hook: assembles the scaled operator of the CURRENT f (as the next sy2d_step would), sets the
// multigrid hierarchy up and applies ONE V-cycle to r (host, [nbatch][nx][ny]) -> z.  Optionally returns
// the scaled operator (w4 = [4][nbatch][nx][ny]: wW, wE, wS, wN) and the row weights om.
int sy2d_debug_vcycle(sy2d_ctx* c, const double* r, double* z, double* w4, double* om) {
  if (!c || !r || !z) return SY2D_ERR_INVALID;
  int rc = ready(c, "sy2d_debug_vcycle");
  if (rc) return rc;
  CU(cudaSetDevice(c->device));
  const int keep = c->opt.precond;
  c->opt.precond = SY2D_PRECOND_MG;
  const bool ok = !c->slab && lockstep_mg(c);
  if (ok) rc = mg_alloc(c);
  c->opt.precond = keep;
  if (!ok) return fail(c, SY2D_ERR_INVALID, "sy2d_debug_vcycle: multigrid does not support this grid");
  if (rc) return rc;
  CU(cudaMemsetAsync(c->d_nactive, 0, 2 * sizeof(int), c->stream));
  CU(cudaMemsetAsync(c->scal, 0, c->nbatch * sizeof(Scal), c->stream));
  AssembleOut o;
  std::memset(&o, 0, sizeof o);
  o.wW = c->wW; o.wE = c->wE; o.wS = c->wS; o.wN = c->wN; o.rhs = c->rhs; o.cs = c->cs; o.om = c->mg_om0;
  o.scal = c->scal; o.part = c->deterministic ? c->part : nullptr; o.part_stride = c->part_stride; o.n_active = c->d_nactive; o.tol = -1.0;   // every problem stays active
  o.local_rows = c->nx;
  k_assemble<0><<<grid_of(c), kBlock, 0, c->stream>>>(c->f, c->yprev, c->tx, c->ty, c->cxy, c->U, c->Ud, geometry(c), o);
  CU(cudaGetLastError());
  mg_setup(c);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(c->p, r, c->total * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  rc = mg_vcycle(c, c->p, c->xl_hat);
  if (rc) return rc;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(z, c->xl_hat, c->total * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (w4) {
    double* src[4] = {c->wW, c->wE, c->wS, c->wN};
    for (int k = 0; k < 4; ++k)
      CU(cudaMemcpyAsync(w4 + (size_t)k * c->total, src[k], c->total * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  }
  if (om) CU(cudaMemcpyAsync(om, c->mg_om0, c->total * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaMemsetAsync(c->scal, 0, c->nbatch * sizeof(Scal), c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return SY2D_OK;
}

int sy2d_bench_kernel(sy2d_ctx* c, int which, int reps, double* ms_per_launch) {
  if (!c || !ms_per_launch || reps < 1) return SY2D_ERR_INVALID;
  int rc = ready(c, "sy2d_bench_kernel");
  if (rc) return rc;
  if (c->slab) return fail(c, SY2D_ERR_INVALID, "sy2d_bench_kernel: not available on a slab context");
  CU(cudaSetDevice(c->device));
  const Geometry geo = geometry(c);
  KrylovVecs k = krylov(c);
  k.freeze_state = 1;
  const dim3 g = grid_of(c);
  const bool vec2 = c->ny % 2 == 0;
  const dim3 g2(capped_blocks(c, c->N / 2, kBlock), (unsigned)c->nbatch, 1);
  AssembleOut o;
  std::memset(&o, 0, sizeof o);
  o.wW = c->wW; o.wE = c->wE; o.wS = c->wS; o.wN = c->wN; o.rhs = c->rhs; o.cs = c->cs;
  o.scal = c->scal; o.part = c->deterministic ? c->part : nullptr; o.part_stride = c->part_stride; o.n_active = c->d_nactive; o.tol = -1.0;  // tol < 0: every problem stays active
  o.local_rows = c->nx;
  auto assemble = [&]() { launch_assembly(c, geo, o, c->opt.reserved[0]); };
  auto launch = [&](int w) {
    switch (w) {
      case SY2D_K_ASSEMBLY: assemble(); break;
      case SY2D_K_P_UPDATE: if (vec2) k_p_update2<<<g2, kBlock, 0, c->stream>>>(k, c->N); else k_p_update<<<g, kBlock, 0, c->stream>>>(k, c->N); break;
      case SY2D_K_SPMV_V: if (vec2) k_spmv_v2<<<g2, kBlock, 0, c->stream>>>(k, c->N, c->ny); else k_spmv_v<<<g, kBlock, 0, c->stream>>>(k, c->N, c->ny); break;
      case SY2D_K_S_UPDATE: if (vec2) k_s_update2<<<g2, kBlock, 0, c->stream>>>(k, c->N); else k_s_update<<<g, kBlock, 0, c->stream>>>(k, c->N); break;
      case SY2D_K_SPMV_T: if (vec2) k_spmv_t2<<<g2, kBlock, 0, c->stream>>>(k, c->N, c->ny); else k_spmv_t<<<g, kBlock, 0, c->stream>>>(k, c->N, c->ny); break;
      default: if (vec2) k_xr_update2<<<g2, kBlock, 0, c->stream>>>(k, c->N); else k_xr_update<<<g, kBlock, 0, c->stream>>>(k, c->N); break;
    }
  };
  if (which < SY2D_K_ASSEMBLY || which > SY2D_K_XR_UPDATE) return fail(c, SY2D_ERR_INVALID, "sy2d_bench_kernel: unknown kernel %d", which);
  // prime: operator + one full iteration so that every vector holds finite data
  CU(cudaMemsetAsync(c->d_nactive, 0, 2 * sizeof(int), c->stream));
  CU(cudaMemsetAsync(c->scal, 0, c->nbatch * sizeof(Scal), c->stream));
  assemble();
  for (int w = SY2D_K_P_UPDATE; w <= SY2D_K_XR_UPDATE; ++w) launch(w);
  launch(SY2D_K_P_UPDATE);
  launch(SY2D_K_SPMV_V);
  launch(SY2D_K_S_UPDATE);
  launch(which);  // warm-up of the kernel under test
  CU(cudaGetLastError());
  CU(cudaEventRecord(c->ev_call0, c->stream));
  for (int r = 0; r < reps; ++r) launch(which);
  CU(cudaEventRecord(c->ev_call1, c->stream));
  CU(cudaEventSynchronize(c->ev_call1));
  CU(cudaGetLastError());
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, c->ev_call0, c->ev_call1));
  *ms_per_launch = (double)ms / reps;
  CU(cudaMemsetAsync(c->scal, 0, c->nbatch * sizeof(Scal), c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return SY2D_OK;
}

int sy2d_measure_peaks(int device, sy2d_peaks* out) {
  sy2d_ctx* c = nullptr;
  if (!out) return SY2D_ERR_INVALID;
  std::memset(out, 0, sizeof *out);
  CU(cudaSetDevice(device));
  int sms = 0, khz = 0;
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  CU(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device));
  out->sm_count = sms;
  out->sm_clock_mhz = khz * 1e-3;
  cudaStream_t st = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  double *a = nullptr, *b = nullptr, *sink = nullptr;
  int rc = SY2D_OK;
  auto body = [&]() -> int {
    CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    const size_t big = (size_t)1 << 27;   // doubles: 1 GB per buffer
    CU(cudaMalloc(reinterpret_cast<void**>(&a), big * sizeof(double)));
    CU(cudaMalloc(reinterpret_cast<void**>(&b), big * sizeof(double)));
    CU(cudaMalloc(reinterpret_cast<void**>(&sink), (size_t)sms * 2 * sizeof(double)));
    CU(cudaMemsetAsync(a, 0, big * sizeof(double), st));
    CU(cudaMemsetAsync(b, 0, big * sizeof(double), st));
    auto timed = [&](auto&& launch, int reps, double* best_ms) -> int {
      *best_ms = 1e30;
      for (int trial = 0; trial < 5; ++trial) {
        CU(cudaEventRecord(e0, st));
        for (int r = 0; r < reps; ++r) launch();
        CU(cudaEventRecord(e1, st));
        CU(cudaEventSynchronize(e1));
        CU(cudaGetLastError());
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        *best_ms = std::min(*best_ms, (double)ms / reps);
      }
      return SY2D_OK;
    };
    double ms = 0.0;
    // shared memory: 2 CTAs of 1024 threads per SM, 96 KB each
    const int smem_bytes = kPeakSmemDoubles * (int)sizeof(double), smem_reps = 400;
    CU(cudaFuncSetAttribute(k_peak_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    int r0 = timed([&] { k_peak_smem<<<sms * 2, 1024, smem_bytes, st>>>(sink, smem_reps); }, 1, &ms);
    if (r0) return r0;
    out->smem_gbs = (double)sms * 2 * smem_reps * (kPeakSmemDoubles / 2) * 8.0 * 4.0 / (ms * 1e-3) / 1e9;
    // L2: 32 MB -> 32 MB, resident after the first pass
    const size_t l2n = (size_t)1 << 22;   // doubles: 32 MB
    r0 = timed([&] { k_peak_copy<<<sms * 8, 256, 0, st>>>(reinterpret_cast<const double2*>(a), reinterpret_cast<double2*>(b), l2n / 2); }, 20, &ms);
    if (r0) return r0;
    out->l2_gbs = 2.0 * l2n * 8.0 / (ms * 1e-3) / 1e9;
    // HBM: 1 GB -> 1 GB
    r0 = timed([&] { k_peak_copy<<<sms * 8, 256, 0, st>>>(reinterpret_cast<const double2*>(a), reinterpret_cast<double2*>(b), big / 2); }, 2, &ms);
    if (r0) return r0;
    out->hbm_gbs = 2.0 * big * 8.0 / (ms * 1e-3) / 1e9;
    return SY2D_OK;
  };
  rc = body();
  if (a) cudaFree(a);
  if (b) cudaFree(b);
  if (sink) cudaFree(sink);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (st) cudaStreamDestroy(st);
  return rc;
}

int sy2d_set_profiling(sy2d_ctx* c, int on) {
  if (!c) return SY2D_ERR_INVALID;
  c->profiling = on != 0;
  c->ev_used = 0;
  std::memset(&c->prof, 0, sizeof c->prof);
  return SY2D_OK;
}

int sy2d_get_profile(sy2d_ctx* c, sy2d_profile* out) {
  if (!c || !out) return SY2D_ERR_INVALID;
  *out = c->prof;
  return SY2D_OK;
}

int sy2d_nccl_unique_id(void* id_out) {
  if (!id_out) return SY2D_ERR_INVALID;
  if (!nccl().ok) { g_create_error = nccl().why; return SY2D_ERR_CUDA; }
  ncclUniqueId id;
  if (nccl().GetUniqueId(&id) != ncclSuccess) { g_create_error = "ncclGetUniqueId failed"; return SY2D_ERR_CUDA; }
  std::memcpy(id_out, &id, sizeof id);
  return SY2D_OK;
}

static int slab_check_shape(int nx_global, int ny, int rank, int nranks) {
  sy2d_ctx* c = nullptr;
  if (nranks < 2 || rank < 0 || rank >= nranks) return fail(c, SY2D_ERR_INVALID, "sy2d_create_slab: need nranks >= 2 and 0 <= rank < nranks");
  if (nx_global / nranks < 2 * kTI) return fail(c, SY2D_ERR_INVALID, "sy2d_create_slab: at least %d rows per rank", 2 * kTI);
  if (ny % 2 || ny < kTJ) return fail(c, SY2D_ERR_INVALID, "sy2d_create_slab: ny must be even and >= %d", kTJ);
  return SY2D_OK;
}

static int slab_finish_create(sy2d_ctx** out, SlabTransport* tp) {
  sy2d_ctx* c = *out;
  c->tp = tp;
  if (cudaMalloc(reinterpret_cast<void**>(&c->d_gather), (size_t)c->nranks * 8 * sizeof(double)) != cudaSuccess) {
    g_create_error = "cudaMalloc failed";
    sy2d_destroy(c);
    *out = nullptr;
    return SY2D_ERR_CUDA;
  }
  c->opt.use_graph = 0;
  return SY2D_OK;
}

int sy2d_create_slab(sy2d_ctx** out, int device, int nx_global, int ny, int rank, int nranks, const void* nccl_id,
                     const double* x_edges, const double* y_edges, double dt) {
  sy2d_ctx* c = nullptr;
  if (!out || !nccl_id) return fail(c, SY2D_ERR_INVALID, "sy2d_create_slab: NULL argument");
  int rc = slab_check_shape(nx_global, ny, rank, nranks);
  if (rc) return rc;
  if (!nccl().ok) return fail(c, SY2D_ERR_CUDA, "sy2d_create_slab: %s", nccl().why.c_str());
  rc = create_impl(out, device, nx_global, ny, 1, x_edges, y_edges, dt, rank, nranks);
  if (rc) return rc;
  c = *out;
  ncclUniqueId id;
  std::memcpy(&id, nccl_id, sizeof id);
  ncclComm_t comm = nullptr;
  ncclResult_t r = nccl().CommInitRank(&comm, nranks, id, rank);
  if (r != ncclSuccess) {
    g_create_error = std::string("ncclCommInitRank failed: ") + nccl().GetErrorString(r);
    sy2d_destroy(c);
    *out = nullptr;
    return SY2D_ERR_CUDA;
  }
  NcclTransport* tp = new NcclTransport;
  tp->comm = comm;
  return slab_finish_create(out, tp);
}

int sy2d_local_group_create(sy2d_local_group** out, int nranks) {
  sy2d_ctx* c = nullptr;
  if (!out || nranks < 2 || nranks > kMgMaxRanks) return fail(c, SY2D_ERR_INVALID, "sy2d_local_group_create: 2 <= nranks <= %d", kMgMaxRanks);
  sy2d_local_group* g = new sy2d_local_group;
  g->nranks = nranks;
  g->ptr.assign(nranks, nullptr);
  g->rows.assign(nranks, 0);
  g->ready.assign(nranks, nullptr);
  g->done.assign(nranks, nullptr);
  *out = g;
  return SY2D_OK;
}

void sy2d_local_group_destroy(sy2d_local_group* g) {
  if (!g) return;
  {
    std::lock_guard<std::mutex> lk(g->mu);
    if (g->attached > 0) return;   // contexts still attached: destroy them first (the group is leaked rather than freed under them)
  }
  delete g;
}

int sy2d_create_slab_local(sy2d_ctx** out, int device, int nx_global, int ny, int rank, int nranks, sy2d_local_group* group,
                           const double* x_edges, const double* y_edges, double dt) {
  sy2d_ctx* c = nullptr;
  if (!out || !group) return fail(c, SY2D_ERR_INVALID, "sy2d_create_slab_local: NULL argument");
  if (group->nranks != nranks) return fail(c, SY2D_ERR_INVALID, "sy2d_create_slab_local: the group was created for %d ranks", group->nranks);
  int rc = slab_check_shape(nx_global, ny, rank, nranks);
  if (rc) return rc;
  {
    std::lock_guard<std::mutex> lk(group->mu);
    if (group->ready[rank]) return fail(c, SY2D_ERR_INVALID, "sy2d_create_slab_local: rank %d is already attached", rank);
  }
  rc = create_impl(out, device, nx_global, ny, 1, x_edges, y_edges, dt, rank, nranks);
  if (rc) return rc;
  c = *out;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (cudaEventCreateWithFlags(&e0, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&e1, cudaEventDisableTiming) != cudaSuccess) {
    g_create_error = "cudaEventCreate failed";
    sy2d_destroy(c);
    *out = nullptr;
    return SY2D_ERR_CUDA;
  }
  LocalTransport* tp = new LocalTransport;
  tp->g = group;
  tp->rank = rank;
  {
    std::lock_guard<std::mutex> lk(group->mu);
    group->ready[rank] = e0;
    group->done[rank] = e1;
    group->attached += 1;
  }
  return slab_finish_create(out, tp);
}

int sy2d_slab_rows(const sy2d_ctx* c, int* i_lo, int* i_hi) {
  if (!c || !i_lo || !i_hi) return SY2D_ERR_INVALID;
  *i_lo = c->slab ? c->i_lo : 0;
  *i_hi = c->slab ? c->i_hi : c->nx;
  return SY2D_OK;
}

}  // extern "C"
