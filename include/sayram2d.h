/*
 * sayram2d.h - C ABI of the B200-native Sayram-2D time-step engine.
 *
 * The reference (xtaohub/Sayram-2D) has no FFI: its seam is the C++ class
 * `Solver` (source/Solver.h:18-25) as used by source/main.cc:51,74,80,83.  This
 * header is what a host-side `Solver` forwards to; sayram2d_b200/host/Solver.h
 * is that forwarding class and keeps the reference's public API unchanged.
 *
 * Conventions
 *  - plain pointers and sizes only; every array argument is CALLER-OWNED HOST
 *    memory (the *_dev variants take device pointers) and is copied during the
 *    call; nothing is retained.
 *  - 2-D fields are (nx, ny) row-major with j (log E) fastest - the reference's
 *    xtensor layout (source/common.h:28) - with an optional leading batch
 *    dimension [nbatch][nx][ny] for ensembles of independent problems that
 *    share one mesh and one set of boundary conditions.
 *  - every function returns SY2D_OK (0) or a negative sy2d_status; the message
 *    of the last failure is available from sy2d_last_error().  No exception
 *    crosses this boundary.  A context is bound to one device and must be used
 *    by one host thread at a time (the reference is single-threaded too).
 *  - there is NO CPU fallback: sy2d_create fails when no CUDA device is usable.
 */
#ifndef SAYRAM2D_H_
#define SAYRAM2D_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sy2d_ctx sy2d_ctx;

typedef enum {
  SY2D_OK = 0,
  SY2D_ERR_INVALID = -1,       /* bad argument (sizes, NULL, non-increasing edges: Grid2D.h:44-67) */
  SY2D_ERR_CUDA = -2,          /* CUDA runtime failure or no device */
  SY2D_ERR_NOT_CONVERGED = -3, /* Krylov solve stopped without converging (maxit, breakdown, NaN) or its true residual
                                  exceeds 1000 tol (the reference's direct LU cannot fail this way).  The failing time step is
                                  NOT committed: f, the predictor state and the step counter stay those of the last completed
                                  step, so the call may be retried with other options.  Engine 2 (batches): a failed member keeps
                                  its last completed step while the others finish the call; stats.steps and the step counter
                                  advance by what EVERY member completed. */
  SY2D_ERR_STATE = -4,         /* call order: coefficients / BCs / f not set yet */
  SY2D_ERR_BC = -5             /* Dirichlet side without data: "Dirichlet BC: missing value." (Solver.cc:393) */
} sy2d_status;

/* source/BCTypes.h:13,16 */
enum { SY2D_XMIN = 0, SY2D_XMAX = 1, SY2D_YMIN = 2, SY2D_YMAX = 3 };
enum { SY2D_DIRICHLET = 0, SY2D_ZEROFLUX = 1 };

/* preconditioner inside the BiCGSTAB loop */
enum {
  SY2D_PRECOND_AUTO = -1,  /* engine 2: XLINE where the grid fits (nx <= 80, ny <= 128), else JACOBI;
                              engine 1: MG where the grid fits (8 <= rows per context <= 8192, ny a multiple of 4 and >= 16), else XLINE */
  SY2D_PRECOND_JACOBI = 0, /* the unit-diagonal scaling itself */
  SY2D_PRECOND_XLINE = 1,  /* right preconditioning by the tridiagonal along i (alpha0): ~4x fewer iterations
                              (engine 1: block-Jacobi with 16-row line segments) */
  SY2D_PRECOND_MG = 2      /* engine 1: one V(1,1) multigrid cycle, semi-coarsening along j (log E), damped
                              whole-x-line Jacobi smoother on every level: O(15) iterations independent of the grid */
};

typedef struct {
  double tol;        /* stop when max_K |r_K| <= tol on the scaled system (unknown ~ O(1)); default 1e-14 */
  int maxit;         /* per linear solve; default 20000 */
  int precond;       /* SY2D_PRECOND_*; default AUTO */
  int predictor;     /* column scale (initial guess) of a step: 0 f^n; 1 f^n y_n with y_n = f^n / f^{n-1} the last per-cell ratio;
                        2 f^n y_n (y_n / y_{n-1}), its geometric extrapolation (one more state array).  Ratios are clamped
                        to [0.5, 2]; the solution does not depend on the choice beyond the tolerance; default 2 */
  int check_every;   /* iterations between host convergence polls; default 16 */
  int use_graph;     /* 1: replay the iteration chunk as a CUDA graph; default 1 */
  int engine;        /* 0 auto (2 when nx*ny <= 16384, else 1), 1 lockstep multi-kernel,
                        2 one persistent CTA per problem (whole time loop in one launch) */
  int mg_levels;     /* SY2D_PRECOND_MG: maximum number of grid levels (0 = default: coarsen to <= 64 columns) */
  int mg_coarse_sweeps; /* SY2D_PRECOND_MG: smoothing sweeps on the coarsest level (0 = default 2) */
  int reserved[3];   /* reserved[0]: assembly kernel of engine 1 - 0 default, 1 one thread per cell, 2 tile kernel with plain loads, 3 warp-marching kernel, 4 TMA-staged tiles (tests); reserved[1]: slab mode,
                        fixed iteration budget without a convergence error (bench); reserved[2]: 1 makes the first attempt of every
                        AUTO multigrid step count as failed, so the x-line fallback runs (tests); 2 (slab mode, multigrid): the smoother's
                        lines end at the slab instead of being coupled across ranks by the spike correction (comparison runs) */
} sy2d_options;

typedef struct {
  long long steps;            /* time steps taken by the call */
  long long iters_total;      /* BiCGSTAB iterations summed over steps (max over the batch per step) */
  int iters_last;             /* iterations of the last step (max over the batch) */
  int restarts_total;         /* time steps redone with the x-line iteration after the multigrid-preconditioned solve
                                 of an AUTO context stopped without converging */
  double resid_last;          /* max over batch and cells of |rhs - A x| after the last solve (true residual) */
  double fmin;                /* min f over batch and cells after the last step */
  long long negatives;        /* number of cells with f < 0 after the last step */
  double seconds_device;      /* CUDA-event time of the whole call on the context's stream */
  long long kernel_launches;  /* kernels of this library launched by the call (graph nodes included) */
  long long iters_sum_all;    /* iterations summed over steps AND batch members (mean = / (steps*nbatch));
                                 lockstep engine: iters_total * nbatch (all members iterate together) */
  int engine;                 /* engine that ran: 1 lockstep multi-kernel, 2 one CTA per problem */
  int precond;                /* preconditioner that ran (SY2D_PRECOND_JACOBI, _XLINE or _MG) */
} sy2d_stats;

/* per-kernel device time of profiled sy2d_step calls (see sy2d_set_profiling) */
enum {
  SY2D_K_ASSEMBLY = 0, /* fused PPFV assembly                          */
  SY2D_K_P_UPDATE,     /* p = r + beta (p - omega v)                   */
  SY2D_K_SPMV_V,       /* v = A p, (rhat, v)                           */
  SY2D_K_S_UPDATE,     /* s = r - alpha v                              */
  SY2D_K_SPMV_T,       /* t = A s, (t,s), (t,t)                        */
  SY2D_K_XR_UPDATE,    /* x, r update, (rhat, r), max|r|               */
  SY2D_K_FINISH,       /* f = c (1 + d), statistics                    */
  SY2D_K_OTHER,        /* true-residual check                          */
  SY2D_K_PROBLEM_STEPS,/* engine 2: whole time steps, one CTA per problem */
  SY2D_K_MG_LINE,      /* multigrid: whole-column Thomas solves (smoother), all levels */
  SY2D_K_MG_RESID,     /* multigrid: residual + restriction / prolongation + residual, all levels */
  SY2D_K_MG_SETUP,     /* multigrid: coarse operators and line factorisations, once per time step */
  SY2D_K_COUNT
};
typedef struct {
  double ms[SY2D_K_COUNT];          /* summed CUDA-event time per kernel */
  long long launches[SY2D_K_COUNT];
  double cells[SY2D_K_COUNT];       /* cells of still-active problems summed over those launches
                                       (exact when check_every == 1, which profiling forces) */
} sy2d_profile;

/* Replaces Mesh(grid, dt) + Solver ctor allocation (Mesh.cc:17-65, Solver.cc:14-34). */
int sy2d_create(sy2d_ctx** out, int device, int nx, int ny, int nbatch,
                const double* x_edges /* nx+1 */, const double* y_edges /* ny+1 */, double dt);
void sy2d_destroy(sy2d_ctx* ctx);
const char* sy2d_last_error(const sy2d_ctx* ctx); /* ctx may be NULL: error of a failed sy2d_create */

int sy2d_default_options(sy2d_options* opt);
int sy2d_set_options(sy2d_ctx* ctx, const sy2d_options* opt);

/* Replaces Equation::G/Dxx/Dxy/Dyy/inv_tau + Solver::update_Lambda (Equation.h:44-50,
 * Solver.cc:57-65).  Each array is [nbatch][nx][ny]; inv_tau may be NULL (= 0, Equation.h:39-40).
 * Call again whenever Equation::update(t) changed a field (Solver.cc:287-288). */
int sy2d_set_coeffs(sy2d_ctx* ctx, const double* G, const double* Dxx, const double* Dxy,
                    const double* Dyy, const double* inv_tau);
/* Same with device pointers (no host round trip). */
int sy2d_set_coeffs_dev(sy2d_ctx* ctx, const double* G, const double* Dxx, const double* Dxy,
                        const double* Dyy, const double* inv_tau);

/* Replaces Equation::bc_type + dirichlet_vertex_value + Solver::fill_vertex_from_bcs
 * (Equation.h:52,60-65, Solver.cc:385-422).  bc_type[side] in SY2D_XMIN..YMAX order; the
 * vertex lines xmin/xmax have ny+1 entries, ymin/ymax nx+1; a line may be NULL only for a
 * ZeroFlux side (else SY2D_ERR_BC).  Shared by all batch members.  Call again when the
 * Dirichlet data change with t. */
int sy2d_set_bc(sy2d_ctx* ctx, const int bc_type[4], const double* xmin, const double* xmax,
                const double* ymin, const double* ymax);

/* Asynchronous flavours for time-dependent cases (Equation::update(t) followed by update_Lambda / update_vertex_f,
 * Solver.cc:286-289): the arrays are copied into pinned staging owned by the context (they are free again on return),
 * uploaded and folded on a separate copy stream into a SECOND set of device buffers while a time step may be running,
 * and become the active set at the start of the NEXT sy2d_step / sy2d_step_host call, whose kernels wait for the upload
 * on the device - the host never does.  These two calls (and only these) may be made from another host thread while
 * sy2d_step is executing on the same context.  A later call before the next step overwrites an earlier one; a
 * synchronous sy2d_set_coeffs / sy2d_set_bc supersedes a pending asynchronous one.  Not available on slab contexts. */
int sy2d_set_coeffs_async(sy2d_ctx* ctx, const double* G, const double* Dxx, const double* Dxy,
                          const double* Dyy, const double* inv_tau);
int sy2d_set_bc_async(sy2d_ctx* ctx, const int bc_type[4], const double* xmin, const double* xmax,
                      const double* ymin, const double* ymax);
long long sy2d_stage_swaps(const sy2d_ctx* ctx);   /* asynchronously staged sets swapped in so far */
/* Number of sy2d_step / sy2d_step_host calls that have passed their swap-in point.  A thread that stages the fields of
 * step n+1 while another thread runs step n must not stage before step n has taken ITS fields: read the counter before
 * starting the step thread and wait until it has advanced (dropin/Solver.cc does exactly this). */
long long sy2d_steps_begun(sy2d_ctx* ctx);

/* Solver::init f_ = eq.init_f (Solver.cc:38-42); [nbatch][nx][ny]. Resets the step counter.
 * f must be finite and > 0 in every cell (SY2D_ERR_INVALID otherwise): the engine solves for the per-cell ratio
 * f^{n+1}/f^n.  The reference's cases add gEPS to f0 (Albert_Young.h:39) and the PPFV scheme keeps f positive. */
int sy2d_set_f(sy2d_ctx* ctx, const double* f);
int sy2d_set_f_dev(sy2d_ctx* ctx, const double* f_dev);
/* Overwrites f from the host WITHOUT resetting the step counter or the predictor state:
 * the host-resident-f flavour of Solver::update() (f_ lives on the host in the reference). */
int sy2d_put_f(sy2d_ctx* ctx, const double* f);
/* Solver::f() (Solver.h:24). */
int sy2d_get_f(sy2d_ctx* ctx, double* f_out);
int sy2d_get_f_dev(sy2d_ctx* ctx, double* f_out_dev);

/* Solver::update() x nsteps (Solver.cc:270-290) without leaving the device; stats may be NULL. */
int sy2d_step(sy2d_ctx* ctx, int nsteps, sy2d_stats* stats);
/* Solver::update() x nsteps with f resident on the HOST, as in the reference (f_ is a host array there):
 * f_in (may be NULL = keep the device state) is uploaded, nsteps are taken, the result is written to
 * f_out (may be NULL).  For batches the copies and the kernels of contiguous sub-batches are pipelined
 * over several streams; pass pinned memory to get the overlap.  Does not reset the step counter. */
int sy2d_step_host(sy2d_ctx* ctx, const double* f_in, double* f_out, int nsteps, sy2d_stats* stats);
/* Solver::t() (Solver.h:23). */
double sy2d_time(const sy2d_ctx* ctx);
long long sy2d_step_count(const sy2d_ctx* ctx);

/* Assembles M(f), R(f) for the CURRENT f exactly as Solver::assemble does (Solver.cc:167-202)
 * and returns them unscaled: diags = [5][nbatch][nx][ny] in the order diag, W(i-1), E(i+1),
 * S(j-1), N(j+1); rhs = [nbatch][nx][ny].  For parity tests; does not advance time. */
int sy2d_dump_operator(sy2d_ctx* ctx, double* diags, double* rhs);
/* vertex_f_ (Solver.cc:292-422) for the current f: [nbatch][nx+1][ny+1]. For parity tests. */
int sy2d_dump_vertex_f(sy2d_ctx* ctx, double* vf);

/* Test hook: the scaled unit-diagonal system A d = rhs of the current f as the lockstep engine
 * assembles it, by the assembly kernel options.reserved[0] selects: w4 = [4][nbatch][nx][ny] = wW, wE, wS, wN;
 * rhs, cs = [nbatch][nx][ny] (cs may be NULL).  The kernels must agree bit for bit.  Does not advance time. */
int sy2d_dump_scaled_operator(sy2d_ctx* ctx, double* w4, double* rhs, double* cs);

/* Test hook for the multigrid preconditioner (SY2D_PRECOND_MG): assembles the scaled operator of the
 * current f, builds the hierarchy and applies ONE V-cycle to r -> z ([nbatch][nx][ny], host).  w4 (may be
 * NULL) receives the scaled unit-diagonal operator [4][nbatch][nx][ny] = wW, wE, wS, wN; om (may be NULL)
 * the row weights M_KK c_K.  Does not advance time. */
int sy2d_debug_vcycle(sy2d_ctx* ctx, const double* r, double* z, double* w4, double* om);

/* Per-kernel CUDA-event timing of subsequent sy2d_step calls (adds event overhead; off by default). */
int sy2d_set_profiling(sy2d_ctx* ctx, int on);
int sy2d_get_profile(sy2d_ctx* ctx, sy2d_profile* out);

/* Sustained time of ONE kernel of the lockstep engine: the operator is assembled from the current f,
 * the Krylov vectors are primed, then kernel `which` (SY2D_K_ASSEMBLY, _P_UPDATE, _SPMV_V, _S_UPDATE,
 * _SPMV_T, _XR_UPDATE) is launched `reps` times back to back between two CUDA events on the context's
 * stream; *ms_per_launch is the average.  Solver state (f, step count) is left untouched. */
int sy2d_bench_kernel(sy2d_ctx* ctx, int which, int reps, double* ms_per_launch);

/* ---- one large grid split into row slabs over the GPUs of a node (BASELINE config 5) ----
 * Rank r holds rows [i_lo, i_hi) of the nx_global x ny grid (contiguous split along i, the slow axis,
 * so a halo line is ny contiguous doubles).  Per BiCGSTAB iteration: two one-line halo exchanges
 * (ncclSend/Recv with both neighbours) and three all-gathers of 5 doubles that every rank reduces in
 * rank order.  NCCL is loaded with dlopen; rank 0 calls sy2d_nccl_unique_id and the host distributes
 * the 128 bytes (bench.py / the tests use torch.distributed for that).  On a slab context the field
 * arguments of sy2d_set_coeffs / set_f / put_f / get_f are the OWNED rows [i_hi - i_lo][ny]; the
 * boundary lines of sy2d_set_bc stay global.  stats.fmin / negatives cover the owned rows only. */
int sy2d_nccl_unique_id(void* id_out /* 128 bytes */);
int sy2d_create_slab(sy2d_ctx** out, int device, int nx_global, int ny, int rank, int nranks, const void* nccl_id,
                     const double* x_edges /* nx_global+1 */, const double* y_edges /* ny+1 */, double dt);
int sy2d_slab_rows(const sy2d_ctx* ctx, int* i_lo, int* i_hi);

/* The same decomposition with an IN-PROCESS transport instead of NCCL: the nranks slab contexts live in one process
 * (one host thread per context; all on one device, or on several), halo lines and the small gathered vectors move with
 * cudaMemcpyAsync between the contexts' buffers, ordered by CUDA events and a host barrier.  Kernels, spike-coupled
 * multigrid and reduction order are exactly those of the NCCL contexts, so a single-GPU box can run and test the slab
 * path at 2..16 ranks.  Every collective call (sy2d_set_coeffs, sy2d_step ...) must be made by all ranks concurrently,
 * each from its own thread.  Destroy the contexts before the group. */
typedef struct sy2d_local_group sy2d_local_group;
int sy2d_local_group_create(sy2d_local_group** out, int nranks);
void sy2d_local_group_destroy(sy2d_local_group* group);
int sy2d_create_slab_local(sy2d_ctx** out, int device, int nx_global, int ny, int rank, int nranks, sy2d_local_group* group,
                           const double* x_edges /* nx_global+1 */, const double* y_edges /* ny+1 */, double dt);

/* Measured bandwidths of the units the kernels are bound by (copy micro-kernels with 8 / 16-byte accesses, best of 5):
 * the denominators of bench.py's rooflines.  smem_gbs: shared memory / L1TEX data pipe, loads + stores, all SMs;
 * l2_gbs: copy inside the L2 (32 MB -> 32 MB), reads + writes; hbm_gbs: copy of 1 GB -> 1 GB. */
typedef struct {
  double smem_gbs, l2_gbs, hbm_gbs;
  int sm_count;
  double sm_clock_mhz;   /* the device's nominal maximum SM clock */
} sy2d_peaks;
int sy2d_measure_peaks(int device, sy2d_peaks* out);

/* Build/device facts: "sm_100a;cuda=12.9;..." */
const char* sy2d_build_info(void);

/* Which assembly kernel the last engine-1 assembly of this context (sy2d_step, sy2d_dump_scaled_operator) ran with:
 * 1 one thread per cell, 2 shared-memory tiles with plain loads, 3 warp-marching strips, 4 TMA-staged tiles (the default
 * where a tensor map exists: ny even, >= 32), 5 TMA column runs, 6 TMA tiles with two cells per thread; 0 = none yet.
 * A forced kernel (options.reserved[0]) that the grid cannot use falls back to the next one in that list - this call says
 * which.  No counterpart in the reference (diagnostic). */
int sy2d_last_assembly_kernel(const sy2d_ctx* ctx);
int sy2d_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SAYRAM2D_H_ */
