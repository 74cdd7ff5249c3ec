"""ctypes binding of include/sayram2d.h (one Engine = one sy2d_ctx on one GPU)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_LIB = None

K_NAMES = ("assembly", "p_update", "spmv_v", "s_update", "spmv_t", "xr_update", "finish", "other", "problem_steps",
           "mg_line", "mg_resid", "mg_setup")
DIRICHLET, ZEROFLUX = 0, 1


class Sy2dError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"sy2d error {code}: {msg}")
        self.code = code


class Options(C.Structure):
    _fields_ = [("tol", C.c_double), ("maxit", C.c_int), ("precond", C.c_int), ("predictor", C.c_int),
                ("check_every", C.c_int), ("use_graph", C.c_int), ("engine", C.c_int), ("mg_levels", C.c_int),
                ("mg_coarse_sweeps", C.c_int), ("reserved", C.c_int * 3)]


class Stats(C.Structure):
    _fields_ = [("steps", C.c_longlong), ("iters_total", C.c_longlong), ("iters_last", C.c_int),
                ("restarts_total", C.c_int), ("resid_last", C.c_double), ("fmin", C.c_double),
                ("negatives", C.c_longlong), ("seconds_device", C.c_double), ("kernel_launches", C.c_longlong),
                ("iters_sum_all", C.c_longlong), ("engine", C.c_int), ("precond", C.c_int)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Peaks(C.Structure):
    _fields_ = [("smem_gbs", C.c_double), ("l2_gbs", C.c_double), ("hbm_gbs", C.c_double), ("sm_count", C.c_int),
                ("sm_clock_mhz", C.c_double)]


class Profile(C.Structure):
    _fields_ = [("ms", C.c_double * len(K_NAMES)), ("launches", C.c_longlong * len(K_NAMES)),
                ("cells", C.c_double * len(K_NAMES))]


def library_path():
    return os.path.join(_PKG, "lib", "libsayram2d_b200.so")


def load_library():
    """Loads the CUDA library; there is no fallback - a missing .so is an error."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `python -m sayram2d_b200.build` "
                          "(the engine has no CPU or PyTorch fallback)")
    lib = C.CDLL(path)
    dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
    sig = {
        "sy2d_create": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, C.c_double]),
        "sy2d_destroy": (None, [vp]),
        "sy2d_last_error": (C.c_char_p, [vp]),
        "sy2d_default_options": (C.c_int, [C.POINTER(Options)]),
        "sy2d_set_options": (C.c_int, [vp, C.POINTER(Options)]),
        "sy2d_set_coeffs": (C.c_int, [vp, dp, dp, dp, dp, dp]),
        "sy2d_set_coeffs_dev": (C.c_int, [vp, vp, vp, vp, vp, vp]),
        "sy2d_set_bc": (C.c_int, [vp, ip, dp, dp, dp, dp]),
        "sy2d_set_coeffs_async": (C.c_int, [vp, dp, dp, dp, dp, dp]),
        "sy2d_set_bc_async": (C.c_int, [vp, ip, dp, dp, dp, dp]),
        "sy2d_stage_swaps": (C.c_longlong, [vp]),
        "sy2d_steps_begun": (C.c_longlong, [vp]),
        "sy2d_set_f": (C.c_int, [vp, dp]),
        "sy2d_set_f_dev": (C.c_int, [vp, vp]),
        "sy2d_put_f": (C.c_int, [vp, dp]),
        "sy2d_get_f": (C.c_int, [vp, dp]),
        "sy2d_get_f_dev": (C.c_int, [vp, vp]),
        "sy2d_step": (C.c_int, [vp, C.c_int, C.POINTER(Stats)]),
        "sy2d_step_host": (C.c_int, [vp, dp, dp, C.c_int, C.POINTER(Stats)]),
        "sy2d_time": (C.c_double, [vp]),
        "sy2d_step_count": (C.c_longlong, [vp]),
        "sy2d_dump_operator": (C.c_int, [vp, dp, dp]),
        "sy2d_dump_vertex_f": (C.c_int, [vp, dp]),
        "sy2d_dump_scaled_operator": (C.c_int, [vp, dp, dp, dp]),
        "sy2d_debug_vcycle": (C.c_int, [vp, dp, dp, dp, dp]),
        "sy2d_set_profiling": (C.c_int, [vp, C.c_int]),
        "sy2d_get_profile": (C.c_int, [vp, C.POINTER(Profile)]),
        "sy2d_nccl_unique_id": (C.c_int, [C.c_char_p]),
        "sy2d_create_slab": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, dp, dp, C.c_double]),
        "sy2d_slab_rows": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "sy2d_local_group_create": (C.c_int, [C.POINTER(vp), C.c_int]),
        "sy2d_local_group_destroy": (None, [vp]),
        "sy2d_create_slab_local": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, dp, dp, C.c_double]),
        "sy2d_bench_kernel": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(C.c_double)]),
        "sy2d_measure_peaks": (C.c_int, [C.c_int, C.POINTER(Peaks)]),
        "sy2d_build_info": (C.c_char_p, []),
        "sy2d_last_assembly_kernel": (C.c_int, [C.c_void_p]),
        "sy2d_device_count": (C.c_int, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _LIB = lib
    return lib


def nccl_unique_id():
    """128-byte NCCL id created by this process (rank 0); distribute it to the other ranks."""
    lib = load_library()
    buf = C.create_string_buffer(128)
    rc = lib.sy2d_nccl_unique_id(buf)
    if rc != 0:
        raise Sy2dError(rc, lib.sy2d_last_error(None).decode())
    return buf.raw


class LocalGroup:
    """In-process slab transport (sy2d_local_group): nranks slab Engines in ONE process, one host thread each."""

    def __init__(self, nranks):
        self.lib = load_library()
        self.nranks = int(nranks)
        self._g = C.c_void_p()
        rc = self.lib.sy2d_local_group_create(C.byref(self._g), self.nranks)
        if rc != 0:
            raise Sy2dError(rc, self.lib.sy2d_last_error(None).decode())

    def close(self):
        if getattr(self, "_g", None):
            self.lib.sy2d_local_group_destroy(self._g)
            self._g = None


def run_local_slabs(nranks, fn, device=0):
    """Runs fn(rank, group) on nranks host threads (one slab context each; ctypes releases the GIL inside the
    library, so the ranks' collective calls meet).  Returns the list of results; re-raises the first exception."""
    import threading
    group = LocalGroup(nranks)
    out, err = [None] * nranks, [None] * nranks

    def work(r):
        try:
            out[r] = fn(r, group)
        except BaseException as ex:  # noqa: BLE001
            err[r] = ex

    ts = [threading.Thread(target=work, args=(r,)) for r in range(nranks)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    group.close()
    for e in err:
        if e is not None:
            raise e
    return out


def measure_peaks(device=0):
    """Measured shared-memory / L2 / HBM copy bandwidths (GB/s) of the device: roofline denominators."""
    lib = load_library()
    p = Peaks()
    rc = lib.sy2d_measure_peaks(int(device), C.byref(p))
    if rc != 0:
        raise Sy2dError(rc, lib.sy2d_last_error(None).decode())
    return {k: getattr(p, k) for k, _ in p._fields_}


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a, shape=None, name="array"):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and tuple(a.shape) != tuple(shape):
        if a.size == int(np.prod(shape)):
            a = a.reshape(shape)
        else:
            raise ValueError(f"{name}: expected shape {tuple(shape)}, got {a.shape}")
    return a


class Engine:
    """Device-resident time stepper: Solver::update() of the reference (Solver.cc:270-290)
    for nbatch independent problems sharing one mesh and one set of BCs."""

    def __init__(self, x_edges, y_edges, dt, nbatch=1, device=0, slab=None):
        """slab = (rank, nranks, nccl_id_bytes | LocalGroup): this engine holds one row slab of the grid
        (sy2d_create_slab / sy2d_create_slab_local); field arguments are then the owned rows [i_lo:i_hi]."""
        self.lib = load_library()
        xe = _f64(x_edges)
        ye = _f64(y_edges)
        self.nx, self.ny, self.nbatch, self.dt = xe.size - 1, ye.size - 1, int(nbatch), float(dt)
        self._ctx = C.c_void_p()
        if slab is None:
            rc = self.lib.sy2d_create(C.byref(self._ctx), int(device), self.nx, self.ny, self.nbatch, _dp(xe), _dp(ye), self.dt)
        else:
            rank, nranks, nccl_id = slab
            if isinstance(nccl_id, LocalGroup):
                rc = self.lib.sy2d_create_slab_local(C.byref(self._ctx), int(device), self.nx, self.ny, int(rank), int(nranks),
                                                     nccl_id._g, _dp(xe), _dp(ye), self.dt)
            else:
                rc = self.lib.sy2d_create_slab(C.byref(self._ctx), int(device), self.nx, self.ny, int(rank), int(nranks),
                                               bytes(nccl_id), _dp(xe), _dp(ye), self.dt)
        if rc != 0:
            msg = self.lib.sy2d_last_error(None).decode()
            self._ctx = None
            raise Sy2dError(rc, msg)
        self.i_lo, self.i_hi = 0, self.nx
        if slab is not None:
            lo, hi = C.c_int(), C.c_int()
            self.lib.sy2d_slab_rows(self._ctx, C.byref(lo), C.byref(hi))
            self.i_lo, self.i_hi = lo.value, hi.value
        self.shape = (self.nbatch, self.i_hi - self.i_lo, self.ny)

    def _check(self, rc):
        if rc != 0:
            raise Sy2dError(rc, self.lib.sy2d_last_error(self._ctx).decode())

    def close(self):
        if getattr(self, "_ctx", None):
            self.lib.sy2d_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- configuration ----------------------------------------------------
    def options(self):
        o = Options()
        self.lib.sy2d_default_options(C.byref(o))
        return o

    def set_options(self, **kw):
        o = getattr(self, "_opt", None) or self.options()
        for k, v in kw.items():
            setattr(o, k, v)
        self._check(self.lib.sy2d_set_options(self._ctx, C.byref(o)))
        self._opt = o
        return o

    def set_coeffs(self, G, Dxx, Dxy, Dyy, inv_tau=None):
        arrs = [_f64(a, self.shape, n) for a, n in ((G, "G"), (Dxx, "Dxx"), (Dxy, "Dxy"), (Dyy, "Dyy"), (inv_tau, "inv_tau"))]
        self._check(self.lib.sy2d_set_coeffs(self._ctx, *[_dp(a) for a in arrs]))

    def set_coeffs_dev(self, G, Dxx, Dxy, Dyy, inv_tau=None):
        """Device pointers (ints, e.g. torch.Tensor.data_ptr())."""
        self._check(self.lib.sy2d_set_coeffs_dev(self._ctx, G, Dxx, Dxy, Dyy, inv_tau))

    def set_bc(self, bc_type, xmin=None, xmax=None, ymin=None, ymax=None):
        t = (C.c_int * 4)(*[int(b) for b in bc_type])
        lines = [_f64(xmin, (self.ny + 1,), "xmin"), _f64(xmax, (self.ny + 1,), "xmax"),
                 _f64(ymin, (self.nx + 1,), "ymin"), _f64(ymax, (self.nx + 1,), "ymax")]
        self._check(self.lib.sy2d_set_bc(self._ctx, t, *[_dp(a) for a in lines]))

    def set_coeffs_async(self, G, Dxx, Dxy, Dyy, inv_tau=None):
        """Stage the NEXT step's fields without waiting for the step in flight (sy2d_set_coeffs_async)."""
        arrs = [_f64(a, self.shape, n) for a, n in ((G, "G"), (Dxx, "Dxx"), (Dxy, "Dxy"), (Dyy, "Dyy"), (inv_tau, "inv_tau"))]
        self._check(self.lib.sy2d_set_coeffs_async(self._ctx, *[_dp(a) for a in arrs]))

    def set_bc_async(self, bc_type, xmin=None, xmax=None, ymin=None, ymax=None):
        t = (C.c_int * 4)(*[int(b) for b in bc_type])
        lines = [_f64(xmin, (self.ny + 1,), "xmin"), _f64(xmax, (self.ny + 1,), "xmax"),
                 _f64(ymin, (self.nx + 1,), "ymin"), _f64(ymax, (self.nx + 1,), "ymax")]
        self._check(self.lib.sy2d_set_bc_async(self._ctx, t, *[_dp(a) for a in lines]))

    def stage_swaps(self):
        return self.lib.sy2d_stage_swaps(self._ctx)

    def steps_begun(self):
        return self.lib.sy2d_steps_begun(self._ctx)

    def step_overlapped(self, stage_next, nsteps=1):
        """Runs step(nsteps) on a helper thread and calls stage_next() - which should use set_coeffs_async /
        set_bc_async - on this thread once the step has taken its own fields (sy2d_steps_begun advanced)."""
        import threading
        import time as _t
        res = {}

        def work():
            try:
                res["st"] = self.step(nsteps)
            except BaseException as ex:  # noqa: BLE001
                res["err"] = ex

        begun = self.steps_begun()
        th = threading.Thread(target=work)
        th.start()
        while self.steps_begun() == begun and th.is_alive():
            _t.sleep(0)
        try:
            stage_next()
        finally:
            th.join()
        if "err" in res:
            raise res["err"]
        return res["st"]

    def set_f(self, f):
        self._check(self.lib.sy2d_set_f(self._ctx, _dp(_f64(f, self.shape, "f"))))

    def put_f(self, f):
        """Upload f without resetting the step counter / predictor (host-resident f)."""
        self._check(self.lib.sy2d_put_f(self._ctx, _dp(_f64(f, self.shape, "f"))))

    def set_f_dev(self, ptr):
        self._check(self.lib.sy2d_set_f_dev(self._ctx, ptr))

    # -- stepping ---------------------------------------------------------
    def step(self, nsteps=1):
        st = Stats()
        rc = self.lib.sy2d_step(self._ctx, int(nsteps), C.byref(st))
        self.last_stats = st.as_dict()
        self._check(rc)
        return self.last_stats

    def step_host(self, f_in, f_out, nsteps=1):
        """Host-resident f: upload f_in, take nsteps, download into f_out (pipelined over
        sub-batches when the arrays are pinned).  Arrays must be C-contiguous float64 of self.shape."""
        st = Stats()
        for a in (f_in, f_out):
            if a is not None and (a.dtype != np.float64 or not a.flags["C_CONTIGUOUS"] or a.size != int(np.prod(self.shape))):
                raise ValueError("step_host: arrays must be C-contiguous float64 of the engine's shape")
        rc = self.lib.sy2d_step_host(self._ctx, _dp(f_in), _dp(f_out), int(nsteps), C.byref(st))
        self.last_stats = st.as_dict()
        self._check(rc)
        return self.last_stats

    def get_f(self, out=None):
        if out is None:
            out = np.empty(self.shape, dtype=np.float64)
        self._check(self.lib.sy2d_get_f(self._ctx, _dp(out)))
        return out

    def get_f_dev(self, ptr):
        self._check(self.lib.sy2d_get_f_dev(self._ctx, ptr))

    def time(self):
        return self.lib.sy2d_time(self._ctx)

    def step_count(self):
        return self.lib.sy2d_step_count(self._ctx)

    # -- parity helpers ---------------------------------------------------
    def dump_operator(self):
        diags = np.empty((5,) + self.shape)
        rhs = np.empty(self.shape)
        self._check(self.lib.sy2d_dump_operator(self._ctx, _dp(diags), _dp(rhs)))
        return dict(diag=diags[0], W=diags[1], E=diags[2], S=diags[3], N=diags[4], R=rhs)

    def dump_scaled_operator(self):
        """Scaled system of the current f by the configured assembly kernel -> (w4, rhs, cs)."""
        w4 = np.empty((4,) + self.shape)
        rhs = np.empty(self.shape)
        cs = np.empty(self.shape)
        self._check(self.lib.sy2d_dump_scaled_operator(self._ctx, _dp(w4), _dp(rhs), _dp(cs)))
        return w4, rhs, cs

    def dump_vertex_f(self):
        vf = np.empty((self.nbatch, self.nx + 1, self.ny + 1))
        self._check(self.lib.sy2d_dump_vertex_f(self._ctx, _dp(vf)))
        return vf

    def debug_vcycle(self, r):
        """One multigrid V-cycle applied to r on the operator of the current f -> (z, w4, om)."""
        r = _f64(r, self.shape, "r")
        z = np.empty(self.shape)
        w4 = np.empty((4,) + self.shape)
        om = np.empty(self.shape)
        self._check(self.lib.sy2d_debug_vcycle(self._ctx, _dp(r), _dp(z), _dp(w4), _dp(om)))
        return z, w4, om

    # -- profiling --------------------------------------------------------
    def set_profiling(self, on=True):
        self._check(self.lib.sy2d_set_profiling(self._ctx, int(bool(on))))

    def profile(self):
        p = Profile()
        self._check(self.lib.sy2d_get_profile(self._ctx, C.byref(p)))
        return {n: {"ms": p.ms[k], "launches": p.launches[k], "cells": p.cells[k]} for k, n in enumerate(K_NAMES)}

    def bench_kernel(self, name, reps=20):
        """Sustained ms per launch of one lockstep-engine kernel (back-to-back launches)."""
        ms = C.c_double()
        self._check(self.lib.sy2d_bench_kernel(self._ctx, K_NAMES.index(name), int(reps), C.byref(ms)))
        return ms.value

    def last_assembly_kernel(self):
        """1 per cell, 2 plain-load tiles, 3 marching warps, 4 TMA tiles, 5 TMA column runs, 6 TMA tiles with two cells per thread."""
        return int(self.lib.sy2d_last_assembly_kernel(self._ctx))

    def build_info(self):
        return self.lib.sy2d_build_info().decode()
