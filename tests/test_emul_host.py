"""CPU check of the DEVICE arithmetic: tests/emul/emul_host.cu runs the same
__host__ __device__ per-cell functions the CUDA kernels run (assemble_row,
scale_row, stencil_apply of sayram2d_b200/csrc/sy2d_kernels.cuh) in serial loops.
Compared with the reference's (M,R) and f from tests/golden.  This harness lives
in tests/ only; the product has no CPU path."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, CASE_OF, bc_for, load_golden, max_rel

SRC = os.path.join(ROOT, "tests", "emul", "emul_host.cu")
SO = os.path.join(ROOT, "tests", "emul", "_emul_host.so")
dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)


def P(a):
    return None if a is None else a.ctypes.data_as(dp)


@pytest.fixture(scope="module")
def emul():
    deps = [SRC] + [os.path.join(ROOT, "sayram2d_b200", "csrc", h) for h in ("sy2d_kernels.cuh", "sy2d_geometry.h")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.run(["/usr/local/cuda/bin/nvcc", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
                        "-Xcompiler", "-fPIC", "-shared", "-o", SO, SRC], check=True)
    lib = C.CDLL(SO)
    lib.emul_create.restype = C.c_void_p
    lib.emul_create.argtypes = [C.c_int, C.c_int, dp, dp, C.c_double, dp, dp, dp, dp, dp, ip, dp, dp, dp, dp]
    lib.emul_assemble.argtypes = [C.c_void_p, dp, dp, dp, dp]
    lib.emul_step.argtypes = [C.c_void_p, dp, dp, C.c_double, C.c_int, C.c_int, dp]
    lib.emul_step.restype = C.c_int
    lib.emul_destroy.argtypes = [C.c_void_p]
    return lib


def make(lib, g, case):
    nx, ny = g["G"].shape
    bct, lines = bc_for(case, g["x_edges"], g["y_edges"])
    bca = np.array(bct, dtype=np.int32)
    return lib.emul_create(nx, ny, P(g["x_edges"]), P(g["y_edges"]), g["meta"]["dt"], P(g["G"]), P(g["Dxx"]), P(g["Dxy"]),
                           P(g["Dyy"]), P(g["inv_tau"]), bca.ctypes.data_as(ip), *[P(l) for l in lines])


@pytest.mark.parametrize("tag", ["ay80", "lc80", "nu48x40", "syn64x48"])
def test_device_assembly_arithmetic_matches_reference(emul, tag):
    import ppfv_oracle as O
    g = load_golden(tag)
    nx, ny = g["G"].shape
    h = make(emul, g, CASE_OF[tag])
    diags = np.empty((5, nx, ny)); R = np.empty((nx, ny)); vf = np.empty((nx + 1, ny + 1))
    f0 = g["f_0"].copy()
    emul.emul_assemble(h, P(f0), P(diags), P(R), P(vf))
    for k, name in enumerate(("diag", "W", "E", "S", "N")):
        ref = g["op1_" + name]
        assert np.max(np.abs(diags[k] - ref)) <= 5e-14 * np.max(np.abs(ref)), name
    assert np.max(np.abs(R - g["op1_R"])) <= 5e-14 * np.max(np.abs(g["op1_R"]))
    # vertex values against the oracle's fill_vertex_from_cells/bcs (Solver.cc:292-422)
    m = O.Mesh(g["x_edges"], g["y_edges"], g["meta"]["dt"])
    bct, lines = bc_for(CASE_OF[tag], g["x_edges"], g["y_edges"])
    eq = O.Equation(m); eq.bc = list(bct); eq.dirichlet_lines = lambda t: lines
    ref_vf = O.fill_vertex_from_bcs(m, eq, O.fill_vertex_from_cells(m, f0), 0.0)
    assert np.max(np.abs(vf - ref_vf)) <= 1e-15 * np.max(np.abs(ref_vf))
    emul.emul_destroy(h)


@pytest.mark.parametrize("tag,nsteps,key,tol", [("ay80", 50, "f_1", 1e-9), ("lc80", 50, "f_1", 1e-9),
                                                ("nu48x40", 20, "f_20", 1e-9), ("syn64x48", 10, "f_10", 1e-9)])
def test_device_iteration_arithmetic_matches_reference(emul, tag, nsteps, key, tol):
    g = load_golden(tag)
    h = make(emul, g, CASE_OF[tag])
    f = g["f_0"].copy(); yprev = np.ones_like(f); res = C.c_double()
    for _ in range(nsteps):
        it = emul.emul_step(h, P(f), P(yprev), 1e-14, 5000, 1, C.byref(res))
        assert it >= 0 and res.value < 1e-13
    assert max_rel(f, g[key]) < tol and (f < 0).sum() == 0
    emul.emul_destroy(h)
