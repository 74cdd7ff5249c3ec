"""Row-slab decomposition (BASELINE config 5) on ONE GPU: the in-process transport
(sy2d_create_slab_local) runs P = 2, 4, 8 slab contexts in one process, so the halo logic, the
all-gathered reductions and the spike-coupled multigrid line solves - the code the NCCL ranks run -
are checked on a single-GPU box against the single-context engine and against the oracle.
Semantics to hold: Solver::update (source/Solver.cc:270-290) of the WHOLE grid."""
import numpy as np
import pytest

from conftest import max_rel

import sayram2d_b200 as sy
from sayram2d_b200 import fields
from sayram2d_b200.engine import run_local_slabs

pytestmark = pytest.mark.gpu

DT = 0.002


def _case(nx, ny):
    xe, ye = fields.uniform_edges(nx, ny)
    Dxx, Dxy, Dyy, inv_tau = fields.synthetic_tensor(xe, ye)
    f0, bct, lines = fields.ay_init_and_bc(xe, ye)
    return dict(xe=xe, ye=ye, G=fields.ay_G(xe, ye), Dxx=Dxx, Dxy=Dxy, Dyy=Dyy, inv_tau=inv_tau, f0=f0, bct=bct, lines=lines)


def _run_slabs(P, cs, nsteps, **opts):
    def work(rank, group):
        eng = sy.Engine(cs["xe"], cs["ye"], DT, slab=(rank, P, group))
        eng.set_options(**opts)
        lo, hi = eng.i_lo, eng.i_hi
        eng.set_coeffs(cs["G"][lo:hi], cs["Dxx"][lo:hi], cs["Dxy"][lo:hi], cs["Dyy"][lo:hi], cs["inv_tau"][lo:hi])
        eng.set_bc(cs["bct"], *cs["lines"])
        eng.set_f(cs["f0"][lo:hi])
        st = eng.step(nsteps)
        f = eng.get_f()[0]
        eng.close()
        return lo, hi, f, st
    res = run_local_slabs(P, work)
    f = np.empty_like(cs["f0"])
    rows = 0
    for lo, hi, fr, st in res:
        f[lo:hi] = fr
        rows += hi - lo
    assert rows == f.shape[0]
    return f, [r[3] for r in res]


def _run_single(cs, nsteps, **opts):
    ref = sy.Engine(cs["xe"], cs["ye"], DT)
    ref.set_options(engine=1, **opts)
    ref.set_coeffs(cs["G"], cs["Dxx"], cs["Dxy"], cs["Dyy"], cs["inv_tau"])
    ref.set_bc(cs["bct"], *cs["lines"])
    ref.set_f(cs["f0"])
    st = ref.step(nsteps)
    f = ref.get_f()[0]
    ref.close()
    return f, st


@pytest.mark.parametrize("P", [2, 4, 8])
@pytest.mark.parametrize("precond,nx,ny,tol", [(2, 1024, 256, 1e-14), (2, 1024, 256, 1e-10), (1, 256, 64, 1e-14)])
def test_local_slabs_match_single_context(P, precond, nx, ny, tol):
    """P slabs == one context: same f to 1e-10, same preconditioner (the spike correction makes the multigrid
    smoother's x-lines exact across the slabs), hence the same iteration counts up to round-off."""
    if nx // P < 16:
        pytest.skip("fewer than 16 rows per slab")
    cs = _case(nx, ny)
    chk = 1 if precond == 2 else 16
    f, sts = _run_slabs(P, cs, 3, precond=precond, tol=tol, check_every=chk)
    fref, st = _run_single(cs, 3, precond=precond, tol=tol)
    assert max_rel(f, fref) < max(1e-10, 100 * tol)
    for s in sts:
        assert s["negatives"] == 0 and s["resid_last"] <= 1000 * tol and s["precond"] == precond and s["steps"] == 3
        assert s["iters_total"] == sts[0]["iters_total"]          # every rank takes the same decisions
        if precond == 2:
            assert abs(s["iters_total"] - st["iters_total"]) <= (1 if tol > 1e-12 else 3), (s["iters_total"], st["iters_total"])
        else:
            assert abs(s["iters_total"] - st["iters_total"]) <= 0.2 * st["iters_total"]


def test_four_slabs_match_the_oracle():
    """192 x 64, P = 4, three steps against the NumPy/SuperLU restatement of the reference."""
    import ppfv_oracle as O
    cs = _case(192, 64)
    f, sts = _run_slabs(4, cs, 3, precond=2, check_every=1)
    m = O.Mesh(cs["xe"], cs["ye"], DT)
    eq = O.Equation(m)
    eq.G, eq.Dxx, eq.Dxy, eq.Dyy, eq.inv_tau = cs["G"], cs["Dxx"], cs["Dxy"], cs["Dyy"], cs["inv_tau"]
    eq.bc = list(cs["bct"]); eq.dirichlet_lines = lambda t: cs["lines"]; eq.init_f = lambda: cs["f0"]
    s = O.Solver(m, eq)
    for _ in range(3):
        s.update()
    assert max_rel(f, s.f) < 1e-9
    assert all(st["negatives"] == 0 for st in sts)


@pytest.mark.parametrize("P", [4, 8])
def test_one_step_at_4096_squared(P):
    """1-context-vs-P-slab agreement at 4096^2 plus the residual certificate."""
    cs = _case(4096, 4096)
    f, sts = _run_slabs(P, cs, 1, precond=2, check_every=1)
    fref, st = _run_single(cs, 1, precond=2)
    assert max_rel(f, fref) < 1e-10
    for s in sts:
        assert s["resid_last"] <= 1e-11 and s["negatives"] == 0
        assert abs(s["iters_total"] - st["iters_total"]) <= 3


def test_slab_failure_is_collective_and_commits_nothing():
    """maxit too small: every rank returns NOT_CONVERGED for the same step and keeps f of t^n."""
    cs = _case(256, 64)

    def work(rank, group):
        eng = sy.Engine(cs["xe"], cs["ye"], DT, slab=(rank, 2, group))
        eng.set_options(precond=2, maxit=2, check_every=1)
        lo, hi = eng.i_lo, eng.i_hi
        eng.set_coeffs(cs["G"][lo:hi], cs["Dxx"][lo:hi], cs["Dxy"][lo:hi], cs["Dyy"][lo:hi], cs["inv_tau"][lo:hi])
        eng.set_bc(cs["bct"], *cs["lines"])
        eng.set_f(cs["f0"][lo:hi])
        code = 0
        try:
            eng.step(1)
        except sy.Sy2dError as e:
            code = e.code
        same = np.array_equal(eng.get_f()[0], cs["f0"][lo:hi])
        cnt = eng.step_count()
        eng.set_options(precond=2, maxit=400, check_every=1)
        st = eng.step(1)
        eng.close()
        return code, same, cnt, st["steps"]
    for code, same, cnt, steps in run_local_slabs(2, work):
        assert code == -3 and same and cnt == 0 and steps == 1
