"""GPU tests of the multigrid preconditioner (SY2D_PRECOND_MG, engine 1) through the C ABI:
the V-cycle against its NumPy restatement (tests/mg_reference.py) on the operator the engine
itself assembled, iteration counts, and agreement of the time steps with the Jacobi iteration
and the reference's direct solve."""
import numpy as np
import pytest

from conftest import max_rel

import mg_reference as MG
import sayram2d_b200 as sy
from sayram2d_b200 import fields

pytestmark = pytest.mark.gpu


def synthetic_engine(nx, ny, nbatch=1, **opts):
    xe, ye = fields.uniform_edges(nx, ny)
    eng = sy.Engine(xe, ye, 0.002, nbatch=nbatch)
    eng.set_options(engine=1, **opts)
    G = fields.ay_G(xe, ye)
    Dxx, Dxy, Dyy, inv_tau = fields.synthetic_tensor(xe, ye)
    rep = lambda a: np.broadcast_to(a, (nbatch,) + a.shape).copy()
    scale = (1.0 + 0.5 * np.arange(nbatch))[:, None, None]
    eng.set_coeffs(rep(G), rep(Dxx) * scale, rep(Dxy) * scale, rep(Dyy) * scale, rep(inv_tau))
    f0, bct, lines = fields.ay_init_and_bc(xe, ye)
    eng.set_bc(bct, *lines)
    eng.set_f(rep(f0))
    return eng


@pytest.mark.parametrize("nx,ny,nbatch", [(128, 64, 1), (100, 48, 2), (37, 16, 1), (1024, 128, 1), (1100, 32, 1), (2048, 32, 1), (520, 24, 1), (4096, 16, 1), (2500, 16, 1), (8192, 32, 1), (2048, 64, 2)])
def test_vcycle_matches_numpy_restatement(nx, ny, nbatch):
    eng = synthetic_engine(nx, ny, nbatch)
    eng.step(2)      # a developed f (predictor column scale active)
    rng = np.random.default_rng(nx * 1000 + ny)
    r = rng.standard_normal((nbatch, nx, ny))
    z, w4, om = eng.debug_vcycle(r)
    nlev = MG.level_count(nx, ny)
    assert nlev >= 2
    for b in range(nbatch):
        levels = MG.hierarchy(w4[0, b], w4[1, b], w4[2, b], w4[3, b], om[b], nlev)
        ref = MG.vcycle(levels, r[b])
        err = np.max(np.abs(z[b] - ref)) / np.max(np.abs(ref))
        assert err < 1e-10, (b, err)
    eng.close()


def test_iterations_are_grid_independent_and_results_agree():
    its = {}
    for n in (128, 256, 512):
        mg = synthetic_engine(n, n, precond=2)
        st = mg.step(3)
        assert st["precond"] == 2 and st["engine"] == 1 and st["negatives"] == 0 and st["resid_last"] < 1e-12
        its[n] = st["iters_last"]
        if n == 128:
            ja = synthetic_engine(n, n, precond=0)
            sj = ja.step(3)
            assert sj["precond"] == 0 and sj["iters_last"] > 5 * st["iters_last"]
            assert max_rel(mg.get_f(), ja.get_f()) < 1e-10
            ja.close()
        mg.close()
    assert max(its.values()) <= 30, its


def test_mg_levels_option_and_unsupported_grids():
    eng = synthetic_engine(64, 64, precond=2, mg_levels=2)
    st2 = eng.step(1)
    eng.close()
    eng = synthetic_engine(64, 64, precond=2)
    st5 = eng.step(1)
    eng.close()
    assert st2["precond"] == 2 and st5["precond"] == 2 and max(st2["iters_last"], st5["iters_last"]) <= 20
    eng = synthetic_engine(64, 50, precond=2)   # ny not a multiple of 4
    with pytest.raises(sy.Sy2dError):
        eng.step(1)
    eng.set_options(engine=1, precond=-1)       # AUTO falls back to the segmented x-line preconditioner
    assert eng.step(1)["precond"] == 1
    eng.close()


@pytest.mark.parametrize("case,nx,ny,stretch", [("AY", 160, 96, 0.6), ("LC", 128, 128, 0.0), ("SYN", 192, 64, -0.4)])
def test_mg_time_steps_match_the_oracle_on_mid_size_grids(case, nx, ny, stretch):
    """Multigrid-preconditioned engine 1 against the oracle's direct solve (SciPy SuperLU) at sizes between the
    reference's 80 x 80 fixtures and the 1024 x 1024 sub-sampled one: warped (non-uniform) grids, the loss-cone
    case with its zero-flux alpha0 = 0 boundary and 49 decades of f, the synthetic tensor with cross terms."""
    import os
    from conftest import ROOT
    import h5min
    import ppfv_oracle as O
    table = h5min.load_d_table(os.path.join(ROOT, "data", "D", "AlbertYoung_chorus.h5"))
    kw = dict(nalpha0=nx, nE=ny, alpha0min=0 if case == "LC" else 5, alpha0max=90, Emin=0.2, Emax=5, T=1.0, nplots=10, nsteps=500)
    p, m, eq = O.build_case(case, None, table, stretch=stretch, **kw)
    ref = O.Solver(m, eq)
    eng = sy.Engine(m.x_edges, m.y_edges, m.dt)
    eng.set_options(engine=1, precond=2)
    eng.set_coeffs(eq.G, eq.Dxx, eq.Dxy, eq.Dyy, eq.inv_tau)
    eng.set_bc(eq.bc, *eq.dirichlet_lines(0.0))
    eng.set_f(ref.f)
    for _ in range(6):
        ref.update()
    st = eng.step(6)
    assert st["precond"] == 2 and st["negatives"] == 0 and st["iters_last"] <= 30
    assert max_rel(eng.get_f()[0], ref.f) < 1e-9
    eng.close()


def test_failed_multigrid_step_is_redone_with_the_xline_iteration():
    """AUTO contexts: a multigrid-preconditioned solve that stops without converging is redone from the same f
    with the segmented x-line iteration (reserved[2] = 1 makes every first attempt count as failed)."""
    a = synthetic_engine(128, 128)            # AUTO -> multigrid
    b = synthetic_engine(128, 128)
    o = b.options(); o.engine = 1; o.reserved[2] = 1
    b._check(b.lib.sy2d_set_options(b._ctx, o)); b._opt = o
    sa, sb = a.step(3), b.step(3)
    assert sa["precond"] == 2 and sa["restarts_total"] == 0
    assert sb["precond"] == 1 and sb["restarts_total"] == 3 and sb["iters_last"] > 3 * sa["iters_last"]
    assert max_rel(a.get_f(), b.get_f()) < 1e-10 and sb["negatives"] == 0
    o.reserved[2] = 0
    b._check(b.lib.sy2d_set_options(b._ctx, o))
    sc = b.step(1)                            # back to multigrid on the next call
    assert sc["precond"] == 2 and sc["restarts_total"] == 0 and sc["iters_last"] <= sa["iters_last"] + 3
    a.close(); b.close()


@pytest.mark.parametrize("n,batch", [(256, 1), (1024, 1), (512, 3)])
def test_fused_coarse_tail_equals_the_separate_launches(n, batch):
    """SY2D_MG_TAIL_NY: the coarse levels of the V-cycle as the stages of ONE kernel (k_mg_tail, a barrier over the CTAs of a
    problem between the stages) - same arithmetic as the separate launches: the same f and the same iteration counts."""
    import os
    out = []
    for tail in ("0", "256"):
        os.environ["SY2D_MG_TAIL_NY"] = tail
        try:
            eng = synthetic_engine(n, n, nbatch=batch)
        finally:
            del os.environ["SY2D_MG_TAIL_NY"]
        eng.set_options(engine=1, precond=2)
        st = eng.step(3)
        out.append((eng.get_f(), st["iters_total"], st["kernel_launches"]))
        assert st["negatives"] == 0 and st["precond"] == 2
        eng.close()
    # (the fine-level dot products are summed with atomics: two runs agree to round-off, not bit for bit)
    assert max_rel(out[1][0], out[0][0]) < 1e-11 and abs(out[0][1] - out[1][1]) <= 2
    assert out[1][2] < out[0][2]      # fewer launches with the fused tail


@pytest.mark.parametrize("n,batch", [(256, 1), (1024, 1), (512, 2)])
def test_prefetching_line_kernel_equals_the_three_phase_kernel(n, batch):
    """SY2D_MG_LINE_PRE: the line kernel that stages the backward factors and the old iterate by cp.async into shared memory and
    scans the carries on all warps does the same arithmetic per row; only the association of the carry composition across blocks of
    32 segments differs (round-off): the same f and the same iteration counts as the kernel with three exposed load phases."""
    import os
    out = []
    for pre in ("0", "1"):
        os.environ["SY2D_MG_LINE_PRE"] = pre
        try:
            eng = synthetic_engine(n, n, nbatch=batch)
        finally:
            del os.environ["SY2D_MG_LINE_PRE"]
        eng.set_options(engine=1, precond=2)
        st = eng.step(3)
        out.append((eng.get_f(), st["iters_total"]))
        assert st["negatives"] == 0 and st["precond"] == 2
        eng.close()
    assert max_rel(out[1][0], out[0][0]) < 1e-11 and abs(out[0][1] - out[1][1]) <= 2
