"""Engine 2 (one problem per CTA / CTA pair): the variants of the ensemble kernel agree with one another and with the
reference's golden run.  The kernels under test replace Solver::update of the reference (source/Solver.cc:270-290) for
independent 80 x 80 problems; the golden arrays are what the reference produced (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from conftest import engine_from_golden, load_golden, max_rel

pytestmark = pytest.mark.gpu


def _ensemble(nbatch, **opts):
    """lc80 fields with D scaled per member (the bench's config-4 construction on a small batch)."""
    from sayram2d_b200 import Engine, fields
    from conftest import bc_for
    g = load_golden("lc80")
    a, b = fields.ensemble_scales(np.arange(nbatch) * 37 % 4096)
    eng = Engine(g["x_edges"], g["y_edges"], g["meta"]["dt"], nbatch=nbatch)
    if opts:
        eng.set_options(**opts)
    sc = lambda arr, s: arr[None] * s[:, None, None]
    one = np.ones(nbatch)
    eng.set_coeffs(sc(g["G"], one), sc(g["Dxx"], a), sc(g["Dxy"], a), sc(g["Dyy"], a), sc(g["inv_tau"], b))
    bct, lines = bc_for("LC", g["x_edges"], g["y_edges"])
    eng.set_bc(bct, *lines)
    eng.set_f(np.ascontiguousarray(sc(g["f_0"], one)))
    return eng


def test_predictor_settings_give_the_same_solution():
    """options.predictor only moves the initial guess of a step: 0 / 1 / 2 agree to the solver tolerance (times the
    conditioning), and the extrapolating predictor does not need more iterations than the plain ratio once it has a history."""
    out, its = {}, {}
    for pred in (0, 1, 2):
        eng = _ensemble(24, predictor=pred)
        eng.step(6)
        st = eng.step(12)
        assert st["engine"] == 2 and st["negatives"] == 0 and st["resid_last"] < 1e-12
        out[pred], its[pred] = eng.get_f(), st["iters_sum_all"]
        eng.close()
    assert max_rel(out[0], out[1]) < 1e-9 and max_rel(out[2], out[1]) < 1e-9
    assert its[1] < its[0] and its[2] <= its[1], its


@pytest.mark.parametrize("pred", [0, 1])
def test_full_run_parity_with_reference_for_the_other_predictors(pred):
    """500 steps of the reference's LC run (snapshots at t = 0.1, 0.5, 1 day) with predictor 0 and 1; the default, 2, is
    what test_gpu_parity.py::test_full_run_parity_with_reference runs."""
    g = load_golden("lc80")
    eng = engine_from_golden(g, "LC", predictor=pred)
    done = 0
    for k, upto in ((1, 50), (5, 250), (10, 500)):
        st = eng.step(upto - done)
        done = upto
        assert st["negatives"] == 0 and st["engine"] == 2
        assert max_rel(eng.get_f()[0], g[f"f_{k}"]) < 1e-8
    eng.close()


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("nbatch", [3, 80])
def test_cta_pair_kernel_matches_the_single_cta_kernel(variant, nbatch):
    """SY2D_XLINE_CLUSTER = 1 (pair of 320-thread CTAs) / 2 (pair of 640-thread CTAs): one 80 x 80 problem per 2-CTA
    cluster, halo columns and reduction partials through distributed shared memory.  Same arithmetic per cell, the
    reductions add in a different order: the solutions agree far below the solver tolerance's effect."""
    ref = _ensemble(nbatch)
    sr = ref.step(7)
    fr = ref.get_f()
    ref.close()
    os.environ["SY2D_XLINE_CLUSTER"] = str(variant)
    try:
        eng = _ensemble(nbatch)
    finally:
        del os.environ["SY2D_XLINE_CLUSTER"]
    st = eng.step(7)
    f = eng.get_f()
    eng.close()
    assert st["engine"] == 2 and st["negatives"] == 0 and st["resid_last"] < 1e-12
    assert max_rel(f, fr) < 1e-10
    assert abs(st["iters_sum_all"] - sr["iters_sum_all"]) <= 0.02 * sr["iters_sum_all"] + nbatch


def test_pair_kernel_results_do_not_depend_on_the_batch():
    """A problem's result must not depend on which pair of CTAs ran which of its time steps (work queue) or on its neighbours
    in the batch: bitwise equal between a 5-member and a 150-member launch."""
    os.environ["SY2D_XLINE_CLUSTER"] = "1"
    try:
        small = _ensemble(5)
        big = _ensemble(150)
    finally:
        del os.environ["SY2D_XLINE_CLUSTER"]
    small.step(5); big.step(5)
    fs, fb = small.get_f(), big.get_f()
    small.close(); big.close()
    assert np.array_equal(fs, fb[:5])
