"""What the engine does when a solve cannot succeed (the reference's direct LU cannot fail this way, so
the contract is this library's own, include/sayram2d.h): bad input is rejected, a NaN iterate is never
reported as converged, a failed step commits nothing and can be retried, and options changed between
calls take effect inside the captured iteration graphs.  Plus the engine-2 work queue: work items of one
time step handed from SM to SM give bit-identical results to one CTA per problem."""
import os

import numpy as np
import pytest

from conftest import bc_for, engine_from_golden, load_golden, max_rel

import sayram2d_b200 as sy
from sayram2d_b200 import fields

pytestmark = pytest.mark.gpu


def test_non_positive_or_non_finite_f_is_rejected():
    g = load_golden("ay80")
    eng = engine_from_golden(g, "AY")
    for bad in (0.0, -1.0e-30, np.nan, np.inf):
        f = g["f_0"].copy()
        f[17, 23] = bad
        for call in (eng.set_f, eng.put_f):
            with pytest.raises(sy.Sy2dError) as e:
                call(f)
            assert e.value.code == -1 and "finite and > 0" in str(e.value)
            with pytest.raises(sy.Sy2dError) as e:      # no usable f any more
                eng.step(1)
            assert e.value.code == -4
            eng.set_f(g["f_0"])
    eng.step(2)
    eng.close()


@pytest.mark.parametrize("engine,precond", [(1, 0), (1, 1), (1, 2), (2, 0), (2, 1)])
def test_nan_iterate_is_an_error_and_commits_nothing(engine, precond):
    """A NaN coefficient turns the Krylov vectors NaN in the first iteration; max|r| must not reduce to 0."""
    g = load_golden("ay80")
    eng = engine_from_golden(g, "AY", engine=engine, precond=precond)
    eng.step(2)
    f2 = eng.get_f()
    Dxx = g["Dxx"].copy()
    Dxx[40, 40] = np.nan
    eng.set_coeffs(g["G"], Dxx, g["Dxy"], g["Dyy"], g["inv_tau"])
    with pytest.raises(sy.Sy2dError) as e:
        eng.step(3)
    assert e.value.code == -3
    assert eng.last_stats["steps"] == 0 and eng.step_count() == 2
    assert np.array_equal(eng.get_f(), f2)                      # f of the last completed step, bit for bit
    eng.set_coeffs(g["G"], g["Dxx"], g["Dxy"], g["Dyy"], g["inv_tau"])
    eng.step(3)
    ref = engine_from_golden(g, "AY", engine=engine, precond=precond)
    ref.step(5)
    assert max_rel(eng.get_f(), ref.get_f()) < 1e-11 and eng.step_count() == 5
    eng.close(); ref.close()


@pytest.mark.parametrize("engine,precond", [(1, 0), (1, 2), (2, 1)])
def test_failed_step_can_be_retried(engine, precond):
    g = load_golden("lc80")
    eng = engine_from_golden(g, "LC", engine=engine, precond=precond)
    eng.step(4)
    f4 = eng.get_f()
    eng.set_options(maxit=2, check_every=1)
    with pytest.raises(sy.Sy2dError) as e:
        eng.step(2)
    assert e.value.code == -3 and eng.step_count() == 4 and np.array_equal(eng.get_f(), f4)
    eng.set_options(maxit=20000, check_every=16)
    eng.step(2)
    ref = engine_from_golden(g, "LC", engine=engine, precond=precond)
    ref.step(6)
    assert max_rel(eng.get_f(), ref.get_f()) < 1e-11
    eng.close(); ref.close()


def test_batch_member_failure_leaves_the_others_complete():
    """Engine 2: one member with a NaN coefficient fails at its first step and keeps its f; the other members finish."""
    g = load_golden("lc80")
    nb = 5
    rep = lambda a: np.broadcast_to(a, (nb,) + a.shape).copy()
    eng = sy.Engine(g["x_edges"], g["y_edges"], g["meta"]["dt"], nbatch=nb)
    Dyy = rep(g["Dyy"]); Dyy[3, 10, 10] = np.nan
    eng.set_coeffs(rep(g["G"]), rep(g["Dxx"]), rep(g["Dxy"]), Dyy, rep(g["inv_tau"]))
    bct, lines = bc_for("LC", g["x_edges"], g["y_edges"])
    eng.set_bc(bct, *lines)
    eng.set_f(rep(g["f_0"]))
    with pytest.raises(sy.Sy2dError) as e:
        eng.step(3)
    assert e.value.code == -3 and eng.last_stats["steps"] == 0 and eng.step_count() == 0
    f = eng.get_f()
    assert np.array_equal(f[3], g["f_0"])
    ref = engine_from_golden(g, "LC")
    ref.step(3)
    for m in (0, 1, 2, 4):
        assert max_rel(f[m], ref.get_f()[0]) < 1e-12
    eng.close(); ref.close()


def test_tolerance_change_reaches_the_captured_graphs():
    """Engine 1 replays CUDA graphs that hold tol / maxit by value: sy2d_set_options must rebuild them."""
    g = load_golden("syn64x48")
    eng = engine_from_golden(g, "AY", engine=1, precond=0, tol=1e-5, check_every=1)
    st = eng.step(1)
    loose = st["iters_total"]
    assert st["resid_last"] > 1e-9
    eng.set_options(tol=1e-14)
    st = eng.step(1)
    assert st["resid_last"] < 1e-13 and st["iters_total"] > loose
    eng.set_options(maxit=3)
    with pytest.raises(sy.Sy2dError) as e:
        eng.step(1)
    assert e.value.code == -3 and eng.last_stats["iters_last"] <= 3
    eng.close()


def _ensemble(nb, chunk):
    os.environ["SY2D_XLINE_CHUNK"] = str(chunk)
    try:
        lc = load_golden("lc80")
        a, b = fields.ensemble_scales(np.arange(nb) * 7 % 4096)
        sc = lambda arr, s: np.ascontiguousarray(arr[None] * s[:, None, None])
        one = np.ones(nb)
        e = sy.Engine(lc["x_edges"], lc["y_edges"], lc["meta"]["dt"], nbatch=nb)
    finally:
        del os.environ["SY2D_XLINE_CHUNK"]
    e.set_coeffs(sc(lc["G"], one), sc(lc["Dxx"], a), sc(lc["Dxy"], a), sc(lc["Dyy"], a), sc(lc["inv_tau"], b))
    bct, lines = bc_for("LC", lc["x_edges"], lc["y_edges"])
    e.set_bc(bct, *lines)
    e.set_f(sc(lc["f_0"], one))
    return e


def test_work_queue_is_bitwise_equal_to_one_cta_per_problem():
    """700 members (4.7 waves of 148 SMs), 7 time steps: items of 1 and of 3 steps migrate between SMs; the arithmetic
    per problem is unchanged, so f must be IDENTICAL to the run where a CTA keeps its problem for the whole call."""
    nb = 700
    outs, its = [], []
    for chunk in (0, 1, 3):
        e = _ensemble(nb, chunk)
        st = e.step(4)
        st2 = e.step(3)
        outs.append(e.get_f())
        its.append((st["iters_sum_all"] + st2["iters_sum_all"], st2["iters_last"], e.step_count()))
        assert st["negatives"] == 0 and st2["resid_last"] < 1e-12 and st["engine"] == 2 and st["precond"] == 1
        e.close()
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    assert its[0] == its[1] == its[2] and its[0][2] == 7


@pytest.mark.gpu
@pytest.mark.parametrize("precond", [1, 0])   # engine 2: x-line kernel, Jacobi kernel
def test_ensemble_statistics_report_the_true_minimum(precond):
    """stats.fmin of an ensemble call is the minimum of f after the call's last step, to the bit (the x-line kernel used to
    reduce 1e300 - f, which rounds every |f| < 1e284 to zero and reported -0.0)."""
    e = _ensemble(40, 1)
    o = e.options(); o.precond = precond
    e._check(e.lib.sy2d_set_options(e._ctx, o)); e._opt = o
    st = e.step(3)
    f = e.get_f()
    assert st["engine"] == 2 and st["negatives"] == 0 and int((f < 0).sum()) == 0
    assert st["fmin"] == float(f.min()) and st["fmin"] > 0.0
    e.close()
