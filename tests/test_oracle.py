"""CPU tests of the oracle itself: the NumPy restatement (oracle/ppfv_oracle.py)
against the golden fixtures that the reference's own sources produced
(tests/golden/make_golden.py), plus the scheme's known-answer properties
 which follow from Solver.cc:99-164."""
import os

import numpy as np
import pytest

import h5min
import ppfv_oracle as O
from conftest import ROOT, load_golden, max_rel

DATA = os.path.join(ROOT, "data")


def test_h5_reader_matches_known_values(d_table):
    t = d_table
    assert t["alpha0"].shape == (91,) and t["E"].shape == (49,)
    assert t["Daa"].shape == t["Dap"].shape == t["Dpp"].shape == (91, 49)
    assert np.array_equal(t["alpha0"], np.arange(91.0))
    assert abs(t["E"][0] - 0.1) < 1e-15 and abs(t["E"][48] - 5.0) < 1e-14
    assert t["Daa"][0, 0] == 3.485e-05 and t["Daa"].max() == 3.146e-04
    assert t["Dap"].min() == -2.799e-05 and t["Dap"].max() == 2.08e-05 and t["Dpp"].max() == 2.934e-05


def test_parameters_follow_reference_rounding(tmp_path):
    ini = tmp_path / "q.ini"
    ini.write_text("[basic]\nrun_id = q\nNALPHA0 = 10\nnE = 12\nalpha0min = 5\nalpha0max = 90\nEmin = 0.2\nEmax = 5\n"
                   "T = 1.0\nnsteps = 505\n[diagnostics]\nnplots = 10\n[diffusion_coefficients]\ndID = X\n")
    p = O.Parameters(str(ini))
    # Parameters.cc:59-60: save_every = int(505/10) = 50, nsteps = 500, dt = T/500
    assert (p.save_every_step, p.nsteps, p.dt) == (50, 500, 1.0 / 500)
    assert p.nalpha0 == 10 and p.nE == 12 and p.dID == "X"
    assert abs(p.alpha0_min - 5 * O.gPI / 180) == 0


@pytest.mark.parametrize("tag,case,ini", [("ay80", "AY", "p.ini"), ("lc80", "LC", "p_AlbertYoungLC.ini")])
def test_case_fields_match_reference(tag, case, ini, d_table):
    g = load_golden(tag)
    p, m, eq = O.build_case(case, os.path.join(DATA, ini), d_table)
    assert np.array_equal(m.x_edges, g["x_edges"]) and np.array_equal(m.y_edges, g["y_edges"])
    for name in ("G", "Dxx", "Dxy", "Dyy", "inv_tau"):
        ref = g[name]
        assert np.max(np.abs(getattr(eq, name) - ref)) <= 4e-15 * np.max(np.abs(ref)) + 0.0, name
    assert max_rel(eq.init_f(), g["f_0"]) < 1e-13


@pytest.mark.parametrize("tag,case,ini", [("ay80", "AY", "p.ini"), ("lc80", "LC", "p_AlbertYoungLC.ini")])
def test_operator_matches_reference(tag, case, ini, d_table):
    g = load_golden(tag)
    p, m, eq = O.build_case(case, os.path.join(DATA, ini), d_table)
    s = O.Solver(m, eq)
    op = s.assemble()
    for k in ("diag", "W", "E", "S", "N", "R"):
        ref = g["op1_" + k]
        assert np.max(np.abs(op[k] - ref)) <= 1e-14 * np.max(np.abs(ref)), k


@pytest.mark.parametrize("tag,case,ini,tol", [("ay80", "AY", "p.ini", 1e-10), ("lc80", "LC", "p_AlbertYoungLC.ini", 2e-9)])
def test_full_run_matches_reference(tag, case, ini, tol, d_table):
    """500 steps of the restatement (SuperLU) vs the reference build (its own LU):
    two direct solvers differ by round-off amplified through the nonlinear scheme."""
    g = load_golden(tag)
    p, m, eq, snaps = O.run(case, os.path.join(DATA, ini), d_table)
    assert len(snaps) == 11
    for k in (1, 5, 10):
        assert max_rel(snaps[k], g[f"f_{k}"]) < tol
    assert (snaps[10] < 0).sum() == 0


def test_nonuniform_grid_matches_reference(d_table):
    g = load_golden("nu48x40")
    p, m, eq, snaps = O.run("AY", None, d_table, run_steps=20, stretch=0.6, nalpha0=48, nE=40, alpha0min=5,
                            alpha0max=90, Emin=0.2, Emax=5, T=1.0, nsteps=500, nplots=10)
    assert np.allclose(m.x_edges, g["x_edges"], rtol=0, atol=1e-15)
    assert max_rel(snaps[1], g["f_1"]) < 1e-11 and max_rel(snaps[20], g["f_20"]) < 1e-10


def test_synthetic_and_ensemble_match_reference(d_table):
    g = load_golden("syn64x48")
    kw = dict(nalpha0=64, nE=48, alpha0min=5, alpha0max=90, Emin=0.2, Emax=5, T=1.0, nplots=10)
    p, m, eq = O.build_case("SYN", None, None, nsteps=500, **kw)
    for name in ("Dxx", "Dxy", "Dyy", "inv_tau"):
        assert np.max(np.abs(getattr(eq, name) - g[name])) <= 1e-14 * np.max(np.abs(g[name]))
    s = O.Solver(m, eq)
    for _ in range(10):
        s.update()
    assert max_rel(s.f, g["f_10"]) < 1e-10
    ens = load_golden("ens_members")
    a, b = O.ensemble_member_scales(2047)
    p, m, eq, snaps = O.run("ENS", os.path.join(DATA, "p_AlbertYoungLC.ini"), d_table, run_steps=50, member=(a, b))
    assert max_rel(snaps[50], ens["f1_m2047"]) < 1e-9


# ---- scheme properties ------------------------------------
def _random_problem(seed, nx=17, ny=13, bc=(O.ZEROFLUX,) * 4, loss=False):
    rng = np.random.default_rng(seed)
    xe = np.concatenate([[0.0], np.cumsum(rng.uniform(0.5, 1.5, nx))])
    ye = np.concatenate([[0.0], np.cumsum(rng.uniform(0.5, 1.5, ny))])
    m = O.Mesh(xe, ye, 0.05)
    eq = O.Equation(m)
    eq.G = rng.uniform(0.5, 2.0, (nx, ny))
    eq.Dxx = rng.uniform(0.1, 3.0, (nx, ny))
    eq.Dyy = rng.uniform(0.1, 3.0, (nx, ny))
    eq.Dxy = rng.uniform(-0.9, 0.9, (nx, ny)) * np.sqrt(eq.Dxx * eq.Dyy)
    if loss:
        eq.inv_tau = rng.uniform(0.0, 4.0, (nx, ny))
    eq.bc = list(bc)
    lines = [rng.uniform(0.1, 1.0, ny + 1), rng.uniform(0.1, 1.0, ny + 1), rng.uniform(0.1, 1.0, nx + 1), rng.uniform(0.1, 1.0, nx + 1)]
    eq.dirichlet_lines = lambda t: lines
    f = rng.uniform(1e-6, 2.0, (nx, ny)) * 10.0 ** rng.uniform(-8, 0, (nx, ny))
    eq.init_f = lambda: f
    return m, eq


@pytest.mark.parametrize("bc", [(1, 1, 1, 1), (0, 1, 0, 0), (0, 0, 0, 0)])
def test_m_matrix_property(bc):
    m, eq = _random_problem(3, bc=bc, loss=True)
    s = O.Solver(m, eq)
    op = s.assemble()
    assert (op["diag"] > 0).all() and (op["R"] >= 0).all()
    for k in ("W", "E", "S", "N"):
        assert (op[k] <= 0).all()
    # strict COLUMN diagonal dominance: column K collects -A_K from its neighbours' rows
    col_off = np.zeros_like(op["diag"])
    col_off[:-1, :] += -op["W"][1:, :]
    col_off[1:, :] += -op["E"][:-1, :]
    col_off[:, :-1] += -op["S"][:, 1:]
    col_off[:, 1:] += -op["N"][:, :-1]
    assert (op["diag"] - col_off > 0).all()
    s.update()
    assert (s.f >= 0).all()


def test_constants_preserved_and_mass_conserved_with_zero_flux():
    m, eq = _random_problem(5)
    const = np.full((m.nx, m.ny), 0.37)
    eq.init_f = lambda: const
    s = O.Solver(m, eq, linear="banded")
    for _ in range(3):
        s.update()
    assert np.max(np.abs(s.f - 0.37)) < 1e-13
    m, eq = _random_problem(6)
    s = O.Solver(m, eq, linear="banded")
    vol = eq.G * m.dx[:, None] * m.dy[None, :]
    mass0 = (vol * s.f).sum()
    for _ in range(3):
        s.update()
    assert abs((vol * s.f).sum() - mass0) < 1e-12 * mass0


def test_reduces_to_five_point_backward_euler():
    nx, ny = 12, 9
    m = O.Mesh(np.linspace(0, 1.2, nx + 1), np.linspace(0, 0.9, ny + 1), 0.01)
    eq = O.Equation(m)
    eq.G[:] = 1.0; eq.Dxx[:] = 2.0; eq.Dyy[:] = 0.5
    rng = np.random.default_rng(0)
    f = rng.uniform(0.5, 1.5, (nx, ny))
    eq.init_f = lambda: f
    s = O.Solver(m, eq)
    op = s.assemble()
    hx, hy = m.dx[0], m.dy[0]
    assert np.allclose(op["W"][1:, :], -2.0 * hy / hx, rtol=1e-12)
    assert np.allclose(op["S"][:, 1:], -0.5 * hx / hy, rtol=1e-12)
    assert np.allclose(op["R"], hx * hy / 0.01 * f, rtol=1e-12)


def test_time_dependent_case_matches_reference(d_table):
    """Equation::update(t) path (Solver.cc:286-289): D(t), 1/tau(t) and the Dirichlet data(t) of the
    Time_Dependent user case, against the reference's own Solver run on the same class."""
    g = load_golden("td64")
    kw = dict(nalpha0=64, nE=64, alpha0min=5, alpha0max=90, Emin=0.2, Emax=5, T=1.0, nplots=10)
    p, m, eq = O.build_case("TD", None, d_table, nsteps=500, **kw)
    s = O.Solver(m, eq)
    assert max_rel(s.f, g["f_0"]) < 1e-13
    for k in (1, 2, 3):
        for _ in range(20):
            s.update()
        assert max_rel(s.f, g[f"f_{k}"]) < 1e-10, k
