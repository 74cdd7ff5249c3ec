"""CPU check of the ensemble kernel's algorithm (tests/xline_reference.py): the pivot-scaled, x-line preconditioned
BiCGSTAB with the deferred solution update reproduces the reference's direct solve of one time step."""
import numpy as np

import xline_reference as XL
from test_mg_reference import scaled_system


def test_pivot_scaled_xline_bicgstab_reproduces_the_reference_step():
    for n in (48, 80):
        w, om, rhs0, c, f_next = scaled_system(n, n)
        rhs = rhs0 - XL.apply_A(w, np.ones_like(rhs0))      # residual of the guess f^{n+1} = c
        x, its, res, dinv, lp, e = XL.solve(w, rhs)
        assert res <= 1e-14 and its < 40                     # unscaled true residual; ~15-23 iterations
        f = c * (1.0 + x)
        assert np.max(np.abs(f - f_next) / np.abs(f_next)) < 1e-10   # the reference's LU result
        # unit-diagonal factors: diagonal of (I + L')(I + U') is 1/d, and 0 < d <= 1 for the M-matrix
        eprev = np.zeros_like(e); eprev[1:] = e[:-1]
        assert np.max(np.abs(dinv - (1.0 + lp * eprev))) < 1e-14
        assert np.all(dinv >= 1.0 - 1e-15) and np.max(dinv) < 2.0


def _ensemble_member_system(member, steps_before=2):
    """Scaled system of ensemble member `member` (BASELINE config 4: LC fields, D x a_m, 1/tau x b_m) after a few steps."""
    import ppfv_oracle as O
    from conftest import bc_for, load_golden
    from sayram2d_b200 import fields
    g = load_golden("lc80")
    a, b = fields.ensemble_scales(np.array([member]))
    m = O.Mesh(g["x_edges"], g["y_edges"], g["meta"]["dt"])
    eq = O.Equation(m)
    eq.G, eq.Dxx, eq.Dxy, eq.Dyy, eq.inv_tau = g["G"], g["Dxx"] * a[0], g["Dxy"] * a[0], g["Dyy"] * a[0], g["inv_tau"] * b[0]
    bct, lines = bc_for("LC", g["x_edges"], g["y_edges"])
    eq.bc = list(bct); eq.dirichlet_lines = lambda t: lines; eq.init_f = lambda: g["f_0"]
    s = O.Solver(m, eq)
    for _ in range(steps_before):
        s.update()
    op, c = s.assemble(), s.f
    om = op["diag"] * c
    w = [np.zeros_like(c) for _ in range(4)]
    w[0][1:] = op["W"][1:] * c[:-1]; w[1][:-1] = op["E"][:-1] * c[1:]
    w[2][:, 1:] = op["S"][:, 1:] * c[:, :-1]; w[3][:, :-1] = op["N"][:, :-1] * c[:, 1:]
    w = [x / om for x in w]
    s.update()
    return w, op["R"] / om, c, s.f


def test_pivot_scaled_solver_on_the_extreme_ensemble_members():
    """Members 0 (weakest diffusion, no loss) and 4095 (D x 10, full loss): the stiff one needs ~50 iterations and its pivots
    go down to 0.43; the scaled stopping rule still delivers the unscaled residual and the reference's f."""
    for member, max_its in ((0, 15), (4095, 70)):
        w, rhs0, c, f_next = _ensemble_member_system(member)
        rhs = rhs0 - XL.apply_A(w, np.ones_like(rhs0))
        x, its, res, dinv, lp, e = XL.solve(w, rhs)
        assert res <= 1e-14 and its <= max_its
        assert np.max(np.abs(c * (1.0 + x) - f_next) / np.abs(f_next)) < 1e-9
        assert 1.0 - 1e-15 <= np.min(dinv) and np.max(dinv) < 4.0      # 0 < d <= 1: r = d r' never exceeds r'
