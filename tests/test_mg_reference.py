"""CPU check of the multigrid algorithm itself (tests/mg_reference.py, the NumPy restatement the GPU
kernels are compared with): BiCGSTAB preconditioned by the V-cycle converges in a grid-independent
number of iterations to the solution of the reference's linear system."""
import numpy as np

import mg_reference as MG
import ppfv_oracle as O


def scaled_system(n_alpha, n_E):
    kw = dict(nalpha0=n_alpha, nE=n_E, alpha0min=5, alpha0max=90, Emin=0.2, Emax=5, T=1.0, nplots=10, nsteps=500)
    p, m, eq = O.build_case("SYN", None, None, **kw)
    s = O.Solver(m, eq)
    op = s.assemble()
    c = s.f
    om = op["diag"] * c
    w = [np.zeros_like(c) for _ in range(4)]
    w[0][1:] = op["W"][1:] * c[:-1]
    w[1][:-1] = op["E"][:-1] * c[1:]
    w[2][:, 1:] = op["S"][:, 1:] * c[:, :-1]
    w[3][:, :-1] = op["N"][:, :-1] * c[:, 1:]
    w = [a / om for a in w]
    s.update()
    return w, om, op["R"] / om, c, s.f


def test_level_count_rule():
    assert MG.level_count(1024, 1024) == 5 and MG.level_count(2048, 2048) == 6 and MG.level_count(80, 80) == 2
    assert MG.level_count(64, 48) == 2 and MG.level_count(64, 50) == 0 and MG.level_count(8192, 64) == 2 and MG.level_count(16384, 64) == 0
    assert MG.level_count(64, 16) == 2 and MG.level_count(1024, 1024, 3) == 3 and MG.level_count(256, 256) == 3


def test_vcycle_preconditioned_bicgstab_solves_the_reference_system():
    its = {}
    for n in (64, 128):
        w, om, R, c, f_next = scaled_system(n, n)
        levels = MG.hierarchy(*w, om, MG.level_count(n, n))
        rhs = R - levels[0].apply(np.ones_like(c))
        d, its[n] = MG.bicgstab_iterations(levels, rhs)
        assert np.max(np.abs(c * (1.0 + d) / f_next - 1.0)) < 1e-10
    assert its[64] <= 20 and its[128] <= 20, its


def test_spike_reduction_reproduces_the_global_line_solve():
    """NumPy restatement of the row-slab line coupling (sy2d_mg.cuh, k_mg_spike_reduced / k_mg_spike_apply): local
    solves g_r, spikes W_r, V_r, the block forward elimination / back substitution over the ranks on the tips, and
    the correction x_r = g_r - W_r bot_{r-1} - V_r top_{r+1} give the solution of the whole tridiagonal line."""
    rng = np.random.default_rng(7)
    for P, rows in ((2, 16), (4, 16), (8, 8), (3, 5)):
        n = P * rows
        wW = -rng.uniform(0.3, 0.499, n); wE = -rng.uniform(0.3, 0.499, n)
        wW[0] = 0.0; wE[-1] = 0.0
        T = np.eye(n) + np.diag(wW[1:], -1) + np.diag(wE[:-1], 1)
        b = rng.standard_normal(n)
        g, W, V = [], [], []
        for r in range(P):
            sl = slice(r * rows, (r + 1) * rows)
            e0 = np.zeros(rows); e0[0] = wW[r * rows]
            el = np.zeros(rows); el[-1] = wE[(r + 1) * rows - 1]
            g.append(np.linalg.solve(T[sl, sl], b[sl]))
            W.append(np.linalg.solve(T[sl, sl], e0))
            V.append(np.linalg.solve(T[sl, sl], el))
        al, be, ga, de = np.zeros(P), np.zeros(P), np.zeros(P), np.zeros(P)
        ap = bp = 0.0
        for r in range(P):   # bot_r = al + be top_{r+1}, top_r = ga + de top_{r+1}
            inv = 1.0 / (1.0 + W[r][0] * bp)
            ga[r] = (g[r][0] - W[r][0] * ap) * inv
            de[r] = -V[r][0] * inv
            al[r] = g[r][-1] - W[r][-1] * ap - W[r][-1] * bp * ga[r]
            be[r] = -W[r][-1] * bp * de[r] - V[r][-1]
            ap, bp = al[r], be[r]
        top, bot = np.zeros(P + 1), np.zeros(P)
        for r in range(P - 1, -1, -1):
            top[r] = ga[r] + de[r] * top[r + 1]
            bot[r] = al[r] + be[r] * top[r + 1]
        x = np.concatenate([g[r] - W[r] * (bot[r - 1] if r else 0.0) - V[r] * top[r + 1] for r in range(P)])
        assert np.max(np.abs(x - np.linalg.solve(T, b))) < 1e-13
