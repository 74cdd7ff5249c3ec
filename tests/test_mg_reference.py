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
    assert MG.level_count(64, 48) == 2 and MG.level_count(64, 50) == 0 and MG.level_count(8192, 64) == 0
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
