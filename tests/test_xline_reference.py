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
