"""Analytic case data for the engine: grids, G, f0, boundary lines and the
synthetic tensors of the benchmark configurations (BASELINE.json configs 3-5,
SURVEY.md section 8d).  NumPy host code; the table-driven Albert & Young D lives in
the C++ host layer (sayram2d_b200/host/Albert_Young.cc), as in the reference.
"""
from __future__ import annotations

import math

import numpy as np

# source/common.h:38-44
gEPS = float(np.finfo(np.float64).eps)
gPI = 3.141592653589793238462
gE0 = 0.511875
DIRICHLET, ZEROFLUX = 0, 1


def uniform_edges(nx, ny, alpha0min_deg=5.0, alpha0max_deg=90.0, Emin=0.2, Emax=5.0):
    """make_uniform of source/main.cc:20-37 (alpha0 in rad, y = log E)."""
    a0, a1 = alpha0min_deg * gPI / 180, alpha0max_deg * gPI / 180
    y0, y1 = math.log(Emin), math.log(Emax)
    xe = a0 + (a1 - a0) / float(nx) * np.arange(nx + 1, dtype=np.float64)
    ye = y0 + (y1 - y0) / float(ny) * np.arange(ny + 1, dtype=np.float64)
    return xe, ye


def centres(e):
    return 0.5 * (e[:-1] + e[1:])


def _e2p(E):  # utils.h:12-14 with gC = 1
    return np.sqrt(E * (E + 2 * gE0))


def _dlogE_dp(logE):  # utils.h:17-20
    E = np.exp(logE)
    return _e2p(E) / (E * (E + gE0))


def ay_G(xe, ye, rows=None):
    """Albert_Young.h:42-45 on cell centres -> (nx, ny); rows=(lo, hi) restricts to a row slab."""
    xc = centres(xe) if rows is None else centres(xe)[rows[0]:rows[1]]
    A, L = np.meshgrid(xc, centres(ye), indexing="ij")
    t = 1.30 - 0.56 * np.sin(A)
    return _e2p(np.exp(L)) ** 2 * t * np.sin(A) * np.cos(A) / _dlogE_dp(L)


def ay_f0(a, logE, loss_cone_deg=5.0):
    """Albert_Young.h:37-40 (loss_cone_deg=5) / Albert_Young_LC.h:37-40 (loss_cone_deg=None)."""
    p = _e2p(np.exp(logE))
    s0 = math.sin(loss_cone_deg * gPI / 180) if loss_cone_deg is not None else 0.0
    return np.exp(-(np.exp(logE) - 0.2) / 0.1) * (np.sin(a) - s0) / (p * p) + gEPS


def ay_init_and_bc(xe, ye, lc=False, rows=None):
    """Initial f on cell centres and the boundary description of the AY / LC cases
    (Albert_Young.cc:42-92, Albert_Young_LC.cc:56-106): returns f0, bc_type[4], lines[4].
    rows=(lo, hi) restricts f0 to a row slab (the boundary lines stay global)."""
    xc = centres(xe) if rows is None else centres(xe)[rows[0]:rows[1]]
    A, L = np.meshgrid(xc, centres(ye), indexing="ij")
    cone = None if lc else 5.0
    f0 = ay_f0(A, L, cone)
    ymin = ay_f0(xe, ye[0], cone)
    ymax = np.zeros(xe.size)
    if lc:
        return f0, [ZEROFLUX, ZEROFLUX, DIRICHLET, DIRICHLET], [None, None, ymin, ymax]
    return f0, [DIRICHLET, ZEROFLUX, DIRICHLET, DIRICHLET], [np.zeros(ye.size), None, ymin, ymax]


def synthetic_tensor(xe, ye, rows=None):
    """BASELINE config 3 / 5: deterministic analytic full tensor with a sign-changing
    cross term and an f/tau loss strip.  Returns Dxx, Dxy, Dyy, inv_tau;
    rows=(lo, hi) restricts to a row slab (normalised coordinates stay global)."""
    xc = centres(xe) if rows is None else centres(xe)[rows[0]:rows[1]]
    A, L = np.meshgrid(xc, centres(ye), indexing="ij")
    xi = (A - xe[0]) / (xe[-1] - xe[0])
    eta = (L - ye[0]) / (ye[-1] - ye[0])
    Dxx = 10.0 * np.exp(-3.0 * eta) * (0.05 + np.sin(gPI * xi) ** 2)
    Dyy = 2.0 * np.exp(-2.0 * eta) * (0.05 + 4.0 * xi * (1.0 - xi))
    rho = 0.8 * np.sin(2.0 * gPI * xi) * np.cos(gPI * eta)
    Dxy = rho * np.sqrt(Dxx * Dyy)
    inv_tau = 5.0 * np.maximum(0.0, 1.0 - xi / 0.1)
    return Dxx, Dxy, Dyy, inv_tau


def ensemble_scales(m):
    """BASELINE config 4 member m: a_m scales D ("L" axis), b_m scales 1/tau ("MLT" axis)."""
    m = np.asarray(m)
    return 0.1 * 100.0 ** ((m % 64) / 63.0), (m // 64) / 63.0
