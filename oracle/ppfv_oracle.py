"""CPU oracle for the Sayram-2D time-step hot path (NumPy/SciPy restatement).

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg may import this module; the product path
(sayram2d_b200/) never does.

PARITY PIN: the reference ships no tests or golden vectors,
so this restatement is pinned against the reference's own sources compiled
here (oracle/_ref, built by oracle/Makefile from /root/reference/source/*.cc
against shim headers); tests/golden/ holds the snapshots that build produced,
and tests/test_oracle.py checks this module against them.

Every function cites the reference file:line it restates (paths relative to
/root/reference/).  Arrays are (nx, ny), row-major, j (log E) fastest, exactly
the reference's xtensor layout (source/common.h:28).
"""
from __future__ import annotations

import math
import numpy as np

# source/common.h:38-44
gEPS = float(np.finfo(np.float64).eps)
gPI = 3.141592653589793238462
gC = 1.0
gE0 = 0.511875
gME = gE0 / (gC * gC)
gRE = 6371000.0

DIRICHLET, ZEROFLUX = 0, 1          # source/BCTypes.h:16
XMIN, XMAX, YMIN, YMAX = 0, 1, 2, 3  # source/BCTypes.h:13


# ----------------------------------------------------------------------------
# Parameters (source/Parameters.cc:35-65, source/Parameters.h:35-44)
# ----------------------------------------------------------------------------
class Parameters:
    def __init__(self, ini_path=None, **kw):
        if ini_path is not None:
            kw = {**read_ini(ini_path), **kw}
        self.run_id = kw.get("run_id", "run")
        self.nalpha0 = int(kw["nalpha0"])
        self.nE = int(kw["nE"])
        self.alpha0_min_deg = float(kw["alpha0min"])
        self.alpha0_max_deg = float(kw["alpha0max"])
        self.Emin = float(kw["Emin"])
        self.Emax = float(kw["Emax"])
        self.T = float(kw["T"])
        nsteps = float(kw["nsteps"])             # a double in the reference (Parameters.h:66)
        self.nplots = int(kw["nplots"])
        self.save_every_step = int(nsteps / self.nplots)       # Parameters.cc:59
        self.nsteps = self.save_every_step * self.nplots       # Parameters.cc:60
        self.dID = kw.get("dID", "AlbertYoung_chorus")

    # Parameters.h:35-44
    @property
    def alpha0_min(self): return self.alpha0_min_deg * gPI / 180
    @property
    def alpha0_max(self): return self.alpha0_max_deg * gPI / 180
    @property
    def logEmin(self): return math.log(self.Emin)
    @property
    def logEmax(self): return math.log(self.Emax)
    @property
    def dt(self): return self.T / self.nsteps


def read_ini(path):
    """Keys the reference reads (Parameters.cc:39-63); mINI semantics that matter
    here: case-insensitive keys, ';' comments, lines without '=' ignored."""
    out = {}
    alias = {"nalpha0": "nalpha0", "ne": "nE", "alpha0min": "alpha0min",
             "alpha0max": "alpha0max", "emin": "Emin", "emax": "Emax", "t": "T",
             "nsteps": "nsteps", "nplots": "nplots", "did": "dID", "run_id": "run_id"}
    with open(path) as fh:
        for line in fh:
            s = line.strip()
            if not s or s[0] in ";[" or "=" not in s:
                continue
            k, v = s.split("=", 1)
            k = k.strip().lower()
            if k in alias:
                out[alias[k]] = v.strip().split()[0] if v.strip() else ""
    return out


# ----------------------------------------------------------------------------
# Grid / Mesh (source/main.cc:20-37, source/Mesh.cc:36-65)
# ----------------------------------------------------------------------------
def make_uniform_edges(p, stretch=0.0):
    """main.cc:20-37.  stretch != 0 warps the interior edges exactly as
    oracle/ref_driver.cc make_grid does (non-uniform weights, Solver.cc:326-381)."""
    dx = (p.alpha0_max - p.alpha0_min) / float(p.nalpha0)
    dy = (p.logEmax - p.logEmin) / float(p.nE)
    xe = p.alpha0_min + dx * np.arange(p.nalpha0 + 1, dtype=np.float64)
    ye = p.logEmin + dy * np.arange(p.nE + 1, dtype=np.float64)
    if stretch != 0.0:
        def warp(e, s):
            a, L, n = e[0], e[-1] - e[0], e.size - 1
            xi = np.arange(1, n) / float(n)
            e[1:n] = a + L * (xi + s * np.sin(2.0 * gPI * xi) / (2.0 * gPI))
        warp(xe, stretch)
        warp(ye, -0.5 * stretch)
    return xe, ye


class Mesh:
    def __init__(self, x_edges, y_edges, dt):
        self.x_edges = np.asarray(x_edges, dtype=np.float64)
        self.y_edges = np.asarray(y_edges, dtype=np.float64)
        if self.x_edges.size < 2 or self.y_edges.size < 2:
            raise RuntimeError("Grid2D: edges must have size >= 2.")       # Grid2D.h:44-50
        if not (np.all(np.diff(self.x_edges) > 0) and np.all(np.diff(self.y_edges) > 0)):
            raise RuntimeError("Grid2D: edges must be strictly increasing")  # Grid2D.h:52-66
        self.nx = self.x_edges.size - 1
        self.ny = self.y_edges.size - 1
        self.dt = float(dt)
        self.dx = self.x_edges[1:] - self.x_edges[:-1]                    # Mesh.cc:47-51
        self.dy = self.y_edges[1:] - self.y_edges[:-1]
        self.x = 0.5 * (self.x_edges[:-1] + self.x_edges[1:])             # Mesh.cc:52
        self.y = 0.5 * (self.y_edges[:-1] + self.y_edges[1:])


# ----------------------------------------------------------------------------
# utils.h
# ----------------------------------------------------------------------------
def e2p(E, E0):   # utils.h:12-14
    return np.sqrt(E * (E + 2 * E0)) / gC


def p2e(p, E0):   # utils.h:7-9
    return np.sqrt(p * p * gC * gC + E0 * E0) - E0


def dlogE_dp(logE, E0):  # utils.h:17-20
    E = np.exp(logE)
    return e2p(E, gE0) * gC * gC / (E * (E + E0))


def _cal_weight(pos, n):
    """utils.h:39-51 with i a size_t: the `i<0` branch is dead, a negative
    floor wraps to a huge value and lands in `i>=n` (i=n-1, w=0)."""
    i = np.floor(pos).astype(np.int64)
    inside = (i >= 0) & (i < n)
    w = np.where(inside, 1.0 - (pos - i), 0.0)
    i = np.where(inside, i, n - 1)
    return i, w


def _interp2d(raw, i0, j0, wi, wj):  # utils.h:24-35
    return (raw[i0, j0] * wi * wj + raw[i0 + 1, j0] * (1 - wi) * wj
            + raw[i0 + 1, j0 + 1] * (1 - wi) * (1 - wj) + raw[i0, j0 + 1] * wi * (1 - wj))


# ----------------------------------------------------------------------------
# Equation cases
# ----------------------------------------------------------------------------
class Equation:
    """Field container mirroring source/Equation.h:27-75."""

    def __init__(self, mesh):
        self.m = mesh
        shp = (mesh.nx, mesh.ny)
        self.G = np.zeros(shp)
        self.Dxx = np.zeros(shp)
        self.Dyy = np.zeros(shp)
        self.Dxy = np.zeros(shp)
        self.inv_tau = np.zeros(shp)            # Equation.h:39-40
        self.bc = [ZEROFLUX] * 4

    def init_f(self):
        raise NotImplementedError

    def dirichlet_lines(self, t):
        """Returns the four boundary vertex lines (xmin[ny+1], xmax[ny+1],
        ymin[nx+1], ymax[nx+1]); None for a side that has no Dirichlet data
        (dirichlet_vertex_value returning false, Equation.h:60-65)."""
        return [None, None, None, None]

    def update(self, t):
        pass


def _ay_G(alpha, logE):  # Albert_Young.h:42-45
    t = 1.30 - 0.56 * np.sin(alpha)
    return e2p(np.exp(logE), gE0) ** 2 * t * np.sin(alpha) * np.cos(alpha) / dlogE_dp(logE, gE0)


def _construct_D(eq, table):
    """Albert_Young.cc:94-134 (identical in Albert_Young_LC.cc:108-148)."""
    m = eq.m
    x_D = table["alpha0"] * gPI / 180            # Albert_Young_IO.cc:22
    y_D = table["E"]
    nxD, nyD = x_D.size, y_D.size
    xminD, xmaxD, yminD, ymaxD = x_D[0], x_D[-1], y_D[0], y_D[-1]
    A, L = np.meshgrid(m.x, m.y, indexing="ij")
    pos_x = (A - xminD) / (xmaxD - xminD) * (nxD - 1)
    pos_y = (L - math.log(yminD)) / (math.log(ymaxD) - math.log(yminD)) * (nyD - 1)
    i0, wi = _cal_weight(pos_x, nxD - 1)
    j0, wj = _cal_weight(pos_y, nyD - 1)
    denorm = gME * gME * gC * gC
    s2d = 3600 * 24
    p = e2p(np.exp(L), gE0)
    dl = dlogE_dp(L, gE0)
    eq.Dxx = _interp2d(table["Daa"], i0, j0, wi, wj) * denorm * s2d / (p * p)
    eq.Dxy = _interp2d(table["Dap"], i0, j0, wi, wj) * denorm * s2d * dl / p
    eq.Dyy = _interp2d(table["Dpp"], i0, j0, wi, wj) * denorm * s2d * dl ** 2


class AlbertYoung(Equation):
    """source/Cases/Albert_Young.{h,cc}."""

    def __init__(self, paras, mesh, table):
        super().__init__(mesh)
        self.p = paras
        A, L = np.meshgrid(mesh.x, mesh.y, indexing="ij")
        self.G = _ay_G(A, L)                                          # Albert_Young.cc:29-40
        _construct_D(self, table)
        self.bc = [DIRICHLET, ZEROFLUX, DIRICHLET, DIRICHLET]         # Albert_Young.cc:42-59

    def _f0(self, a, logE):  # Albert_Young.h:37-40
        p = e2p(np.exp(logE), gE0)
        return np.exp(-(np.exp(logE) - 0.2) / 0.1) * (np.sin(a) - math.sin(5 * gPI / 180)) / (p * p) + gEPS

    def init_f(self):  # Albert_Young.cc:20-26
        A, L = np.meshgrid(self.m.x, self.m.y, indexing="ij")
        return self._f0(A, L)

    def dirichlet_lines(self, t):  # Albert_Young.cc:64-92, Albert_Young.h:52-62
        m = self.m
        return [np.zeros(m.ny + 1), None,
                self._f0(m.x_edges, self.p.logEmin), np.zeros(m.nx + 1)]


class TimeDependent(AlbertYoung):
    """The time-dependent user case of oracle/ref_driver.cc (Time_Dependent): AY with
    D(t) = D0 (1 + 0.5 sin(2 pi t / 0.1)), 1/tau(t) = 3 sin^2(pi t / 0.05) on the first quarter of the
    alpha0 rows, low-energy Dirichlet line x exp(-2 t).  Exercises Solver.cc:286-289."""

    def __init__(self, paras, mesh, table):
        super().__init__(paras, mesh, table)
        self.D0 = (self.Dxx.copy(), self.Dxy.copy(), self.Dyy.copy())
        self.update(0.0)

    def update(self, t):
        a = 1.0 + 0.5 * math.sin(2.0 * gPI * t / 0.1)
        s = math.sin(gPI * t / 0.05)
        self.Dxx, self.Dxy, self.Dyy = a * self.D0[0], a * self.D0[1], a * self.D0[2]
        self.inv_tau = np.zeros_like(self.Dxx)
        self.inv_tau[: self.m.nx // 4, :] = 3.0 * s * s

    def dirichlet_lines(self, t):
        lines = super().dirichlet_lines(t)
        lines[2] = lines[2] * math.exp(-2.0 * t)
        return lines


class AlbertYoungLC(Equation):
    """source/Cases/Albert_Young_LC.{h,cc}."""

    def __init__(self, paras, mesh, table):
        super().__init__(mesh)
        self.p = paras
        A, L = np.meshgrid(mesh.x, mesh.y, indexing="ij")
        self.G = _ay_G(A, L)
        Lsh = 4.5                                                      # Albert_Young_LC.cc:42
        a_lc = math.asin((Lsh ** 5 * (4 * Lsh - 3)) ** -0.25)          # :43
        pp = e2p(np.exp(L), gE0)
        tau_b = self._bounce_period(A, pp, Lsh)
        self.inv_tau = np.where(A < a_lc, 4.0 / tau_b, 0.0)            # :45-53
        _construct_D(self, table)
        self.bc = [ZEROFLUX, ZEROFLUX, DIRICHLET, DIRICHLET]          # :56-73

    @staticmethod
    def _bounce_period(a0, p, L):  # Albert_Young_LC.cc:150-160
        T0, T1 = 1.3802, 0.7405
        y = np.sin(a0)
        Ty = T0 - 0.5 * (T0 - T1) * (y + np.sqrt(y))
        return 4 * L * gRE * ((gE0 + p2e(p, gE0)) / (gC * gC)) / p * Ty / (3e8 * 3600 * 24)

    def _f0(self, a, logE):  # Albert_Young_LC.h:37-40
        p = e2p(np.exp(logE), gE0)
        return np.exp(-(np.exp(logE) - 0.2) / 0.1) * np.sin(a) / (p * p) + gEPS

    def init_f(self):
        A, L = np.meshgrid(self.m.x, self.m.y, indexing="ij")
        return self._f0(A, L)

    def dirichlet_lines(self, t):  # Albert_Young_LC.cc:78-106
        m = self.m
        return [None, None, self._f0(m.x_edges, self.p.logEmin), np.zeros(m.nx + 1)]


class SyntheticTensor(Equation):
    """BASELINE config 3: AY domain/BCs/G/f0 with a
    deterministic analytic full tensor and an f/tau loss strip.  Not a
    reference case: it plugs into the reference's Equation interface the same
    way a user case would (README.md:148-168)."""

    def __init__(self, paras, mesh):
        super().__init__(mesh)
        self.p = paras
        A, L = np.meshgrid(mesh.x, mesh.y, indexing="ij")
        self.G = _ay_G(A, L)
        xi = (A - mesh.x_edges[0]) / (mesh.x_edges[-1] - mesh.x_edges[0])
        eta = (L - mesh.y_edges[0]) / (mesh.y_edges[-1] - mesh.y_edges[0])
        self.Dxx = 10.0 * np.exp(-3.0 * eta) * (0.05 + np.sin(gPI * xi) ** 2)
        self.Dyy = 2.0 * np.exp(-2.0 * eta) * (0.05 + 4.0 * xi * (1.0 - xi))
        rho = 0.8 * np.sin(2.0 * gPI * xi) * np.cos(gPI * eta)
        self.Dxy = rho * np.sqrt(self.Dxx * self.Dyy)
        self.inv_tau = 5.0 * np.maximum(0.0, 1.0 - xi / 0.1)
        self.bc = [DIRICHLET, ZEROFLUX, DIRICHLET, DIRICHLET]
        self._ay = AlbertYoung.__new__(AlbertYoung)
        self._ay.m, self._ay.p = mesh, paras

    def init_f(self):
        return AlbertYoung.init_f(self._ay)

    def dirichlet_lines(self, t):
        return AlbertYoung.dirichlet_lines(self._ay, t)


# ----------------------------------------------------------------------------
# Solver (source/Solver.cc)
# ----------------------------------------------------------------------------
def update_Lambda(eq):
    """Solver.cc:57-65: Lambda = G * [[Dxx, Dxy], [Dxy, Dyy]]."""
    return eq.Dxx * eq.G, eq.Dxy * eq.G, eq.Dyy * eq.G


def fill_vertex_from_cells(m, f):
    """Solver.cc:299-383."""
    nx, ny = m.nx, m.ny
    v = np.empty((nx + 1, ny + 1))
    v[0, 0] = f[0, 0]; v[nx, 0] = f[nx - 1, 0]; v[0, ny] = f[0, ny - 1]; v[nx, ny] = f[nx - 1, ny - 1]
    xv = m.x_edges[1:nx][:, None]; yv = m.y_edges[1:ny][None, :]
    xL = m.x[:nx - 1][:, None]; xR = m.x[1:][:, None]
    yB = m.y[:ny - 1][None, :]; yT = m.y[1:][None, :]
    wxL = (xR - xv) / (xR - xL); wxR = (xv - xL) / (xR - xL)
    wyB = (yT - yv) / (yT - yB); wyT = (yv - yB) / (yT - yB)
    v[1:nx, 1:ny] = (wxL * wyB * f[:-1, :-1] + wxR * wyB * f[1:, :-1]
                     + wxL * wyT * f[:-1, 1:] + wxR * wyT * f[1:, 1:])          # :346
    w0 = (m.y[1:] - m.y_edges[1:ny]) / (m.y[1:] - m.y[:-1])
    w1 = (m.y_edges[1:ny] - m.y[:-1]) / (m.y[1:] - m.y[:-1])
    v[0, 1:ny] = w0 * f[0, :-1] + w1 * f[0, 1:]                                 # :361
    v[nx, 1:ny] = w0 * f[nx - 1, :-1] + w1 * f[nx - 1, 1:]                      # :364
    w0 = (m.x[1:] - m.x_edges[1:nx]) / (m.x[1:] - m.x[:-1])
    w1 = (m.x_edges[1:nx] - m.x[:-1]) / (m.x[1:] - m.x[:-1])
    v[1:nx, 0] = w0 * f[:-1, 0] + w1 * f[1:, 0]                                 # :377
    v[1:nx, ny] = w0 * f[:-1, ny - 1] + w1 * f[1:, ny - 1]                      # :380
    return v


def fill_vertex_from_bcs(m, eq, v, t):
    """Solver.cc:385-422; order XMIN, XMAX, YMIN, YMAX (later wins at corners)."""
    lines = eq.dirichlet_lines(t)
    for side in (XMIN, XMAX, YMIN, YMAX):
        if eq.bc[side] != DIRICHLET:
            continue
        if lines[side] is None:
            raise RuntimeError("Dirichlet BC: missing value.")                   # :393
        if side == XMIN: v[0, :] = lines[side]
        elif side == XMAX: v[m.nx, :] = lines[side]
        elif side == YMIN: v[:, 0] = lines[side]
        else: v[:, m.ny] = lines[side]
    return v


def _edge_geometry(m, inbr):
    """Mesh.cc:67-155: vertices A,B (coordinates and vertex indices), length and
    outward normal of face `inbr` (0=im/W, 1=jp/N, 2=ip/E, 3=jm/S) of every cell."""
    nx, ny = m.nx, m.ny
    I, J = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    if inbr == 0:   ai, aj, bi, bj, n = I, J + 1, I, J, (-1.0, 0.0)          # :94-97
    elif inbr == 1: ai, aj, bi, bj, n = I + 1, J + 1, I, J + 1, (0.0, 1.0)   # :112-115
    elif inbr == 2: ai, aj, bi, bj, n = I + 1, J, I + 1, J + 1, (1.0, 0.0)   # :130-133
    else:           ai, aj, bi, bj, n = I, J, I + 1, J, (0.0, -1.0)          # :148-151
    Ax, Ay = m.x_edges[ai], m.y_edges[aj]
    Bx, By = m.x_edges[bi], m.y_edges[bj]
    length = np.sqrt((Bx - Ax) ** 2 + (By - Ay) ** 2)                         # Edge.h:43
    return (ai, aj, Ax, Ay), (bi, bj, Bx, By), length, n


def a_sigma(m, Lam, vf, inbr):
    """Solver.cc:68-97 for every cell at once: returns a_A, a_B, a_sigma."""
    Lxx, Lxy, Lyy = Lam
    (ai, aj, Ax, Ay), (bi, bj, Bx, By), length, (n0, n1) = _edge_geometry(m, inbr)
    Kx = m.x[:, None]; Ky = m.y[None, :]
    vkbx, vkby = Bx - Kx, By - Ky
    vkax, vkay = Ax - Kx, Ay - Ky
    rbx, rby = vkby, -vkbx
    rax, ray = vkay, -vkax
    nLx = n0 * Lxx + n1 * Lxy          # n^T Lambda
    nLy = n0 * Lxy + n1 * Lyy
    aA = length * (nLx * rbx + nLy * rby) / (vkax * rbx + vkay * rby)         # :84
    aB = length * (nLx * rax + nLy * ray) / (vkbx * rax + vkby * ray)         # :85
    return aA, aB, aA * vf[ai, aj] + aB * vf[bi, bj]                          # :96


def assemble(m, eq, f, vf, Lam=None):
    """Solver.cc:167-267 -> the operator as five (nx,ny) diagonals + R.

    Returns dict(diag, W, E, S, N, R): row K=(i,j) of M is
      diag*f(i,j) + W*f(i-1,j) + E*f(i+1,j) + S*f(i,j-1) + N*f(i,j+1).
    """
    nx, ny = m.nx, m.ny
    if Lam is None:
        Lam = update_Lambda(eq)
    a = [a_sigma(m, Lam, vf, k) for k in range(4)]
    diag = np.zeros((nx, ny)); oW = np.zeros((nx, ny)); oE = np.zeros((nx, ny))
    oS = np.zeros((nx, ny)); oN = np.zeros((nx, ny)); R = np.zeros((nx, ny))

    def pair(aK, aL, fK, fL):
        """apply_inner_face_pair, Solver.cc:99-141 + calculate_mu Solver.h:63-67."""
        aAK, aBK, asK = aK
        aAL, aBL, asL = aL
        denom = np.abs(asK) + np.abs(asL) + 2 * gEPS
        muK = (np.abs(asL) + gEPS) / denom
        muL = 1 - muK
        B = muL * asL - muK * asK
        Babs = np.abs(B)
        Bp = (Babs + B) / 2.0
        Bm = (Babs - B) / 2.0
        AK = muK * (aAK + aBK) + Bp / (fK + gEPS)
        AL = muL * (aAL + aBL) + Bm / (fL + gEPS)
        return AK, AL

    # west faces, i>=1: K=(i,j) inbr=0, L=(i-1,j) rinbr=2          Solver.cc:173-175
    aK = tuple(q[1:, :] for q in a[0]); aL = tuple(q[:-1, :] for q in a[2])
    AK, AL = pair(aK, aL, f[1:, :], f[:-1, :])
    diag[1:, :] += AK; oW[1:, :] -= AL; diag[:-1, :] += AL; oE[:-1, :] -= AK      # :136-139
    # south faces, j>=1: K=(i,j) inbr=3, L=(i,j-1) rinbr=1         Solver.cc:178-182
    aK = tuple(q[:, 1:] for q in a[3]); aL = tuple(q[:, :-1] for q in a[1])
    AK, AL = pair(aK, aL, f[:, 1:], f[:, :-1])
    diag[:, 1:] += AK; oS[:, 1:] -= AL; diag[:, :-1] += AL; oN[:, :-1] -= AK

    def dirichlet(sl, inbr):
        """apply_dirichlet_face, Solver.cc:143-164."""
        aA, aB, asK = (q[sl] for q in a[inbr])
        B = -asK
        Babs = np.abs(B)
        Bp = (Babs + B) / 2.0
        Bm = (Babs - B) / 2.0
        diag[sl] += aA + aB + Bp / (f[sl] + gEPS)
        R[sl] += Bm

    # Solver.cc:204-267
    if eq.bc[XMIN] == DIRICHLET: dirichlet((0, slice(None)), 0)
    if eq.bc[XMAX] == DIRICHLET: dirichlet((nx - 1, slice(None)), 2)
    if eq.bc[YMIN] == DIRICHLET: dirichlet((slice(None), 0), 3)
    if eq.bc[YMAX] == DIRICHLET: dirichlet((slice(None), ny - 1), 1)

    UKK = eq.G * (m.dx[:, None] * m.dy[None, :] / m.dt)                        # :193, Mesh.h:61-63
    diag += UKK * (1 + m.dt * eq.inv_tau)                                       # :195
    R += UKK * f                                                                # :197
    return dict(diag=diag, W=oW, E=oE, S=oS, N=oN, R=R)


def to_sparse(op):
    """5 diagonals -> scipy CSC in the reference's numbering K = j*nx + i
    (Mesh.h:66-68), which is what Eigen's COLAMD sees."""
    import scipy.sparse as sp
    nx, ny = op["diag"].shape
    N = nx * ny
    K = (np.arange(ny)[None, :] * nx + np.arange(nx)[:, None])
    rows = [K.ravel()]; cols = [K.ravel()]; vals = [op["diag"].ravel()]
    for key, (di, dj) in (("W", (-1, 0)), ("E", (1, 0)), ("S", (0, -1)), ("N", (0, 1))):
        I, J = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
        ok = (I + di >= 0) & (I + di < nx) & (J + dj >= 0) & (J + dj < ny)
        rows.append(K[ok]); cols.append(((J + dj) * nx + (I + di))[ok]); vals.append(op[key][ok])
    return sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N, N))


def apply_operator(op, x):
    """y = M x with the 5-diagonal layout (used by tests as a residual check)."""
    y = op["diag"] * x
    y[1:, :] += op["W"][1:, :] * x[:-1, :]
    y[:-1, :] += op["E"][:-1, :] * x[1:, :]
    y[:, 1:] += op["S"][:, 1:] * x[:, :-1]
    y[:, :-1] += op["N"][:, :-1] * x[:, 1:]
    return y


class Solver:
    """source/Solver.{h,cc}: Solver(m, eq); update(); t(); f()."""

    def __init__(self, mesh, eq, linear="splu"):
        self.m, self.eq = mesh, eq
        self.istep = 0
        self.linear = linear
        self.f = np.array(eq.init_f(), dtype=np.float64)                        # Solver.cc:38-42
        self._refresh()

    def t(self):
        return self.istep * self.m.dt

    def _refresh(self):
        self.Lam = update_Lambda(self.eq)                                       # :288
        self.vf = fill_vertex_from_bcs(self.m, self.eq,
                                       fill_vertex_from_cells(self.m, self.f), self.t())  # :289

    def assemble(self):
        return assemble(self.m, self.eq, self.f, self.vf, self.Lam)

    def update(self):
        """Solver.cc:270-290."""
        op = self.assemble()
        nx, ny = self.m.nx, self.m.ny
        if self.linear == "splu":
            from scipy.sparse.linalg import splu
            lu = splu(to_sparse(op), permc_spec="COLAMD")    # Eigen SparseLU is a port of SuperLU
            sol = lu.solve(op["R"].T.ravel())                # R_(K), K = j*nx+i
            self.f = sol.reshape(ny, nx).T.copy()            # :280-284
        else:
            self.f = banded_solve(op)
        self.istep += 1
        self.eq.update(self.t())
        self._refresh()
        return op


def banded_solve(op):
    """Dense-band LU in (i*ny + j) numbering via LAPACK gbsv (pivoting is
    immaterial: M is strictly column diagonally dominant)."""
    from scipy.linalg import solve_banded
    nx, ny = op["diag"].shape
    N = nx * ny
    ab = np.zeros((2 * ny + 1, N))
    ab[ny, :] = op["diag"].ravel()
    # entry (r, c) is stored at ab[ny + r - c, c]
    w = op["W"].ravel(); e = op["E"].ravel(); s = op["S"].ravel(); n = op["N"].ravel()
    idx = np.arange(N)
    c = idx - ny; ok = c >= 0;                     ab[2 * ny, c[ok]] = w[ok]
    c = idx + ny; ok = c < N;                      ab[0, c[ok]] = e[ok]
    ok = (idx % ny) != 0;                          ab[ny + 1, idx[ok] - 1] = s[ok]
    ok = (idx % ny) != ny - 1;                     ab[ny - 1, idx[ok] + 1] = n[ok]
    return solve_banded((ny, ny), ab, op["R"].ravel()).reshape(nx, ny)


def build_case(case, ini_path=None, table=None, stretch=0.0, member=None, **overrides):
    """Parameters -> make_uniform -> Mesh -> Equation, as main.cc:41-51 does.
    case "ENS" = BASELINE config 4 member: LC with D scaled by member[0] and
    inv_tau by member[1]."""
    p = Parameters(ini_path, **overrides)
    xe, ye = make_uniform_edges(p, stretch)
    m = Mesh(xe, ye, p.dt)
    if case == "AY":
        eq = AlbertYoung(p, m, table)
    elif case == "LC":
        eq = AlbertYoungLC(p, m, table)
    elif case == "SYN":
        eq = SyntheticTensor(p, m)
    elif case == "TD":
        eq = TimeDependent(p, m, table)
    elif case == "ENS":
        eq = AlbertYoungLC(p, m, table)
        a, b = member
        eq.Dxx = eq.Dxx * a; eq.Dxy = eq.Dxy * a; eq.Dyy = eq.Dyy * a
        eq.inv_tau = eq.inv_tau * b
    else:
        raise ValueError(case)
    return p, m, eq


def ensemble_member_scales(mth):
    """a_m = 0.1*100^((m mod 64)/63), b_m = (m div 64)/63."""
    return 0.1 * 100.0 ** ((mth % 64) / 63.0), (mth // 64) / 63.0


def run(case, ini_path=None, table=None, run_steps=None, linear="splu", stretch=0.0, member=None, **overrides):
    """main.cc:41-85: returns snapshots /f/0 ... /f/nplots (or every step if run_steps given)."""
    p, m, eq = build_case(case, ini_path, table, stretch=stretch, member=member, **overrides)
    s = Solver(m, eq, linear=linear)
    snaps = [s.f.copy()]
    total = p.nsteps if run_steps is None else run_steps
    for tstep in range(1, total + 1):
        s.update()
        if run_steps is not None or tstep % p.save_every_step == 0:
            snaps.append(s.f.copy())
    return p, m, eq, snaps
