"""NumPy restatement of the multigrid V-cycle of sayram2d_b200/csrc/sy2d_mg.cuh (tests only).

Same hierarchy (pairs along j, unit-diagonal levels with row weights om, theta = 1/2 rescaling of the
inter-aggregate couplings), same cycle (V(1,1), damped whole-x-line Jacobi, omega = 0.7, 2 sweeps on
the coarsest level); written with whole-array operations, so it shares no code and no evaluation
order with the kernels.  Arrays are (nx, ny) for one problem."""
import numpy as np

OMEGA, THETA, COARSE_SWEEPS = 0.7, 0.5, 2


def level_count(nx, ny, want=0):
    """The engine's rule: coarsen to <= 64 columns (at least two levels); want > 0 caps the count instead."""
    if nx > 8192 or nx < 8 or ny % 4 or ny < 16:
        return 0
    nlev, cap = 1, (want if want > 0 else 8)
    while nlev < cap and ny % 4 == 0 and ny >= 16 and (want > 0 or ny > 64 or nlev < 2):
        ny //= 2
        nlev += 1
    return nlev if nlev >= 2 else 0


class Level:
    def __init__(self, wW, wE, wS, wN, om):
        self.wW, self.wE, self.wS, self.wN, self.om = wW, wE, wS, wN, om
        nx = wW.shape[0]
        self.l = np.zeros_like(wW)
        self.dinv = np.zeros_like(wW)
        d = np.ones(wW.shape[1])
        self.dinv[0] = 1.0
        for i in range(1, nx):
            self.l[i] = wW[i] / d
            d = 1.0 - self.l[i] * wE[i - 1]
            self.dinv[i] = 1.0 / d

    def apply(self, x):
        y = x.copy()
        y[1:] += self.wW[1:] * x[:-1]
        y[:-1] += self.wE[:-1] * x[1:]
        y[:, 1:] += self.wS[:, 1:] * x[:, :-1]
        y[:, :-1] += self.wN[:, :-1] * x[:, 1:]
        return y

    def line(self, b):
        nx = b.shape[0]
        z = b.copy()
        for i in range(1, nx):
            z[i] -= self.l[i] * z[i - 1]
        z[nx - 1] *= self.dinv[nx - 1]
        for i in range(nx - 2, -1, -1):
            z[i] = (z[i] - self.wE[i] * z[i + 1]) * self.dinv[i]
        return z

    def coarsen(self):
        a, b = slice(0, None, 2), slice(1, None, 2)
        om, wS, wN = self.om, self.wS, self.wN
        s_full = om[:, a] * wS[:, a]        # row J, column J-1
        n_full = om[:, b] * wN[:, b]        # row J, column J+1
        d = om[:, a] * (1.0 + wN[:, a]) + om[:, b] * (1.0 + wS[:, b])
        d[:, :-1] += (1.0 - THETA) * s_full[:, 1:]
        d[:, 1:] += (1.0 - THETA) * n_full[:, :-1]
        cW = (om[:, a] * self.wW[:, a] + om[:, b] * self.wW[:, b]) / d
        cE = (om[:, a] * self.wE[:, a] + om[:, b] * self.wE[:, b]) / d
        return Level(cW, cE, THETA * s_full / d, THETA * n_full / d, d)


def hierarchy(wW, wE, wS, wN, om, nlev):
    levels = [Level(wW, wE, wS, wN, om)]
    while len(levels) < nlev:
        levels.append(levels[-1].coarsen())
    return levels


def vcycle(levels, r, k=0):
    L = levels[k]
    z = OMEGA * L.line(r)
    if k == len(levels) - 1:
        for _ in range(COARSE_SWEEPS - 1):
            z = z + OMEGA * L.line(r - L.apply(z))
        return z
    res = L.om * (r - L.apply(z))
    rc = (res[:, 0::2] + res[:, 1::2]) / levels[k + 1].om
    z = z + np.repeat(vcycle(levels, rc, k + 1), 2, axis=1)
    return z + OMEGA * L.line(r - L.apply(z))


def bicgstab_iterations(levels, rhs, tol=1e-14, maxit=200):
    """Right-preconditioned BiCGSTAB exactly as the engine runs it; returns (x, iterations)."""
    A = levels[0]
    x = np.zeros_like(rhs)
    r = rhs.copy()
    rho = alpha = omega = 1.0
    p = v = None
    for it in range(1, maxit + 1):
        rho_new = np.vdot(rhs, r)
        p = r.copy() if it == 1 else r + (rho_new / rho) * (alpha / omega) * (p - omega * v)
        ph = vcycle(levels, p)
        v = A.apply(ph)
        alpha = rho_new / np.vdot(rhs, v)
        s = r - alpha * v
        sh = vcycle(levels, s)
        t = A.apply(sh)
        omega = np.vdot(t, s) / np.vdot(t, t)
        x += alpha * ph + omega * sh
        r = s - omega * t
        rho = rho_new
        if np.max(np.abs(r)) <= tol:
            return x, it
    return x, maxit
