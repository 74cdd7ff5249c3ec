"""Research prototype (NumPy, CPU): y-semi-coarsening multigrid with x-line smoothing as the
preconditioner of the engine's BiCGSTAB on the scaled system A d = rhs.

    python profiles/proto_mg.py [n] [steps]

Counts iterations to max|r| <= 1e-14 for (a) full x-line Jacobi preconditioning (what engine 2
does), (b) V(1,1) cycles.  Uses the oracle only to produce the step's (M, R) - not product code."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import ppfv_oracle as O  # noqa: E402


def thomas_factor(w, d, e):
    """LU of tridiag(w, d, e) along axis 0 for every column."""
    nx = d.shape[0]
    l = np.zeros_like(d)
    dd = np.zeros_like(d)
    dd[0] = d[0]
    for i in range(1, nx):
        l[i] = w[i] / dd[i - 1]
        dd[i] = d[i] - l[i] * e[i - 1]
    return l, 1.0 / dd


def thomas_solve(l, dinv, e, b):
    nx = b.shape[0]
    z = b.copy()
    for i in range(1, nx):
        z[i] -= l[i] * z[i - 1]
    z[nx - 1] *= dinv[nx - 1]
    for i in range(nx - 2, -1, -1):
        z[i] = (z[i] - e[i] * z[i + 1]) * dinv[i]
    return z


SEG = int(os.environ.get("SEG", "16"))


class Level:
    def __init__(self, d, w, e, s, n):
        self.d, self.w, self.e, self.s, self.n = d, w, e, s, n
        self.l, self.dinv = thomas_factor(w, d, e)
        ws, es = w.copy(), e.copy()      # segmented lines: couplings cut every SEG rows
        ws[0::SEG] = 0.0
        es[SEG - 1::SEG] = 0.0
        self.es = es
        self.ls, self.dinvs = thomas_factor(ws, d, es)

    def apply(self, x):
        y = self.d * x
        y[1:] += self.w[1:] * x[:-1]
        y[:-1] += self.e[:-1] * x[1:]
        y[:, 1:] += self.s[:, 1:] * x[:, :-1]
        y[:, :-1] += self.n[:, :-1] * x[:, 1:]
        return y

    def yoff(self, x):
        y = np.zeros_like(x)
        y[:, 1:] += self.s[:, 1:] * x[:, :-1]
        y[:, :-1] += self.n[:, :-1] * x[:, 1:]
        return y

    def segline(self, b):
        return thomas_solve(self.ls, self.dinvs, self.es, b)

    def line(self, b):
        return thomas_solve(self.l, self.dinv, self.e, b)


def coarsen(L, theta=0.5):
    ny = L.d.shape[1]
    assert ny % 2 == 0
    a, b = slice(0, ny, 2), slice(1, ny, 2)
    w = L.w[:, a] + L.w[:, b]
    e = L.e[:, a] + L.e[:, b]
    s_full = L.s[:, a].copy()
    n_full = L.n[:, b].copy()
    d = L.d[:, a] + L.d[:, b] + L.n[:, a] + L.s[:, b]
    # rescale the couplings between aggregates by theta, keep the column sums
    d[:, :-1] += (1 - theta) * s_full[:, 1:]   # entry (row J+1, col J)
    d[:, 1:] += (1 - theta) * n_full[:, :-1]   # entry (row J-1, col J)
    return Level(d, w, e, theta * s_full, theta * n_full)


def restrict(r):
    return r[:, 0::2] + r[:, 1::2]


def prolong(zc):
    return np.repeat(zc, 2, axis=1)


def vcycle(levels, k, r, smoother="zebra", omega=0.8, nu=1):
    L = levels[k]
    ny = r.shape[1]
    if ny == 1 or k == len(levels) - 1:
        if ny == 1:
            return L.line(r)
        z = np.zeros_like(r)
        for _ in range(4):
            z = smooth(L, z, r, smoother, omega)
        return z
    z = np.zeros_like(r)
    for _ in range(nu):
        z = smooth(L, z, r, smoother, omega, first=True)
    res = r - L.apply(z)
    zc = vcycle(levels, k + 1, restrict(res), smoother, omega, nu)
    z = z + prolong(zc)
    for _ in range(nu):
        z = smooth(L, z, r, smoother, omega, reverse=True)
    return z


def smooth(L, z, r, kind, omega, first=False, reverse=False):
    if kind == "segjacobi":
        return z + omega * L.segline(r - L.apply(z))
    if kind == "jacobi":
        return z + omega * L.line(r - L.apply(z))
    # zebra line Gauss-Seidel over columns j (even then odd; reversed on the way up)
    z = z.copy()
    order = (1, 0) if reverse else (0, 1)
    for par in order:
        rhs = r - L.yoff(z)
        zz = L.line(rhs)
        z[:, par::2] = zz[:, par::2]
    return z


def bicgstab(Aop, b, prec, tol=1e-14, maxit=3000):
    x = np.zeros_like(b)
    r = b.copy()
    rhat = b.copy()
    rho = alpha = omega = 1.0
    p = np.zeros_like(b)
    v = np.zeros_like(b)
    for it in range(1, maxit + 1):
        rho_new = np.vdot(rhat, r)
        beta = (rho_new / rho) * (alpha / omega) if it > 1 else 0.0
        p = r + beta * (p - omega * v) if it > 1 else r.copy()
        ph = prec(p)
        v = Aop(ph)
        alpha = rho_new / np.vdot(rhat, v)
        s = r - alpha * v
        sh = prec(s)
        t = Aop(sh)
        omega = np.vdot(t, s) / np.vdot(t, t)
        x += alpha * ph + omega * sh
        r = s - omega * t
        rho = rho_new
        if np.max(np.abs(r)) <= tol:
            return x, it
    return x, maxit


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    p, m, eq = O.build_case("SYN", None, None, nalpha0=n, nE=n, alpha0min=5, alpha0max=90, Emin=0.2, Emax=5, T=1.0, nplots=10, nsteps=500)
    sol = O.Solver(m, eq, linear="splu")
    for step in range(nsteps):
        op = sol.assemble()
        f = sol.f
        c = f.copy()
        # B = M diag(c): the row-unscaled form used on all MG levels; A = diag(1/B_KK) B
        d = op["diag"] * c
        w = np.zeros_like(c); e = np.zeros_like(c); s = np.zeros_like(c); nn = np.zeros_like(c)
        w[1:] = op["W"][1:] * c[:-1]
        e[:-1] = op["E"][:-1] * c[1:]
        s[:, 1:] = op["S"][:, 1:] * c[:, :-1]
        nn[:, :-1] = op["N"][:, :-1] * c[:, 1:]
        fine = Level(d, w, e, s, nn)
        om = d
        A = Level(np.ones_like(d), w / om, e / om, s / om, nn / om)
        rhs = op["R"] / om - A.apply(np.ones_like(c))
        print(f"step {step}: n={n} max|rhs|={np.abs(rhs).max():.3e}")
        t0 = time.time()
        if n <= 512:
            _, it = bicgstab(A.apply, rhs, A.line)
            print(f"  full x-line Jacobi: {it} iterations ({time.time() - t0:.1f} s)")
        for theta in (0.5,):
            levels = [fine]
            while levels[-1].d.shape[1] > 1 and levels[-1].d.shape[1] % 2 == 0:
                levels.append(coarsen(levels[-1], theta))
            for kind, omg, nu in (("segjacobi", 0.7, 1),):
                for depth in (5,):
                    lv = levels[:depth]
                    t0 = time.time()
                    if os.environ.get("UNW"):
                        lvA = [A]
                        while len(lvA) < depth:
                            lvA.append(coarsen(lvA[-1], theta))
                        xs, it = bicgstab(A.apply, rhs, lambda q: vcycle(lvA, 0, q, kind, omg, nu), maxit=400)
                    else:
                        xs, it = bicgstab(A.apply, rhs, lambda q: vcycle(lv, 0, om * q, kind, omg, nu), maxit=400)
                    true = np.abs(rhs - A.apply(xs)).max()
                    print(f"  MG theta={theta} {kind} omega={omg} nu={nu} levels={len(lv)}: {it} iterations, true resid {true:.2e} ({time.time() - t0:.1f} s)")
        if step + 1 < nsteps:
            sol.update()


if __name__ == "__main__":
    main()
