"""NumPy restatement of the ensemble kernel's linear solve (sayram2d_b200/csrc/sy2d_xline_kernel.cuh): BiCGSTAB on the
f-scaled unit-diagonal system, right-preconditioned by the x-line tridiagonal T = tridiag(wW, 1, wE), in the
PIVOT-SCALED form the kernel iterates on (rows divided by the pivots d_i of T's LU, so that both triangular factors
have a unit diagonal), with the solution accumulated as y = sum alpha p + omega s and x = T'^-1 y formed once.
Test infrastructure only (tests/test_xline_reference.py); arrays are [nx][ny], lines run along axis 0."""
import numpy as np


def factor_pivot_scaled(wW, wE):
    """l'_i = wW_i / d_i, e_i = wE_i / d_i, 1/d_i with d_i = 1 - wW_i e_{i-1} (d_0 = 1)."""
    nx = wW.shape[0]
    lp, e, dinv = np.zeros_like(wW), np.zeros_like(wW), np.zeros_like(wW)
    eprev = np.zeros(wW.shape[1])
    for i in range(nx):
        dinv[i] = 1.0 / (1.0 - wW[i] * eprev)
        lp[i] = wW[i] * dinv[i]
        e[i] = wE[i] * dinv[i]
        eprev = e[i]
    return lp, e, dinv


def solve_unit_lu(lp, e, b):
    """(I + L')(I + U') x = b: z_i = b_i - l'_i z_{i-1}, x_i = z_i - e_i x_{i+1}."""
    nx = b.shape[0]
    z = b.copy()
    for i in range(1, nx):
        z[i] = b[i] - lp[i] * z[i - 1]
    x = z.copy()
    for i in range(nx - 2, -1, -1):
        x[i] = z[i] - e[i] * x[i + 1]
    return x


def apply_A(w, x):
    wW, wE, wS, wN = w
    out = x.copy()
    out[1:] += wW[1:] * x[:-1]; out[:-1] += wE[:-1] * x[1:]
    out[:, 1:] += wS[:, 1:] * x[:, :-1]; out[:, :-1] += wN[:, :-1] * x[:, 1:]
    return out


def solve(w, rhs, tol=1e-14, maxit=500):
    """Returns (x, iterations, max|rhs - A x|) of A x = rhs the way the kernel computes it."""
    wW, wE, wS, wN = w
    lp, e, dinv = factor_pivot_scaled(wW, wE)
    wSp, wNp = wS * dinv, wN * dinv
    r = rhs * dinv                       # r0 = rhs' ; rhat = r0
    rhat = r.copy()

    def apply_precond(b):                # b -> (phat, A' phat) with A' phat = b + wS' phat_S + wN' phat_N
        hat = solve_unit_lu(lp, e, b)
        out = b.copy()
        out[:, 1:] += wSp[:, 1:] * hat[:, :-1]; out[:, :-1] += wNp[:, :-1] * hat[:, 1:]
        return out

    rho = np.sum(rhat * r)
    p = v = None
    alpha = omega = 1.0
    beta = 0.0
    y = np.zeros_like(r)
    it = 0
    while np.max(np.abs(r)) > tol and it < maxit:
        p = r.copy() if it == 0 else r + beta * (p - omega * v)
        v = apply_precond(p)
        alpha = rho / np.sum(rhat * v)
        s = r - alpha * v
        t = apply_precond(s)
        omega = np.sum(t * s) / np.sum(t * t)
        y += alpha * p + omega * s
        r = s - omega * t
        rho_new = np.sum(rhat * r)
        beta = (rho_new / rho) * (alpha / omega)
        rho = rho_new
        it += 1
    x = solve_unit_lu(lp, e, y)
    return x, it, float(np.max(np.abs(rhs - apply_A(w, x)))), dinv, lp, e
