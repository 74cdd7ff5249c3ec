"""Consistency of sy2d_stats across engines: fmin == min f, negatives == count(f < 0), step counter, time."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sayram2d_b200 as sy
from sayram2d_b200 import fields

def run(nx, ny, nbatch, engine, precond, steps=3):
    xe, ye = fields.uniform_edges(nx, ny)
    Dxx, Dxy, Dyy, inv_tau = fields.synthetic_tensor(xe, ye)
    G = fields.ay_G(xe, ye)
    f0, bct, lines = fields.ay_init_and_bc(xe, ye)
    rep = lambda a: np.broadcast_to(a, (nbatch,) + a.shape).copy()
    scale = (1.0 + 0.25 * np.arange(nbatch))[:, None, None]
    eng = sy.Engine(xe, ye, 0.002, nbatch=nbatch)
    eng.set_options(engine=engine, precond=precond)
    eng.set_coeffs(rep(G), rep(Dxx) * scale, rep(Dxy) * scale, rep(Dyy) * scale, rep(inv_tau))
    eng.set_bc(bct, *lines)
    eng.set_f(rep(f0))
    st = eng.step(steps)
    f = eng.get_f()
    ok = st["fmin"] == float(f.min()) and st["negatives"] == int((f < 0).sum()) and eng.step_count() == steps and st["steps"] == steps
    print(nx, ny, nbatch, "engine", st["engine"], "precond", st["precond"], "fmin", st["fmin"], float(f.min()), "neg", st["negatives"], "iters", st["iters_total"], st["iters_last"], st["iters_sum_all"], "resid", st["resid_last"], "OK" if ok else "MISMATCH", flush=True)
    eng.close()
    return ok

allok = True
for cfg in [(80, 80, 5, 2, 1), (80, 80, 5, 2, 0), (64, 48, 3, 2, -1), (30, 20, 4, 2, -1), (80, 80, 2, 1, 0), (256, 256, 1, 1, 2), (200, 120, 2, 1, 1), (96, 61, 3, 1, 0), (100, 70, 3, 1, -1), (9, 1, 2, 0, -1), (1, 9, 2, 0, -1)]:
    try:
        allok &= run(*cfg)
    except Exception as e:
        print(cfg, "ERROR", e); allok = False
print("ALL OK" if allok else "SOME MISMATCH")
