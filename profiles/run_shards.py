"""Per-shard cost of the 4096-member ensemble (BASELINE config 4) on ONE GPU: the 8 shards an 8-GPU run gives its ranks,
contiguous (members [512 r, 512 (r+1))) against strided (members r, r + 8, ...).  ms per time step and iterations."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import sayram2d_b200 as sy
from sayram2d_b200 import fields

def run(members):
    g = bench.lc_fields()
    nb = len(members)
    a, b = fields.ensemble_scales(np.asarray(members))
    sc = lambda arr, s: arr[None] * s[:, None, None]
    one = np.ones(nb)
    eng = sy.Engine(g["x_edges"], g["y_edges"], bench.DT, nbatch=nb)
    eng.set_coeffs(sc(g["G"], one), sc(g["Dxx"], a), sc(g["Dxy"], a), sc(g["Dyy"], a), sc(g["inv_tau"], b))
    _, bct, lines = fields.ay_init_and_bc(g["x_edges"], g["y_edges"], lc=True)
    eng.set_bc(bct, *lines)
    eng.set_f(np.ascontiguousarray(sc(g["f_0"], one)))
    eng.step(5)
    st = eng.step(20)
    eng.close()
    return 1e3 * st["seconds_device"] / 20, st["iters_sum_all"] / (20 * nb)

out = {"contiguous": [], "strided": []}
for r in range(8):
    out["contiguous"].append(run(list(range(512 * r, 512 * (r + 1)))))
    out["strided"].append(run(list(range(r, 4096, 8))))
print(json.dumps(out))
