"""Small engine-1 run for compute-sanitizer (memcheck / racecheck): the assembly kernels 2 .. 6 on a ragged grid, a few multigrid time steps
with the deterministic cross-CTA sums:  compute-sanitizer --tool racecheck python profiles/run_sanitize.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SY2D_DETERMINISTIC"] = "1"
import numpy as np
import sayram2d_b200 as sy
from sayram2d_b200 import fields

nx, ny = 72, 136
xe, ye = fields.uniform_edges(nx, ny)
eng = sy.Engine(xe, ye, 0.002)
Dxx, Dxy, Dyy, inv_tau = fields.synthetic_tensor(xe, ye)
eng.set_coeffs(fields.ay_G(xe, ye), Dxx, Dxy, Dyy, inv_tau)
f0, bct, lines = fields.ay_init_and_bc(xe, ye)
eng.set_bc(bct, *lines)
eng.set_f(f0)
eng.set_options(engine=1, precond=2, use_graph=0)
print(eng.step(2))
ref = None
for variant in (2, 3, 4, 5, 6):
    o = eng.options(); o.engine = 1; o.reserved[0] = variant
    eng._check(eng.lib.sy2d_set_options(eng._ctx, o)); eng._opt = o
    got = eng.dump_scaled_operator()
    print(variant, eng.last_assembly_kernel(), float(np.sum(got[1])))
    if ref is None:
        ref = got
    elif variant != 3:
        assert all(np.array_equal(a, b) for a, b in zip(got, ref))
eng.close()

# the ensemble kernel (engine 2: work queue, bulk-copy assembly ring, tensor-memory scratchpad) on a few members
import bench  # noqa: E402
bench.NB_TOTAL = 4096
eng2, _ = bench.make_ensemble(0, 6, 0)
print(eng2.step(2))
eng2.close()
