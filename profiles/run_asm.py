"""Two steps of a small synthetic grid on engine 1 (assembly variant in argv[3]: 0 TMA, 1 per cell, 2 tiled)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sayram2d_b200 as sy
from sayram2d_b200 import fields

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ny = int(sys.argv[2]) if len(sys.argv) > 2 else 64
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 0
xe, ye = fields.uniform_edges(nx, ny)
eng = sy.Engine(xe, ye, 0.002)
o = eng.options(); o.engine = 1; o.reserved[0] = variant
eng._check(eng.lib.sy2d_set_options(eng._ctx, o)); eng._opt = o
Dxx, Dxy, Dyy, inv_tau = fields.synthetic_tensor(xe, ye)
eng.set_coeffs(fields.ay_G(xe, ye), Dxx, Dxy, Dyy, inv_tau)
f0, bct, lines = fields.ay_init_and_bc(xe, ye)
eng.set_bc(bct, *lines)
eng.set_f(f0)
print(eng.step(2))
print("f checksum", float(np.sum(eng.get_f())))
eng.close()
