import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np
import bench
bench.NB_TOTAL = 4096
for nb in (512, 1184, 2048, 4096):
    eng, _ = bench.make_ensemble(0, nb, 0)
    eng.step(5)
    st = eng.step(20)
    print(nb, st["iters_sum_all"] / (nb * 20), st["iters_total"] / 20, 1e3 * st["seconds_device"] / 20, flush=True)
    f = eng.get_f()
    eng.close()
    if nb == 512: f512 = f
    else: print("  max rel diff of first 512 members vs the 512-member run:", float(np.max(np.abs(f[:512] - f512) / np.abs(f512))))
