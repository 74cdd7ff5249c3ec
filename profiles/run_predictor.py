"""Column-scale predictor of the ensemble kernel (options.predictor = 0 / 1 / 2): iterations and ms per step, and
the difference of the solutions (they must agree to the solver tolerance).
    python profiles/run_predictor.py [members] [steps] [first member]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

members = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
bench.NB_TOTAL = 4096
stride = 4096 // members
out = {}
for pred in (0, 1, 2):
    eng, _ = bench.make_ensemble(lo, lo + members, 0)
    eng.set_options(predictor=pred)
    eng.step(5)
    ms, its = [], []
    for rep in range(3):
        st = eng.step(steps)
        ms.append(1e3 * st["seconds_device"] / steps)
        its.append(st["iters_sum_all"] / (members * steps))
    out[pred] = eng.get_f()
    print({"predictor": pred, "members": members, "ms_per_step": [round(m, 4) for m in ms], "mean_iters": [round(i, 3) for i in its],
           "negatives": st["negatives"]}, flush=True)
    eng.close()
for pred in (0, 2):
    print("max rel diff predictor", pred, "vs 1:", float(np.max(np.abs(out[pred] - out[1]) / np.abs(out[1]))))
