"""Ensemble kernel on one CTA per problem against the CTA-pair (cluster) kernels: ms per step, iterations, agreement.
    python profiles/run_cluster.py [members] [steps] [variants, e.g. 012]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

members = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
variants = [int(c) for c in (sys.argv[3] if len(sys.argv) > 3 else "01")]
bench.NB_TOTAL = 4096
out = {}
for cl in variants:
    os.environ["SY2D_XLINE_CLUSTER"] = str(cl)
    eng, _ = bench.make_ensemble(0, members, 0)
    eng.step(5)
    ms, its = [], []
    for rep in range(3):
        st = eng.step(steps)
        ms.append(1e3 * st["seconds_device"] / steps)
        its.append(st["iters_sum_all"] / (members * steps))
    out[cl] = eng.get_f()
    print({"cluster": cl, "members": members, "ms_per_step": [round(m, 4) for m in ms], "mean_iters": [round(i, 3) for i in its],
           "negatives": st["negatives"], "resid_last": st["resid_last"]}, flush=True)
    eng.close()
for cl in variants[1:]:
    print("max rel diff variant", cl, "vs", variants[0], ":", float(np.max(np.abs(out[cl] - out[variants[0]]) / np.abs(out[variants[0]]))))
