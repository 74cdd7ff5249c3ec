"""Iterations per time step of the n x n grid (multigrid-preconditioned engine 1) for predictor 1 and 2, step by step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
for pred in (1, 2):
    eng, _ = bench.make_grid(n, 0)
    eng.set_options(engine=1, precond=2, predictor=pred)
    its, ms = [], []
    for s in range(steps):
        st = eng.step(1)
        its.append(st["iters_total"]); ms.append(round(1e3 * st["seconds_device"], 2))
    print({"n": n, "predictor": pred, "iters": its, "ms": ms}, flush=True)
    eng.close()
