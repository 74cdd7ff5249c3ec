"""Mean BiCGSTAB iterations per step over time steps 6-25 and 26-45 for predictor 1 and 2, by blocks of 512 ensemble members."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

bench.NB_TOTAL = 4096
for lo in range(0, 4096, 512):
    row = {}
    for pred in (1, 2):
        eng, _ = bench.make_ensemble(lo, lo + 512, 0)
        eng.set_options(predictor=pred)
        eng.step(5)
        a = eng.step(20)["iters_sum_all"] / (512 * 20)
        b = eng.step(20)["iters_sum_all"] / (512 * 20)
        row[pred] = (round(a, 2), round(b, 2))
        eng.close()
    print(lo, row, flush=True)
