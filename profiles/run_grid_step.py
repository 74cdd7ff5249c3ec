"""One time step of the n x n synthetic grid with the lockstep engine (for ncu captures)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
eng, _ = bench.make_grid(n, 0)
eng.set_options(engine=1, use_graph=0, maxit=int(sys.argv[2]) if len(sys.argv) > 2 else 20000, check_every=8)
if len(sys.argv) > 3:
    eng.set_options(precond=int(sys.argv[3]))
if os.environ.get("SY2D_EVENTS"):
    eng.set_profiling(True)
try:
    st = eng.step(1)
    print(st)
except Exception as ex:  # a small maxit is used to bound the profiled run
    print("stopped:", ex)
if os.environ.get("SY2D_EVENTS"):
    for k, v in eng.profile().items():
        if v["launches"]:
            print(k, v["launches"], "launches", round(1e3 * v["ms"] / v["launches"], 1), "us/launch")
eng.close()
