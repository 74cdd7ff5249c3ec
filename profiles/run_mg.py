"""Time steps of the n x n synthetic grid with the multigrid-preconditioned lockstep engine:
    python profiles/run_mg.py [n] [steps] [precond] [mg_levels] [mg_coarse_sweeps]
prints ms/step, iterations, and (SY2D_EVENTS=1) the per-class CUDA-event times of one more step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
precond = int(sys.argv[3]) if len(sys.argv) > 3 else 2
levels = int(sys.argv[4]) if len(sys.argv) > 4 else 0
eng, _ = bench.make_grid(n, 0)
sweeps = int(sys.argv[5]) if len(sys.argv) > 5 else 0
eng.set_options(engine=1, precond=precond, mg_levels=levels, mg_coarse_sweeps=sweeps, use_graph=0 if os.environ.get("SY2D_NOGRAPH") else 1)
eng.step(2)
st = eng.step(steps)
print(f"n={n} precond={st['precond']} levels={levels} sweeps={sweeps}: {1e3 * st['seconds_device'] / steps:.3f} ms/step, "
      f"{st['iters_total'] / steps:.1f} iterations/step, {n * n * steps / st['seconds_device'] / 1e6:.1f} M cell-updates/s, "
      f"launches {st['kernel_launches']}, resid {st['resid_last']:.2e}, negatives {st['negatives']}")
if os.environ.get("SY2D_EVENTS"):
    eng.set_profiling(True)
    eng.step(1)
    for k, v in eng.profile().items():
        if v["launches"]:
            print(f"  {k:14s} {v['launches']:5d} launches {1e3 * v['ms'] / v['launches']:8.1f} us/launch {v['ms']:8.3f} ms")
eng.close()
