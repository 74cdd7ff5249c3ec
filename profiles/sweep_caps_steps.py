"""Whole time steps (multigrid-preconditioned lockstep engine) for several grid caps of the grid-stride kernels:
    python profiles/sweep_caps_steps.py 512 1024 2048"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

for n in [int(a) for a in sys.argv[1:]] or [1024]:
    for cap in (2, 3, 4, 6, 8, 16):
        os.environ["SY2D_CTAS_PER_SM"] = str(cap)
        eng, _ = bench.make_grid(n, 0)
        eng.set_options(engine=1)
        eng.step(2)
        st = eng.step(4)
        row = [f"{name}={1e3 * eng.bench_kernel(name, 20):.2f}" for name in ("spmv_v", "spmv_t", "xr_update")]
        print(n, "cap", cap, f"{1e3 * st['seconds_device'] / 4:.3f} ms/step {st['iters_total'] / 4:.1f} it/step", " ".join(row), flush=True)
        eng.close()
