"""Sustained per-kernel times (sy2d_bench_kernel) of the lockstep kernels for several grid caps:
    python profiles/sweep_caps.py 1024 [4096]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

for n in [int(a) for a in sys.argv[1:]] or [1024]:
    for cap in (4, 6, 8, 12, 16, 24, 32):
        os.environ["SY2D_CTAS_PER_SM"] = str(cap)
        os.environ["SY2D_ASM_CTAS_PER_SM"] = str(min(cap, 4))
        eng, _ = bench.make_grid(n, 0)
        eng.set_options(engine=1)
        row = [f"{name}={1e3 * eng.bench_kernel(name, 20):.2f}" for name in ("assembly", "spmv_v", "spmv_t", "xr_update", "p_update", "s_update")]
        print(n, "cap", cap, " ".join(row), "us", flush=True)
        eng.close()
    for acap in (1, 2, 3):
        os.environ["SY2D_CTAS_PER_SM"] = "16"
        os.environ["SY2D_ASM_CTAS_PER_SM"] = str(acap)
        eng, _ = bench.make_grid(n, 0)
        eng.set_options(engine=1)
        print(n, "asm cap", acap, f"assembly={1e3 * eng.bench_kernel('assembly', 20):.2f} us", flush=True)
        eng.close()
