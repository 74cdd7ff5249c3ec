"""Sustained launches of the engine-1 assembly kernels (variant 4 = TMA tiles in strided order, 5 = column runs, 6 = two cells per thread on 8 x 64 tiles) on n x n
synthetic grids, and the column-run kernel with several boundary-tile weights:  python profiles/run_asm_cmp.py [reps]"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if len(sys.argv) > 2 and sys.argv[1] == "one":
    import bench  # noqa: E402
    n, variant = int(sys.argv[2]), int(sys.argv[3])
    eng, _ = bench.make_grid(n, 0)
    o = eng.options(); o.engine = 1; o.reserved[0] = variant
    eng._check(eng.lib.sy2d_set_options(eng._ctx, o))
    eng.bench_kernel("assembly", 20)
    t = min(eng.bench_kernel("assembly", 200 if n <= 2048 else 50) for _ in range(3))
    print(n, variant, os.environ.get("SY2D_COL_EDGE_WEIGHT", "-"), os.environ.get("SY2D_WIDE_CTAS_PER_SM", "-"), round(1e3 * t, 2), "us", flush=True)
    eng.close()
else:
    for n in (1024, 2048, 4096):
        for variant, env in ((4, {}), (5, {}), (6, {}), (6, {"SY2D_WIDE_CTAS_PER_SM": "1"})):
            subprocess.run([sys.executable, __file__, "one", str(n), str(variant)], env={**os.environ, **env})
