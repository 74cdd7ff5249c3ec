"""Sustained launches of ONE engine-1 assembly kernel on an n x n synthetic grid (for ncu captures):
    ncu --set full --clock-control none --import-source on -k regex:k_assemble_march -s 3 -c 1 -o gpurun_out/x \
        python profiles/run_asm_variant.py 4096 3
variant: 2 tile kernel with plain loads, 3 warp-marching kernel, 4 TMA-staged tiles."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 3
eng, _ = bench.make_grid(n, 0)
o = eng.options(); o.engine = 1; o.reserved[0] = variant
eng._check(eng.lib.sy2d_set_options(eng._ctx, o))
print(n, variant, round(1e3 * eng.bench_kernel("assembly", 6), 2), "us")
eng.close()
