"""Launches the lockstep-engine kernels on an n x n synthetic grid (for ncu captures):
    ncu --set full -k regex:k_assemble_tiled -c 1 python profiles/run_kernels.py 2048
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
eng, _ = bench.make_grid(n, 0)
eng.set_options(engine=1)
for name in ("assembly", "spmv_v", "spmv_t", "xr_update", "p_update", "s_update"):
    print(name, round(1e3 * eng.bench_kernel(name, 5), 2), "us")
eng.close()
