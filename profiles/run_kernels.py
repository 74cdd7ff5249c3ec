"""Launches the lockstep-engine kernels on an n x n synthetic grid (for ncu captures):
    ncu --set full -k regex:k_assemble_tiled -c 1 python profiles/run_kernels.py 2048
Prints the sustained time per launch of the Jacobi kernels (sy2d_bench_kernel) and then runs a
few x-line iterations of one time step so that the k_xl_* kernels appear in the capture too."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
eng, _ = bench.make_grid(n, 0)
eng.set_options(engine=1)
for name in ("assembly", "spmv_v", "spmv_t", "xr_update", "p_update", "s_update"):
    sys.stdout.write(f"{name} {round(1e3 * eng.bench_kernel(name, 5), 2)} us\n")
eng.set_options(engine=1, use_graph=0, maxit=8, check_every=8)
try:
    eng.step(1)
except Exception as ex:
    sys.stdout.write(f"stopped after 8 x-line iterations: {ex}\n")
eng.close()
