import sys, os
sys.path.insert(0, '/root/repo')
import bench
for n in (1024, 4096):
    eng, _ = bench.make_grid(n, 0)
    eng.set_options(engine=1, precond=2)
    eng.step(2)
    out = {k: round(1e3 * min(eng.bench_kernel(k, 100 if n == 1024 else 30) for _ in range(3)), 2) for k in ("assembly", "p_update", "spmv_v", "s_update", "spmv_t", "xr_update")}
    st = eng.step(5)
    print(n, os.environ.get("SY2D_DETERMINISTIC", "0"), out, round(1e3 * st["seconds_device"] / 5, 3), "ms/step", flush=True)
    eng.close()
