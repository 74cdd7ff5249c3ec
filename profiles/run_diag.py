"""Scratch diagnostics run on the GPU box (numbers quoted in profiles/r02_*.md)."""
import json
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import sayram2d_b200 as sy
from sayram2d_b200 import fields
import bench

what = sys.argv[1] if len(sys.argv) > 1 else "all"

if what in ("random", "all"):
    import test_gpu_parity as T
    for (seed, nx, ny, bc) in [(1, 33, 21, (0, 0, 0, 0)), (2, 7, 50, (1, 1, 1, 1)), (3, 64, 64, (0, 1, 1, 0)), (6, 2, 2, (0, 1, 0, 1)), (7, 45, 36, (0, 0, 0, 0))]:
        for precond in (0, 1, 2):
            m, eq, lines, f = T._random_case(seed, nx, ny, bc)
            eng = sy.Engine(m.x_edges, m.y_edges, m.dt)
            try:
                eng.set_options(engine=1, precond=precond)
                eng.set_coeffs(eq.G, eq.Dxx, eq.Dxy, eq.Dyy, eq.inv_tau)
                eng.set_bc(bc, *[l if b == 0 else None for l, b in zip(lines, bc)])
                eng.set_f(f)
                st = eng.step(5)
                print("random", seed, nx, ny, precond, "ok", st["iters_total"], st["resid_last"])
            except Exception as ex:
                print("random", seed, nx, ny, precond, "ERR", str(ex)[:200], getattr(eng, "last_stats", None))
            eng.close()

if what in ("asm", "all"):
    for n in (1024, 2048, 4096):
        eng, f0 = bench.make_grid(n, 0)
        eng.set_options(engine=1)
        row = {}
        for name, variant in (("tiled", 2), ("march", 3), ("tma", 4)):
            o = eng.options(); o.engine = 1; o.reserved[0] = variant
            eng._check(eng.lib.sy2d_set_options(eng._ctx, o))
            ms = eng.bench_kernel("assembly", 20)
            row[name] = round(1e3 * ms, 2)
        print("asm_sustained_us", n, row, "frac104", {k: round(n * n * 104 / (v * 1e-6) / 1e9 / 6539.9, 3) for k, v in row.items()})
        eng.close()

if what in ("pipe", "all"):
    import torch
    for members in (512, 4096):
        for chunks in ((3, 6, 12, 24) if members == 512 else (16, 27, 32, 48)):
            os.environ["SY2D_PIPE_CHUNKS"] = str(chunks)
            bench.NB_TOTAL = members
            eng, f0 = bench.make_ensemble(0, members, 0)
            pin_in = torch.empty((members, 80, 80), dtype=torch.float64).pin_memory()
            pin_out = torch.empty((members, 80, 80), dtype=torch.float64).pin_memory()
            pin_in.numpy()[...] = f0
            h_in, h_out = pin_in.numpy(), pin_out.numpy()
            eng.step(3)
            for _ in range(3):
                eng.step_host(h_in, h_out, 1); h_in, h_out = h_out, h_in
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(20):
                eng.step_host(h_in, h_out, 1); h_in, h_out = h_out, h_in
            torch.cuda.synchronize()
            e2e = (time.perf_counter() - t0) / 20
            st = eng.step(20)
            print("pipe", members, chunks, "e2e_ms", round(1e3 * e2e, 3), "device_ms", round(1e3 * st["seconds_device"] / 20, 3))
            eng.close()
    del os.environ["SY2D_PIPE_CHUNKS"]
