#!/usr/bin/env python
"""Benchmark of the Sayram-2D time-step hot path on B200 (see DESIGN.md "measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): cell-updates/sec = cells advanced one full implicit time
step (PPFV assembly + complete linear solve) per second, whole job.

Workload of the headline line ("ensemble4096"): BASELINE config 4 - 4096
independent (L, MLT) problems on the 80x80 grid of p_AlbertYoungLC.ini (member m:
D scaled by a_m, 1/tau by b_m), sharded contiguously over the N ranks with no
data-path collective ("scaling": "strong" - the total is fixed at 4096 members).
A "step" is one time step of every member.  At N=1 the line also carries
"grid1024": BASELINE config 3 (single 1024x1024 grid, synthetic full tensor +
loss), the configuration the HBM-roofline target is quoted on.

--impl reference times the reference's own CPU implementation (oracle/_ref: the
unmodified Solver.cc etc. compiled against shim headers) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NB_TOTAL = 4096
NX = NY = 80
DT = 0.002
METRIC = "cell_updates_per_sec"
UNIT = "cell-updates/s"
REF_DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
LC_INI = os.path.join(ROOT, "data", "p_AlbertYoungLC.ini")
REF_SUBSTEPS = 40   # --impl reference: one "step" of the CPU arm = this many time steps of its sample members

# algorithmic HBM bytes per cell and launch (DESIGN.md "kernels"; fp64 = 8 B)
BYTES_PER_CELL = {"assembly": 104, "p_update": 32, "spmv_v": 56, "s_update": 24, "spmv_t": 48, "xr_update": 56,
                  "finish": 40, "other": 56, "problem_steps": None}
# lockstep engine with the multigrid preconditioner (sy2d_mg.cuh).  mg_* classes are per LEVEL cell (the
# library counts nx * ny_level cells for a launch on a level): line solves read rhs + 3 LU factors and
# write z (40 B; +z, +zc/2 on the way up: 52 B) - 46 B on average; residual kernels read r, z, 4 weights,
# om or zc and write the restricted residual / t: 66 B; setup per FINE cell: coarsening 60 B per level cell
# and line LU 40 B per level cell, summed over the levels (factor ~1.94): ~190 B.
BYTES_PER_CELL_MG = {"assembly": 112, "spmv_t": 56, "xr_update": 72, "mg_line": 46, "mg_resid": 66, "mg_setup": 190}
PRECOND_NAMES = {0: "jacobi", 1: "xline16", 2: "multigrid"}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def shard(n, rank, world):
    """Contiguous member range of this rank (sayram2d_b200.shard is the tested copy)."""
    from sayram2d_b200.shard import shard_range
    return shard_range(n, rank, world)


def bind_to_gpu_numa_node(torch, local_rank):
    """Run this rank (and first-touch its pinned host buffers) on the CPUs of the NUMA node its GPU hangs off: with
    one rank per GPU the end-to-end loop moves 2 x 26-210 MB per step and rank over PCIe, and buffers on the far
    socket make that cross the inter-socket link.  Best effort: returns the node or None."""
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        dev = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{dev}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        self.path = tempfile.mktemp(prefix="sy2d_clocks_", suffix=".csv")
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(index)], stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1])); smax.append(float(p[2]))
                except ValueError:
                    continue
                for k, nm in enumerate(names):
                    if p[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            busy = [s for s in sm if s > 0.5 * max(sm)] or sm
            out.update(sm_mhz=statistics.median(busy), sm_max_mhz=max(smax), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


# ----------------------------------------------------------------------------- workloads
def lc_fields():
    g = np.load(os.path.join(ROOT, "tests", "golden", "lc80.npz"))
    return {k: g[k] for k in ("x_edges", "y_edges", "G", "Dxx", "Dxy", "Dyy", "inv_tau", "f_0")}


def make_ensemble(lo, hi, device):
    """BASELINE config 4 members [lo, hi): LC case fields as the reference computes them
    (tests/golden/lc80.npz inputs), D scaled by a_m and 1/tau by b_m."""
    import sayram2d_b200 as sy
    from sayram2d_b200 import fields
    g = lc_fields()
    nb = hi - lo
    a, b = fields.ensemble_scales(np.arange(lo, hi))
    sc = lambda arr, s: arr[None] * s[:, None, None]
    one = np.ones(nb)
    eng = sy.Engine(g["x_edges"], g["y_edges"], DT, nbatch=nb, device=device)
    eng.set_coeffs(sc(g["G"], one), sc(g["Dxx"], a), sc(g["Dxy"], a), sc(g["Dyy"], a), sc(g["inv_tau"], b))
    _, bct, lines = fields.ay_init_and_bc(g["x_edges"], g["y_edges"], lc=True)
    eng.set_bc(bct, *lines)
    f0 = np.ascontiguousarray(sc(g["f_0"], one))
    eng.set_f(f0)
    return eng, f0


def make_grid(n, device):
    """BASELINE config 3: AY domain/BCs/G/f0 on an n x n grid with the synthetic tensor."""
    import sayram2d_b200 as sy
    from sayram2d_b200 import fields
    xe, ye = fields.uniform_edges(n, n)
    eng = sy.Engine(xe, ye, DT, nbatch=1, device=device)
    Dxx, Dxy, Dyy, inv_tau = fields.synthetic_tensor(xe, ye)
    eng.set_coeffs(fields.ay_G(xe, ye), Dxx, Dxy, Dyy, inv_tau)
    f0, bct, lines = fields.ay_init_and_bc(xe, ye)
    eng.set_bc(bct, *lines)
    eng.set_f(f0)
    return eng, f0


# lockstep engine with the segmented x-line preconditioner: the five kernel classes move these bytes
BYTES_PER_CELL_XLINE = {"p_update": 64, "spmv_v": 48, "s_update": 80, "spmv_t": 40, "xr_update": 56}


# The engine-2 x-line kernel (k_problem_xline) keeps a problem's state on chip: what it moves goes through the L1TEX data
# pipe of its SM (shared memory), not through HBM.  Algorithmic bytes, counted from the kernel source (DESIGN.md section 4,
# "engine 2"): per cell and BiCGSTAB iteration 11 shared-memory accesses of 8 bytes (two publishes of the line solutions, four
# neighbour reads, p written twice and read three times) + two 2-byte reads of the 16-bit shadow residual = 92 B; the other
# 14 accesses of an iteration (wS' and wN' twice, y read and written, the sweep factors l' and e twice per line solve) go to
# TENSOR MEMORY (tcgen05.ld / st, sy2d_tmem.cuh) and do not pass through the L1TEX data pipe, so they are not counted; per cell
# and time step (block assembly from the staged inputs, factorisation, pivot scaling, factors into tensor memory, final solve
# and update) 78 accesses = 624 B.
XLINE_L1_BYTES_PER_CELL_ITER = 92
XLINE_L1_BYTES_PER_CELL_STEP = 624
HBM_COMPULSORY_BYTES_PER_CELL_STEP = 88   # tx, ty, cxy, U, Ud, f, yprev, ylast in; f, yprev, ylast out


def roofline_from_profile(prof, only=None, stats=None, cells_per_problem=None, peaks=None):
    """achieved = algorithmic bytes / CUDA-event time per kernel, against the measured peak of the unit that bounds it:
    HBM for the streaming kernels of the lockstep engine, the L1TEX data pipe (shared memory) for the engine-2 kernel."""
    peak, peak_src = measured_peak()
    rows = {}
    for name, p in prof.items():
        if p["launches"] == 0 or p["ms"] <= 0:
            continue
        if name == "problem_steps":
            xline = stats.get("precond", 0) == 1
            cell_iters = cells_per_problem * stats["iters_sum_all"]
            sec = p["ms"] * 1e-3
            hbm = HBM_COMPULSORY_BYTES_PER_CELL_STEP * p["cells"] / sec / 1e9
            row = {"ms_total": round(p["ms"], 3), "launches": p["launches"], "precond": "xline" if xline else "jacobi",
                   "mean_iters_per_step": cell_iters / p["cells"], "cell_steps": p["cells"], "cell_iterations": cell_iters,
                   "hbm_compulsory": {"bytes_per_cell_step": HBM_COMPULSORY_BYTES_PER_CELL_STEP, "achieved": round(hbm, 1),
                                      "frac": round(hbm / peak, 4), "peak": peak, "unit": "GB/s"}}
            if xline:
                nbytes = XLINE_L1_BYTES_PER_CELL_STEP * p["cells"] + XLINE_L1_BYTES_PER_CELL_ITER * cell_iters
                gbs = nbytes / sec / 1e9
                l1peak = (peaks or {}).get("smem_gbs")
                row.update({"bound": "l1tex", "achieved": round(gbs, 1), "peak": round(l1peak, 1) if l1peak else None,
                            "frac": round(gbs / l1peak, 4) if l1peak else None, "unit": "GB/s", "bytes_total": nbytes,
                            "bytes_per_cell_iteration": XLINE_L1_BYTES_PER_CELL_ITER, "bytes_per_cell_step": XLINE_L1_BYTES_PER_CELL_STEP})
            rows[name] = row
            continue
        bpc = BYTES_PER_CELL.get(name)
        if stats and stats.get("precond", 0) == 1 and stats.get("engine", 0) == 1:
            bpc = BYTES_PER_CELL_XLINE.get(name, bpc)
        if stats and stats.get("precond", 0) == 2 and stats.get("engine", 0) == 1:
            bpc = BYTES_PER_CELL_MG.get(name, bpc)
        if bpc is None:
            continue
        gbs = p["cells"] * bpc / (p["ms"] * 1e-3) / 1e9
        rows[name] = {"achieved": round(gbs, 1), "frac": round(gbs / peak, 4), "ms_total": round(p["ms"], 3),
                      "launches": p["launches"], "us_per_launch": round(1e3 * p["ms"] / p["launches"], 2),
                      "bytes_per_cell": bpc}
    cand = {k: v for k, v in rows.items() if only is None or k in only or k == "problem_steps"}
    dom = max(cand, key=lambda k: cand[k]["ms_total"]) if cand else None
    return rows, dom, peak, peak_src


def sy_measure_peaks(device):
    import sayram2d_b200 as sy
    try:
        return sy.measure_peaks(device)
    except Exception as ex:  # noqa: BLE001
        return {"error": str(ex)[:200]}


def profile_pass(eng, steps):
    eng.set_profiling(True)
    st = eng.step(steps)
    prof = eng.profile()
    eng.set_profiling(False)
    return prof, st


# ----------------------------------------------------------------------------- CPU reference
def run_ref_driver(case, ini, steps, skip, extra=(), timeout=3600):
    with tempfile.TemporaryDirectory() as work:
        os.symlink(os.path.join(ROOT, "data", "D"), os.path.join(work, "D"))
        cmd = [REF_DRIVER, "--case", case, "--ini", ini, "--out", os.path.join(work, "o"), "--steps", str(steps),
               "--skip", str(skip), "--every", "1000000", *map(str, extra)]
        res = subprocess.run(cmd, cwd=work, capture_output=True, text=True, timeout=timeout, check=True)
        return json.loads(res.stdout.strip().splitlines()[-1])


def spawn_ref_driver(case, ini, steps, skip, extra, work):
    cmd = [REF_DRIVER, "--case", case, "--ini", ini, "--out", os.path.join(work, "o"), "--steps", str(steps),
           "--skip", str(skip), "--every", "1000000", *map(str, extra)]
    return subprocess.Popen(cmd, cwd=work, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)


def cpu_baseline_ensemble(members=(0, 2047, 4095), steps=400):
    """Bounded sample of the same workload: `members` of the ensemble, `steps` time steps
    each, one after the other on ONE core (the reference is single-threaded: no `#pragma omp`
    in its sources and Eigen's SparseLU is sequential - BASELINE.md section 2)."""
    from sayram2d_b200 import fields
    if not os.path.exists(REF_DRIVER):
        return {"value": None, "unit": UNIT, "cores": 1, "kind": "reference+shim-LU", "sample": "oracle/_ref/ref_driver missing"}
    wall, cells = 0.0, 0
    for m in members:
        a, b = fields.ensemble_scales(m)
        r = run_ref_driver("ENS", LC_INI, steps, 0, ("--member", repr(float(a)), repr(float(b))))
        wall += r["loop_wall_s"]
        cells += r["nx"] * r["ny"] * r["timed_steps"]
    return {"value": cells / wall, "unit": UNIT, "cores": 1, "kind": "reference+shim-LU",
            "sample": f"members {list(members)} x {steps} steps of the 4096-member ensemble, sequentially on 1 core "
                      f"({wall:.1f} s of oracle/_ref/ref_driver = reference Solver.cc + shim LU)"}


def cpu_baseline_grid(n, steps=1):
    if not os.path.exists(REF_DRIVER):
        return {"value": None, "unit": UNIT, "cores": 1, "kind": "reference+shim-LU", "sample": "oracle/_ref/ref_driver missing"}
    with tempfile.TemporaryDirectory() as d:
        ini = os.path.join(d, "syn.ini")
        open(ini, "w").write(f"[basic]\nrun_id = syn{n}\nnalpha0 = {n}\nnE = {n}\nalpha0min = 5\nalpha0max = 90\nEmin = 0.2\n"
                             f"Emax = 5\nT = 1.0\nnsteps = 500\n[diagnostics]\nnplots = 10\n[diffusion_coefficients]\n"
                             f"dID = AlbertYoung_chorus\n")
        r = run_ref_driver("SYN", ini, steps, 0)
    return {"value": r["nx"] * r["ny"] * r["timed_steps"] / r["loop_wall_s"], "unit": UNIT, "cores": 1, "kind": "reference+shim-LU",
            "sample": f"{steps} time step(s) of the {n}x{n} grid ({r['loop_wall_s']:.1f} s; LU factor {r['lu_factor_s']:.1f} s, "
                      f"nnz(L+U)={r['nnz_LU']})"}


def reference_arm(args):
    """The reference's CPU implementation on the same workload/metric, using every host core
    it can: the path itself is single-threaded, but ensemble members are independent, so one
    process per core each advances its own member.  A "step" = one time step of that sample."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "gpu_launches": 0,
            "config": {"workload": f"ensemble{NB_TOTAL}", "members": NB_TOTAL, "grid": [NX, NY], "dt": DT}}
    if not os.path.exists(REF_DRIVER):
        line["unavailable"] = "oracle/_ref/ref_driver not built (make -C oracle ref needs /root/reference)"
        print(json.dumps(line), flush=True)
        return
    from sayram2d_b200 import fields
    cores = os.cpu_count() or 1
    procs = min(cores, 64)
    members = [int(round(k * (NB_TOTAL - 1) / max(procs - 1, 1))) for k in range(procs)]
    t0 = time.perf_counter()
    with tempfile.TemporaryDirectory() as base:
        ps = []
        for k, m in enumerate(members):
            work = os.path.join(base, f"w{k}")
            os.makedirs(work)
            os.symlink(os.path.join(ROOT, "data", "D"), os.path.join(work, "D"))
            a, b = fields.ensemble_scales(m)
            ps.append(spawn_ref_driver("ENS", LC_INI, (args.warmup + args.steps) * REF_SUBSTEPS, args.warmup * REF_SUBSTEPS,
                                       ("--member", repr(float(a)), repr(float(b))), work))
        outs = [json.loads(p.communicate()[0].strip().splitlines()[-1]) for p in ps]
    wall = max(o["loop_wall_s"] for o in outs)
    value = procs * NX * NY * args.steps * REF_SUBSTEPS / wall
    line["config"]["sample_per_step"] = f"{REF_SUBSTEPS} time steps of {procs} members (one per host core)"
    line.update(value=value, ms_per_step=1e3 * wall / args.steps,
                cpu_baseline={"value": value, "unit": UNIT, "cores": procs, "kind": "reference+shim-LU",
                              "sample": f"{procs} of the 4096 members (one oracle/_ref/ref_driver process per core, "
                                        f"({args.warmup}+{args.steps}) x {REF_SUBSTEPS} time steps each; the direct LU costs the same "
                                        f"every step); total wall {time.perf_counter() - t0:.1f} s"},
                e2e={"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- ours
def ours(args):
    import torch
    import torch.distributed as dist
    rank, local_rank, world = dist_env()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo", rank=rank, world_size=world)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_node = bind_to_gpu_numa_node(torch, local_rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    lo, hi = shard(NB_TOTAL, rank, world)
    eng, f0 = make_ensemble(lo, hi, local_rank)
    nb = hi - lo
    cells_total = NB_TOTAL * NX * NY

    # ---- device-resident timing (value) ----
    eng.step(args.warmup)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    t0 = time.perf_counter()
    st = eng.step(args.steps)
    barrier()
    wall = time.perf_counter() - t0
    dev_s = max_over_ranks(st["seconds_device"])
    wall_s = max_over_ranks(wall)
    iters = max_over_ranks(st["iters_total"]) / args.steps
    iters_mean = sum_over_ranks(st["iters_sum_all"]) / (NB_TOTAL * args.steps)
    launches = sum_over_ranks(st["kernel_launches"])
    negatives = sum_over_ranks(st["negatives"])
    value = cells_total * args.steps / dev_s

    # ---- end-to-end through the public call with HOST buffers ----
    pin_in = torch.empty((nb, NX, NY), dtype=torch.float64).pin_memory()
    pin_out = torch.empty((nb, NX, NY), dtype=torch.float64).pin_memory()
    pin_in.numpy()[...] = eng.get_f()
    h_in, h_out = pin_in.numpy(), pin_out.numpy()
    for _ in range(max(1, min(args.warmup, 3))):
        eng.step_host(h_in, h_out, 1); h_in, h_out = h_out, h_in
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        # Solver::update() with f resident on the host: sy2d_step_host uploads this step's f from pinned
        # memory, advances it and downloads the result (sub-batches pipelined over streams); returns
        # after the last D2H copy has landed
        eng.step_host(h_in, h_out, 1)
        h_in, h_out = h_out, h_in
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop() if sampler else None
    e2e = {"value": cells_total * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": cells_total * 8,
           "d2h_bytes_per_step": cells_total * 8, "ms_per_step": 1e3 * e2e_s / args.steps,
           "host_buffers": "pinned" + (f", rank bound to the GPU's NUMA node ({numa_node})" if numa_node is not None else "")}
    # The host-side ceiling of that loop: the same bytes per step and rank, pinned host <-> device in both directions at
    # once on two streams, all ranks concurrently, NO kernels.  An end-to-end step cannot be faster than this; on a box
    # whose ranks share one host memory system / PCIe root it, not the GPU, is what the N-GPU end-to-end number runs into.
    d_buf = torch.empty((nb, NX, NY), dtype=torch.float64, device=dev)
    d_buf2 = torch.empty((nb, NX, NY), dtype=torch.float64, device=dev)
    s_up, s_dn = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    def copy_step():
        with torch.cuda.stream(s_up):
            d_buf.copy_(pin_in, non_blocking=True)
        with torch.cuda.stream(s_dn):
            pin_out.copy_(d_buf2, non_blocking=True)
    for _ in range(3):
        copy_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        copy_step()
    barrier()
    copy_s = max_over_ranks(time.perf_counter() - t0)
    e2e["copy_ceiling"] = {"ms_per_step": 1e3 * copy_s / args.steps, "value": cells_total * args.steps / copy_s, "unit": UNIT,
                           "aggregate_GBps_each_way": cells_total * 8 * args.steps / copy_s / 1e9,
                           "what": "pinned H2D + D2H of one step's f per rank, both directions concurrently, all ranks at once, no kernels"}
    del d_buf, d_buf2

    # ---- per-kernel roofline (separate profiled pass, CUDA events around every launch) ----
    peaks = sy_measure_peaks(local_rank) if rank == 0 else None
    prof, pst = profile_pass(eng, 2)
    rows, dom, peak, peak_src = roofline_from_profile(prof, only=("p_update", "spmv_v", "s_update", "spmv_t", "xr_update", "assembly"),
                                                      stats=pst, cells_per_problem=NX * NY, peaks=peaks)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"ensemble{NB_TOTAL}", "members": NB_TOTAL, "members_per_gpu": nb, "grid": [NX, NY], "dt": DT,
                       "case": "Albert_Young_LC fields, D x a_m, 1/tau x b_m (SURVEY 8d config 4)", "parallelism": f"ensemble-shard x{world}",
                       "solver": "BiCGSTAB (x-line preconditioned in engine 2) on the f-scaled unit-diagonal system, max|r|<=1e-14",
                       "l2": f"working set {nb * NX * NY * 8 * 19 / 1e6:.0f} MB per GPU > 126 MB L2: no flush needed"},
            "iters_per_step": iters, "iters_per_step_mean": iters_mean, "engine": st["engine"], "negatives": int(negatives), "wall_ms_per_step": 1e3 * wall_s / args.steps,
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
    if dom and rank == 0:
        r = rows[dom]
        if dom == "problem_steps" and r.get("bound") == "l1tex":
            # the dominant kernel is bound by the L1TEX data pipe (ncu: l1tex throughput is the top unit, DRAM ~4 %):
            # achieved = algorithmic shared-memory + L2-backed bytes / kernel time, peak = the measured shared-memory
            # copy bandwidth of this device (sy2d_measure_peaks); what crosses HBM is listed beside it
            traffic, traffic_src = None, None
            try:  # L1TEX bytes per launch from the committed ncu --set full capture, scaled to this launch's work
                tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))["k_problem_xline"]
                traffic = tj["l1tex_bytes_per_cell_iteration"] * r["cell_iterations"] / r["launches"]
                traffic_src = tj["source"]
                unit_busy = tj.get("ncu_l1tex_throughput_pct")
            except Exception:
                unit_busy = None
            line["roofline"] = {"bound": "l1tex", "kernel": "k_problem_xline", "achieved": r["achieved"], "peak": r["peak"],
                                "unit": "GB/s", "frac": r["frac"], "traffic": traffic, "traffic_source": traffic_src,
                                "ncu_unit_busy_pct": unit_busy,   # l1tex__throughput of the committed capture: what the unit itself reports (wavefront granularity: a 2-byte read costs a wavefront like an 8-byte one)
                                "bytes_per_launch": r["bytes_total"] / r["launches"], "ms_per_launch": r["ms_total"] / r["launches"],
                                "model": f"{XLINE_L1_BYTES_PER_CELL_STEP} B per cell-step + {XLINE_L1_BYTES_PER_CELL_ITER} B per cell-iteration through "
                                         "the L1TEX data pipe (11 shared-memory 8-byte accesses and two 2-byte reads per cell and iteration; 14 more 8-byte accesses go to tensor memory, not counted)",
                                "peak_source": "measured: sy2d_measure_peaks shared-memory copy (8-byte accesses, loads + stores, all SMs)",
                                "hbm_compulsory": r["hbm_compulsory"], "measured_peaks": peaks, "kernels": {"ensemble_xline": r}}
        else:
            line["roofline"] = {"bound": "hbm", "kernel": dom, "achieved": r.get("achieved"), "peak": peak, "unit": "GB/s",
                                "frac": r.get("frac"), "traffic": None, "peak_source": peak_src, "measured_peaks": peaks, "kernels": rows}
    eng.close()

    if rank == 0 and world == 1:
        if not args.no_grid1024:
            line["grid1024"] = grid_object(1024, local_rank, args, torch)
            line["grid4096"] = big_grid_object(4096, local_rank)
            if "roofline" in line:   # the north-star target's kernels, next to the dominant kernel of the headline workload
                line["roofline"]["kernels"].update(north_star_rows(line["grid1024"], line["grid4096"], peak))
        line["cpu_baseline"] = cpu_baseline_ensemble() if not args.no_cpu else None
    if world > 1 and args.slab_n > 0:
        slab = slab_object(args.slab_n, args.slab_iters, rank, world, local_rank, dist, torch)
        try:
            par = slab_parity(4096, 2, rank, world, local_rank, dist, torch)
            if par:
                slab["parity"] = par
                slab["parity_max_rel"] = par["parity_max_rel"]
                slab["iters"] = {"slab": par["iters_slab"], "single": par["iters_single"]}
        except Exception as ex:  # noqa: BLE001
            slab["parity"] = {"unavailable": str(ex)[:200]}
        line["slab"] = slab
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def grid_object(n, device, args, torch):
    """BASELINE config 3 on one GPU: cell-updates/s, iterations, and the roofline of the
    assembly and SpMV kernels.  L2 is flushed (256 MB write) between timed steps."""
    eng, f0 = make_grid(n, device)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=f"cuda:{device}")
    steps = max(2, min(args.steps, 5))
    eng.step(max(1, min(args.warmup, 3)))
    dev_s, iters = 0.0, 0
    for _ in range(steps):
        flush.zero_(); torch.cuda.synchronize()
        st = eng.step(1)
        dev_s += st["seconds_device"]; iters += st["iters_total"]
    flush.zero_(); torch.cuda.synchronize()
    prof, pst = profile_pass(eng, 1)
    rows, dom, peak, peak_src = roofline_from_profile(prof, stats=pst, cells_per_problem=n * n)
    out = {"workload": f"grid{n}", "value": n * n * steps / dev_s, "unit": UNIT, "steps": steps, "ms_per_step": 1e3 * dev_s / steps,
           "iters_per_step": iters / steps, "negatives": st["negatives"], "resid_last": st["resid_last"],
           "precond": PRECOND_NAMES.get(st.get("precond", 0), "?"),
           "kernel_launches_per_step": st["kernel_launches"],
           "l2": "256 MB flush between timed steps; the working set (~25 fine-grid arrays, 210 MB, plus the coarse levels) cycles through L2 within a step",
           "roofline": {"bound": "hbm", "peak": peak, "unit": "GB/s", "peak_source": peak_src, "dominant": dom, "kernels": rows}}
    # The event bracket around a launch (record, kernel, record on a plain stream) is longer than the kernel's slot in the
    # real step, which replays CUDA graphs: the difference per launch, measured here as
    #   (sum of all brackets of the profiled step - device time of the graph-replayed step) / launches,
    # is taken off every bracket, so that the in-step times add up to the real step (both sides CUDA events, same run).
    tot_ms = sum(r["ms_total"] for r in rows.values())
    tot_launches = sum(r["launches"] for r in rows.values())
    step_ms = 1e3 * dev_s / steps
    over_us = max(0.0, 1e3 * (tot_ms - step_ms) / max(tot_launches, 1))
    out["roofline"]["bracket"] = {"sum_of_brackets_ms": round(tot_ms, 3), "graph_step_ms": round(step_ms, 3), "launches": tot_launches,
                                  "overhead_us_per_launch": round(over_us, 2),
                                  "what": "us_per_launch = bracketed CUDA-event time - overhead_us_per_launch; the in-step times then add up to graph_step_ms"}
    for name, r in rows.items():
        if "us_per_launch" not in r:
            continue
        r["us_per_launch_bracketed"] = r["us_per_launch"]
        r["us_per_launch"] = round(max(r["us_per_launch"] - over_us, 0.1), 2)
        r["achieved_bracketed"], r["frac_bracketed"] = r["achieved"], r["frac"]
        gbs = r["achieved_bracketed"] * r["us_per_launch_bracketed"] / r["us_per_launch"]
        r["achieved"], r["frac"] = round(gbs, 1), round(gbs / peak, 4)
    out["roofline"]["sustained"] = sustained_kernels(eng, n * n, peak)
    out["time_dependent"] = time_dependent_leg(eng, n, torch)
    eng.close()
    if not args.no_cpu and not args.no_grid_cpu:
        out["cpu_baseline"] = cpu_baseline_grid(n, 1)
    return out


# the figures the north-star roofline target is quoted with - assembly 96 B per cell (104 with the
# predictor's yprev, 112 with the multigrid row weights: what the kernel really moves), SpMV 56 B per cell
SURVEY_BYTES = {"assembly": 96, "spmv_v": 56, "spmv_t": 56}


def north_star_rows(grid1024, grid4096, peak):
    """The kernels BASELINE.json's roofline target names (PPFV assembly, stencil SpMV) on the 1024^2 grid it is quoted on:
    IN-STEP (CUDA events around every launch of a real time step, L2 flushed before the step - the bracket includes the
    launch gap) and SUSTAINED (back-to-back launches; at 1024^2 the SpMV working set fits the 126 MB L2), with the
    DRAM-bound 4096^2 figures alongside.  frac uses SURVEY 8(d)'s bytes, frac_moved what the kernel really moves."""
    rows = {}
    for gname, g in (("grid1024", grid1024), ("grid4096", grid4096)):
        if not g:
            continue
        cells = int(gname[4:]) ** 2
        ins = (g.get("roofline") or {}).get("kernels", {})
        sus = (g.get("roofline") or {}).get("sustained", {}) if gname == "grid1024" else g.get("sustained", {})
        for k, sb in SURVEY_BYTES.items():
            for mode, src in (("in_step", ins), ("sustained", sus)):
                r = src.get(k)
                if not r or not r.get("us_per_launch"):
                    continue
                us = r["us_per_launch"]
                gbs = cells * sb / (us * 1e-6) / 1e9
                rows[f"{gname}_{k}_{mode}"] = {"us_per_launch": us, "bytes_per_cell": sb, "achieved": round(gbs, 1), "peak": peak, "unit": "GB/s",
                                               "frac": round(gbs / peak, 4), "bytes_per_cell_moved": r["bytes_per_cell"],
                                               "frac_moved": r["frac"], "launches": r.get("launches")}
                if "us_per_launch_bracketed" in r:
                    rows[f"{gname}_{k}_{mode}"].update({"us_per_launch_bracketed": r["us_per_launch_bracketed"],
                                                        "frac_bracketed": round(cells * sb / (r["us_per_launch_bracketed"] * 1e-6) / 1e9 / peak, 4)})
    return rows


def time_dependent_leg(eng, n, torch, steps=6):
    """Equation::update(t) on the device (Solver.cc:286-289): per step new D fields and Dirichlet data.  Blocking route:
    sy2d_set_coeffs + sy2d_set_bc then the step.  Overlapped route: the NEXT step's fields go through
    sy2d_set_coeffs_async / sy2d_set_bc_async (second buffer set, copy stream) from this thread while the step runs on
    another host thread.  Host-side evaluation of the fields is excluded (two precomputed field sets alternate)."""
    from sayram2d_b200 import fields
    xe, ye = fields.uniform_edges(n, n)
    G = fields.ay_G(xe, ye)
    Dxx, Dxy, Dyy, inv_tau = fields.synthetic_tensor(xe, ye)
    _, bct, lines = fields.ay_init_and_bc(xe, ye)
    sets = [(G, Dxx * a, Dxy * a, Dyy * a, inv_tau) for a in (1.0, 1.02)]
    out = {}
    try:
        eng.step(1)
        t0 = time.perf_counter()
        for k in range(steps):
            eng.set_coeffs(*sets[k % 2]); eng.set_bc(bct, *lines)
            eng.step(1)
        out["blocking_ms_per_step"] = 1e3 * (time.perf_counter() - t0) / steps
        eng.set_coeffs_async(*sets[0]); eng.set_bc_async(bct, *lines)
        eng.step(1)
        t0 = time.perf_counter()
        for k in range(steps):
            def stage_next(k=k):
                eng.set_coeffs_async(*sets[(k + 1) % 2]); eng.set_bc_async(bct, *lines)
            eng.step_overlapped(stage_next)
        out["overlapped_ms_per_step"] = 1e3 * (time.perf_counter() - t0) / steps
        eng.set_coeffs(*sets[0]); eng.set_bc(bct, *lines)
        t0 = time.perf_counter()
        eng.step(steps)
        out["static_ms_per_step"] = 1e3 * (time.perf_counter() - t0) / steps
        out["staged_bytes_per_step"] = 5 * n * n * 8
        out["note"] = "wall clock per time step incl. staging 5 fields + 4 boundary lines from pageable host arrays"
    except Exception as ex:  # noqa: BLE001
        out["unavailable"] = str(ex)[:200]
    return out


def sustained_kernels(eng, cells, peak, reps=20):
    """Back-to-back launches of one kernel between two CUDA events (sy2d_bench_kernel): the
    kernel's sustained rate without per-launch event gaps."""
    rows = {}
    for name in ("assembly", "p_update", "spmv_v", "s_update", "spmv_t", "xr_update"):
        ms = eng.bench_kernel(name, reps)
        gbs = cells * BYTES_PER_CELL[name] / (ms * 1e-3) / 1e9
        rows[name] = {"us_per_launch": round(1e3 * ms, 2), "achieved": round(gbs, 1), "frac": round(gbs / peak, 4),
                      "bytes_per_cell": BYTES_PER_CELL[name]}
    return rows


def big_grid_object(n, device):
    """DRAM-bound kernel rates: the same kernels on an n x n grid whose per-kernel inputs
    (n = 4096: 0.8-1.7 GB) are far larger than the 126 MB L2."""
    eng, f0 = make_grid(n, device)
    eng.set_options(engine=1)
    peak, peak_src = measured_peak()
    out = {"workload": f"grid{n}", "l2": f"per-kernel inputs {n * n * 48 / 1e6:.0f}-{n * n * 104 / 1e6:.0f} MB >> 126 MB L2: DRAM-bound, no flush needed",
           "peak": peak, "unit": "GB/s", "peak_source": peak_src, "sustained": sustained_kernels(eng, n * n, peak, reps=10)}
    # whole time steps at this size (multigrid-preconditioned BiCGSTAB; no CPU figure: the direct LU needs > 100 GB here)
    eng.step(1)
    st = eng.step(2)
    out.update({"value": n * n * st["steps"] / st["seconds_device"], "value_unit": UNIT, "ms_per_step": 1e3 * st["seconds_device"] / st["steps"],
                "iters_per_step": st["iters_total"] / st["steps"], "precond": PRECOND_NAMES.get(st.get("precond", 0), "?"),
                "resid_last": st["resid_last"], "negatives": st["negatives"]})
    eng.close()
    return out


def slab_object(n, iters, rank, world, local_rank, dist, torch):
    """BASELINE config 5: one n x n grid (config-3 fields) split into row slabs over the ranks, NCCL
    one-line halo exchange + all-gathered dots.  A direct solve is infeasible at this size and the
    Jacobi-BiCGSTAB needs O(1e4) iterations per step, so the run uses a FIXED ITERATION BUDGET:
    the first `iters` BiCGSTAB iterations of one time step (216 B/cell algorithmic traffic each)."""
    import sayram2d_b200 as sy
    from sayram2d_b200 import fields
    from sayram2d_b200.shard import slab_range
    ids = [sy.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    xe, ye = fields.uniform_edges(n, n)
    lo, hi = slab_range(n, rank, world)
    eng = sy.Engine(xe, ye, DT, device=local_rank, slab=(rank, world, ids[0]))
    assert (eng.i_lo, eng.i_hi) == (lo, hi)
    Dxx, Dxy, Dyy, inv_tau = fields.synthetic_tensor(xe, ye, rows=(lo, hi))
    eng.set_coeffs(fields.ay_G(xe, ye, rows=(lo, hi)), Dxx, Dxy, Dyy, inv_tau)
    del Dxx, Dxy, Dyy, inv_tau
    f0, bct, lines = fields.ay_init_and_bc(xe, ye, rows=(lo, hi))
    eng.set_bc(bct, *lines)
    eng.set_f(f0)
    del f0
    o = eng.options(); o.maxit = 16; o.check_every = 16; o.reserved[1] = 1; o.precond = 1   # the segmented x-line iteration
    eng._check(eng.lib.sy2d_set_options(eng._ctx, o))
    eng.step(1)                                   # warm-up: 16 iterations
    eng.set_f(fields.ay_init_and_bc(xe, ye, rows=(lo, hi))[0])
    o.maxit = iters
    eng._check(eng.lib.sy2d_set_options(eng._ctx, o))
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    st = eng.step(1)
    t = torch.tensor([st["seconds_device"]], dtype=torch.float64, device=f"cuda:{local_rank}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t.item())
    its = st["iters_total"]
    peak, _ = measured_peak()
    per_iter = 288.0 if st.get("precond", 0) == 1 else 216.0
    gbs = n * n * (104 + per_iter * its) / sec / 1e9
    out = {"workload": f"slab{n}", "n_gpus": world, "rows_per_gpu": hi - lo, "iterations": its, "seconds": sec,
           "ms_per_iteration": 1e3 * sec / max(its, 1), "cell_iterations_per_sec": n * n * its / sec,
           "algorithmic_GBps_aggregate": gbs, "frac_of_aggregate_hbm_peak": gbs / (peak * world),
           "precond": "xline16" if st.get("precond", 0) == 1 else "jacobi", "bytes_per_cell_iteration": per_iter,
           "budget": f"fixed iteration budget: assembly + first {its} BiCGSTAB iterations of one time step (not converged by design)",
           "exchange": "per iteration: 2 one-line halo ncclSend/Recv pairs (128 KB lines at 16384) + 3 all-gathers of 5 doubles"}
    # The same grid SOLVED: whole time steps with the multigrid-preconditioned BiCGSTAB (whole-line smoother made exact
    # across the ranks by the spike correction, every residual exchanges one halo row per level); value = cell-updates/s
    # of converged steps.
    try:
        o.maxit = 400; o.check_every = 4; o.reserved[1] = 0; o.precond = 2
        eng._check(eng.lib.sy2d_set_options(eng._ctx, o))
        eng.set_f(fields.ay_init_and_bc(xe, ye, rows=(lo, hi))[0])
        eng.step(1)                               # warm-up (allocations, first-step transient)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        st = eng.step(2)
        t = torch.tensor([st["seconds_device"]], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        neg = torch.tensor([float(st["negatives"])], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(neg, op=dist.ReduceOp.SUM)
        sec = float(t.item())
        out["solved"] = {"precond": "multigrid, whole x-lines coupled across the ranks (spike correction)", "steps": st["steps"], "ms_per_step": 1e3 * sec / st["steps"],
                         "iters_per_step": st["iters_total"] / st["steps"], "value": n * n * st["steps"] / sec, "unit": UNIT,
                         "resid_last": st["resid_last"], "negatives": int(neg.item())}
    except Exception as ex:  # e.g. rows per rank beyond what the line kernel covers
        out["solved"] = {"unavailable": str(ex)[:200]}
    eng.close()
    return out


def slab_parity(n, steps, rank, world, local_rank, dist, torch):
    """Parity certificate of the slab path: `steps` time steps of the n x n grid as `world` row slabs
    over NCCL against the SAME steps in one context on rank 0; max relative difference of f over the whole grid."""
    import sayram2d_b200 as sy
    from sayram2d_b200 import fields
    from sayram2d_b200.shard import slab_range
    dev = f"cuda:{local_rank}"
    ids = [sy.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    xe, ye = fields.uniform_edges(n, n)
    lo, hi = slab_range(n, rank, world)
    eng = sy.Engine(xe, ye, DT, device=local_rank, slab=(rank, world, ids[0]))
    eng.set_options(precond=2, check_every=1)
    Dxx, Dxy, Dyy, inv_tau = fields.synthetic_tensor(xe, ye, rows=(lo, hi))
    eng.set_coeffs(fields.ay_G(xe, ye, rows=(lo, hi)), Dxx, Dxy, Dyy, inv_tau)
    f0, bct, lines = fields.ay_init_and_bc(xe, ye, rows=(lo, hi))
    eng.set_bc(bct, *lines)
    eng.set_f(f0)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    st = eng.step(steps)
    rows_max = -(-n // world)
    mine = torch.zeros((rows_max, n), dtype=torch.float64, device=dev)
    mine[: hi - lo] = torch.from_numpy(eng.get_f()[0]).to(dev)
    eng.close()
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    out = None
    if rank == 0:
        f = np.empty((n, n))
        for r in range(world):
            l, h = slab_range(n, r, world)
            f[l:h] = parts[r][: h - l].cpu().numpy()
        ref = sy.Engine(xe, ye, DT, device=local_rank)
        ref.set_options(engine=1, precond=2, check_every=1)
        D = fields.synthetic_tensor(xe, ye)
        ref.set_coeffs(fields.ay_G(xe, ye), *D)
        f0g, bct, lines = fields.ay_init_and_bc(xe, ye)
        ref.set_bc(bct, *lines)
        ref.set_f(f0g)
        rs = ref.step(steps)
        fr = ref.get_f()[0]
        ref.close()
        out = {"grid": n, "steps": steps, "parity_max_rel": float(np.max(np.abs(f - fr) / np.abs(fr))),
               "iters_slab": st["iters_total"], "iters_single": rs["iters_total"], "resid_slab": st["resid_last"], "resid_single": rs["resid_last"],
               "ms_per_step_slab": 1e3 * st["seconds_device"] / steps, "ms_per_step_single": 1e3 * rs["seconds_device"] / steps,
               "timing_note": "both runs poll the host after EVERY iteration (check_every = 1, so that the iteration counts can be compared exactly) and include the first-use graph captures: the ms figures are not throughput numbers, slab.solved is",
               "what": f"{steps} time steps of {n}x{n} as {world} NCCL row slabs against one context on rank 0 (multigrid-preconditioned BiCGSTAB on both)"}
    dist.barrier()
    return out


def main():
    global NB_TOTAL
    # stdout carries ONE JSON line: native code that writes to file descriptor 1 (NCCL's version banner, whatever its debug
    # level says) is sent to stderr; Python's own stdout keeps the real descriptor
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        # NCCL prints its version banner on stdout, where the JSON line goes (the level may also come from an nccl.conf,
        # which an explicit environment value overrides); a level the user set explicitly is left alone
        os.environ["NCCL_DEBUG"] = "WARN"
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-grid1024", action="store_true")
    ap.add_argument("--members", type=int, default=NB_TOTAL, help="ensemble size (profiling runs use a smaller one)")
    ap.add_argument("--slab-n", type=int, default=16384, help="N>1 only: side of the single grid split into row slabs (0 = skip)")
    ap.add_argument("--slab-iters", type=int, default=48, help="iteration budget of the slab run")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline legs")
    ap.add_argument("--no-grid-cpu", action="store_true", help="skip the ~40 s CPU baseline of the 1024^2 grid")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    NB_TOTAL = args.members
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
