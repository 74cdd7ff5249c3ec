#!/usr/bin/env python
"""Generates the golden fixtures in this directory from the reference itself.

Runs oracle/_ref/ref_driver - the reference's UNMODIFIED Solver.cc / Mesh.cc /
Parameters.cc / Cases/*.cc compiled against oracle/shim (see oracle/Makefile) -
and packs its outputs into small .npz files.  Needs /root/reference (to build
oracle/_ref) and therefore only runs in the build container:

    make -C oracle ref && python tests/golden/make_golden.py

Fixtures (all f64, (nx,ny) row-major j-fastest = reference layout):
  ay80.npz / lc80.npz : data/p.ini / data/p_AlbertYoungLC.ini, 500 steps; inputs
                        (edges, G, Dxx, Dxy, Dyy, inv_tau), snapshots f_0,f_1,f_5,f_10
                        (= t 0, 0.1, 0.5, 1.0 day) and the assembled (M,R) of
                        steps 1 and 250 as five diagonals + R.
  nu48x40.npz         : AY case on a warped (non-uniform) 48x40 grid, 20 steps.
  syn64x48.npz        : BASELINE config 3 tensor on a 64x48 grid, 10 steps.
  ens_members.npz     : BASELINE config 4 members 0, 63, 2047, 4095 after 500 steps.
  syn1024_sub.npz     : config 3 at 1024x1024, 3 steps, f sub-sampled every 8 cells.
  td64.npz            : time-dependent user case (D(t), 1/tau(t), Dirichlet data(t); ref_driver.cc
                        Time_Dependent) on a 64x64 grid, 60 steps, snapshots every 20 steps.
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from ppfv_oracle import ensemble_member_scales  # noqa: E402


def ini_text(run_id, nx, ny, amin, nsteps=500, nplots=10):
    return (f"[basic]\nrun_id = {run_id}\nnalpha0 = {nx}\nnE = {ny}\nalpha0min = {amin}\nalpha0max = 90\n"
            f"Emin = 0.2\nEmax = 5\nT = 1.0\nnsteps = {nsteps}\n[diagnostics]\nnplots = {nplots}\n"
            f"[diffusion_coefficients]\ndID = AlbertYoung_chorus\n")


def drive(work, case, ini, out, *extra):
    cmd = [DRIVER, "--case", case, "--ini", ini, "--out", out, *map(str, extra)]
    res = subprocess.run(cmd, cwd=work, check=True, capture_output=True, text=True)
    return json.loads(res.stdout.strip().splitlines()[-1])


def load(work, out, names):
    return {n: np.load(os.path.join(work, out, n + ".npy")) for n in names}


INPUTS = ["x_edges", "y_edges", "G", "Dxx", "Dxy", "Dyy", "inv_tau"]
OPS = ["diag", "W", "E", "S", "N", "R"]


def main():
    if not os.path.exists(DRIVER):
        sys.exit("build the oracle first: make -C oracle ref")
    with tempfile.TemporaryDirectory() as work:
        os.symlink(os.path.join(ROOT, "data", "D"), os.path.join(work, "D"))
        for tag, case, ini in (("ay80", "AY", "p.ini"), ("lc80", "LC", "p_AlbertYoungLC.ini")):
            meta = drive(work, case, os.path.join(ROOT, "data", ini), tag, "--dump-op", "1,250")
            d = load(work, tag, INPUTS + ["f_0", "f_1", "f_5", "f_10"] + [f"op{s}_{n}" for s in (1, 250) for n in OPS])
            np.savez_compressed(os.path.join(HERE, tag + ".npz"), meta=json.dumps(meta), **d)
            print(tag, meta)

        open(os.path.join(work, "nu.ini"), "w").write(ini_text("nu48x40", 48, 40, 5))
        meta = drive(work, "AY", "nu.ini", "nu", "--steps", 20, "--stretch", 0.6, "--dump-op", "1,20")
        d = load(work, "nu", INPUTS + ["f_0", "f_1", "f_20"] + [f"op{s}_{n}" for s in (1, 20) for n in OPS])
        np.savez_compressed(os.path.join(HERE, "nu48x40.npz"), meta=json.dumps(meta), stretch=0.6, **d)
        print("nu48x40", meta)

        open(os.path.join(work, "syn.ini"), "w").write(ini_text("syn64x48", 64, 48, 5))
        meta = drive(work, "SYN", "syn.ini", "syn", "--steps", 10, "--dump-op", "1,10")
        d = load(work, "syn", INPUTS + ["f_0", "f_1", "f_10"] + [f"op{s}_{n}" for s in (1, 10) for n in OPS])
        np.savez_compressed(os.path.join(HERE, "syn64x48.npz"), meta=json.dumps(meta), **d)
        print("syn64x48", meta)

        ens = {}
        for mth in (0, 63, 2047, 4095):
            a, b = ensemble_member_scales(mth)
            meta = drive(work, "ENS", os.path.join(ROOT, "data", "p_AlbertYoungLC.ini"), f"ens{mth}",
                         "--member", repr(a), repr(b))
            ens[f"f10_m{mth}"] = np.load(os.path.join(work, f"ens{mth}", "f_10.npy"))
            ens[f"f1_m{mth}"] = np.load(os.path.join(work, f"ens{mth}", "f_1.npy"))
            print("ens", mth, a, b, meta)
        np.savez_compressed(os.path.join(HERE, "ens_members.npz"), **ens)

        open(os.path.join(work, "td.ini"), "w").write(ini_text("td64", 64, 64, 5))
        meta = drive(work, "TD", "td.ini", "td", "--steps", 60, "--every", 20)
        d = load(work, "td", ["x_edges", "y_edges", "G", "f_0", "f_1", "f_2", "f_3"])
        np.savez_compressed(os.path.join(HERE, "td64.npz"), meta=json.dumps(meta), **d)
        print("td64", meta)
        if "--only-td" in sys.argv:
            return

        open(os.path.join(work, "syn1024.ini"), "w").write(ini_text("syn1024", 1024, 1024, 5))
        meta = drive(work, "SYN", "syn1024.ini", "syn1024", "--steps", 3)
        sub = {f"f_{k}": np.load(os.path.join(work, "syn1024", f"f_{k}.npy"))[3::8, 5::8].copy() for k in (0, 1, 3)}
        np.savez_compressed(os.path.join(HERE, "syn1024_sub.npz"), meta=json.dumps(meta), **sub)
        print("syn1024", meta)


if __name__ == "__main__":
    main()
