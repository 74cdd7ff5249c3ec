"""Builds the CUDA library and the C++ host binary in-tree (nvcc, sm_100a).

    python -m sayram2d_b200.build [--force]

nvcc cross-compiles without a GPU; the resulting files travel to the GPU box with
the repo snapshot (they are git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "lib", "libsayram2d_b200.so")
BIN = os.path.join(PKG, "bin", "sayram2d")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CUDA_SOURCES = [os.path.join(PKG, "csrc", "sy2d_api.cu")]
CUDA_HEADERS = sorted(os.path.join(PKG, "csrc", h) for h in os.listdir(os.path.join(PKG, "csrc")) if h.endswith((".cuh", ".h"))) + [
    os.path.join(ROOT, "include", "sayram2d.h")]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    print("+", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)


def build_library(force=False):
    if force or _stale(LIB, CUDA_SOURCES + CUDA_HEADERS):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        _run([NVCC, "-O3", "-std=c++17", *ARCH, "-lineinfo", "-Xcompiler", "-fPIC", "-shared",
              "-o", LIB, *CUDA_SOURCES, "-ldl"])
    return LIB


DROPIN = os.path.join(PKG, "dropin")  # Solver.{h,cc}: compile against these host classes OR the reference's


def host_sources():
    hdir = os.path.join(PKG, "host")
    if not os.path.isdir(hdir):
        return []
    return sorted(os.path.join(hdir, f) for f in os.listdir(hdir) if f.endswith(".cc")) + [os.path.join(DROPIN, "Solver.cc")]


def build_host(force=False):
    """C++ host layer (reference API) + main, linked against the CUDA library."""
    srcs = host_sources()
    if not srcs:
        return None
    hdir = os.path.join(PKG, "host")
    deps = srcs + [os.path.join(hdir, f) for f in os.listdir(hdir) if f.endswith(".h")] + [LIB, os.path.join(DROPIN, "Solver.h")]
    if force or _stale(BIN, deps):
        os.makedirs(os.path.dirname(BIN), exist_ok=True)
        _run(["/usr/bin/g++", "-O2", "-std=c++17", "-I" + hdir, "-I" + DROPIN, "-I" + os.path.join(ROOT, "include"), *srcs,
              "-L" + os.path.dirname(LIB), "-lsayram2d_b200", "-Wl,-rpath,$ORIGIN/../lib", "-o", BIN])
    return BIN


TD_BIN = os.path.join(PKG, "bin", "sayram2d_td")


def build_examples(force=False):
    """examples/time_dependent_case.cc: a user-defined time-dependent Equation run with the GPU Solver."""
    hdir = os.path.join(PKG, "host")
    src = os.path.join(PKG, "examples", "time_dependent_case.cc")
    if not os.path.exists(src):
        return None
    srcs = [s for s in host_sources() if not s.endswith(os.sep + "main.cc")] + [src]
    deps = srcs + [os.path.join(hdir, f) for f in os.listdir(hdir) if f.endswith(".h")] + [LIB, os.path.join(DROPIN, "Solver.h")]
    if force or _stale(TD_BIN, deps):
        os.makedirs(os.path.dirname(TD_BIN), exist_ok=True)
        _run(["/usr/bin/g++", "-O2", "-std=c++17", "-I" + hdir, "-I" + DROPIN, "-I" + os.path.join(ROOT, "include"), *srcs,
              "-L" + os.path.dirname(LIB), "-lsayram2d_b200", "-Wl,-rpath,$ORIGIN/../lib", "-o", TD_BIN])
    return TD_BIN


def build_all(force=False):
    build_library(force)
    build_host(force)
    build_examples(force)


if __name__ == "__main__":
    build_all("--force" in sys.argv)
