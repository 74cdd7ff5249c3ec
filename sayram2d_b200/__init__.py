"""sayram2d_b200 - B200-native (sm_100a) time-step engine for Sayram-2D.

The product is the CUDA library behind the C ABI of include/sayram2d.h
(sayram2d_b200/lib/libsayram2d_b200.so, built by sayram2d_b200.build) and the C++
host layer in sayram2d_b200/host/ that keeps the reference's Parameters / Mesh /
Equation / Solver API.  This Python package is a thin ctypes binding used by the
tests and bench.py; it has no CPU path and raises if the library is missing.
"""
from .engine import (Engine, Sy2dError, Options, load_library, library_path, nccl_unique_id, K_NAMES,  # noqa: F401
                     LocalGroup, run_local_slabs, measure_peaks)
from . import fields  # noqa: F401

__all__ = ["Engine", "Sy2dError", "Options", "load_library", "library_path", "fields", "K_NAMES"]
