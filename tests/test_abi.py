"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every
symbol include/sayram2d.h declares, validates arguments the way the reference
does (Grid2D.h:44-67) and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT

import sayram2d_b200 as sy


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sayram2d.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sy2d_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(sy.library_path())
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sayram2d.h but not exported"


def test_struct_layouts_match_header():
    text = open(os.path.join(ROOT, "include", "sayram2d.h")).read()
    assert ctypes.sizeof(sy.Options) == 8 + 4 * 8 + 4 * 3 + 4  # double + 8 ints + reserved[3] + padding
    assert "reserved[3]" in text
    n_k = len(re.findall(r"^\s*SY2D_K_[A-Z_]+\b", text, flags=re.M)) - 1  # minus SY2D_K_COUNT
    assert n_k == len(sy.K_NAMES)


def test_build_info_names_the_target():
    info = sy.load_library().sy2d_build_info().decode()
    assert "sm_100a" in info and "fp64" in info


def test_argument_validation_messages_follow_reference():
    lib = sy.load_library()
    with pytest.raises(sy.Sy2dError) as e:
        sy.Engine(np.array([0.0, 1.0, 1.0]), np.linspace(0, 1, 4), 0.1)
    assert e.value.code == -1 and "x_edges must be strictly increasing at i=1" in str(e.value)  # Grid2D.h:52-57
    with pytest.raises(sy.Sy2dError) as e:
        sy.Engine(np.linspace(0, 1, 4), np.array([0.0, 0.5, 0.4]), 0.1)
    assert "y_edges must be strictly increasing at j=1" in str(e.value)                          # Grid2D.h:60-65
    with pytest.raises(sy.Sy2dError) as e:
        sy.Engine(np.linspace(0, 1, 4), np.linspace(0, 1, 4), 0.0)
    assert e.value.code == -1
    o = sy.Options()
    assert lib.sy2d_default_options(ctypes.byref(o)) == 0
    assert o.tol == 1e-14 and o.maxit >= 1000 and o.predictor == 2


def test_no_cpu_fallback():
    lib = sy.load_library()
    if lib.sy2d_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(sy.Sy2dError) as e:
        sy.Engine(np.linspace(0, 1, 9), np.linspace(0, 1, 9), 0.1)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_never_touches_the_oracle():
    """The package must not import, link or execute anything under oracle/ (or tests/)."""
    pkg = os.path.join(ROOT, "sayram2d_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.replace("the oracle", "").lower() or "ppfv_oracle" not in src, f
                assert "ppfv_oracle" not in src and "oracle/" not in src and "_emul_host" not in src, f
