import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


# which reference case produced each fixture -> boundary description for the engine
CASE_OF = {"ay80": "AY", "lc80": "LC", "nu48x40": "AY", "syn64x48": "AY", "syn1024_sub": "AY"}


def load_golden(tag):
    g = np.load(os.path.join(GOLDEN, tag + ".npz"))
    d = {k: g[k] for k in g.files}
    if "meta" in d:
        d["meta"] = json.loads(str(d["meta"]))
    return d


def bc_for(case, xe, ye):
    """Boundary types and Dirichlet vertex lines of the AY / LC cases
    (Albert_Young.cc:42-92, Albert_Young_LC.cc:56-106) from the package's fields module."""
    from sayram2d_b200 import fields
    _, bct, lines = fields.ay_init_and_bc(xe, ye, lc=(case == "LC"))
    return bct, lines


def engine_from_golden(g, case, nbatch=1, device=0, **opts):
    from sayram2d_b200 import Engine
    xe, ye = g["x_edges"], g["y_edges"]
    dt = g["meta"]["dt"] if "meta" in g else 0.002
    eng = Engine(xe, ye, dt, nbatch=nbatch, device=device)
    if opts:
        eng.set_options(**opts)
    rep = lambda a: np.broadcast_to(a, (nbatch,) + a.shape).copy()
    eng.set_coeffs(rep(g["G"]), rep(g["Dxx"]), rep(g["Dxy"]), rep(g["Dyy"]), rep(g["inv_tau"]))
    bct, lines = bc_for(case, xe, ye)
    eng.set_bc(bct, *lines)
    eng.set_f(rep(g["f_0"]))
    return eng


def max_rel(a, b):
    return float(np.max(np.abs(a - b) / np.abs(b)))


@pytest.fixture(scope="session")
def d_table():
    import h5min
    return h5min.load_d_table(os.path.join(ROOT, "data", "D", "AlbertYoung_chorus.h5"))
