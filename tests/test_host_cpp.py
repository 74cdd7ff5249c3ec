"""CPU tests of the C++ host layer (sayram2d_b200/host): Parameters / Ini_reader / Grid2D /
Mesh / Albert_Young / Albert_Young_LC keep the reference's API and reproduce the reference's
case data (compared with the fields the reference build dumped into tests/golden)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden, max_rel

HOST = os.path.join(ROOT, "sayram2d_b200", "host")
EXE = os.path.join(ROOT, "tests", "host", "_host_check")


@pytest.fixture(scope="module")
def host_check():
    srcs = [os.path.join(HOST, f) for f in ("Albert_Young.cc", "Albert_Young_IO.cc", "Mesh.cc", "Parameters.cc")]
    srcs.append(os.path.join(ROOT, "tests", "host", "host_check.cc"))
    deps = srcs + [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".h")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-I" + HOST, *srcs, "-o", EXE], check=True)
    return EXE


def test_validation_and_mesh_api(host_check):
    out = subprocess.run([host_check, "errors"], capture_output=True, text=True)
    assert out.returncode == 0 and "FAIL" not in out.stdout, out.stdout


def test_ini_reader_and_parameters(host_check, tmp_path):
    ini = tmp_path / "q.ini"
    ini.write_text("; comment\n[basic]\nrun_id = q\nNALPHA0 = 10\nnE = 12\nalpha0min = 5\nalpha0max = 90\nEmin = 0.2\nEmax = 5\n"
                   "T = 1.0\nnsteps = 505\nflag = no\n;commented = 1\n[diagnostics]\nnplots = 10\n[diffusion_coefficients]\ndID = X\n")
    out = subprocess.run([host_check, "ini", str(ini)], capture_output=True, text=True, cwd=tmp_path)
    assert out.returncode == 0 and "FAIL" not in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("case,ini,tag", [("AY", "p.ini", "ay80"), ("LC", "p_AlbertYoungLC.ini", "lc80")])
def test_case_data_matches_reference(host_check, tmp_path, case, ini, tag):
    os.symlink(os.path.join(ROOT, "data", "D"), tmp_path / "D")
    out = subprocess.run([host_check, "dump", case, os.path.join(ROOT, "data", ini), str(tmp_path)], capture_output=True, text=True, cwd=tmp_path)
    assert out.returncode == 0, out.stderr
    assert "dt 0.002 static 1" in out.stdout
    g = load_golden(tag)
    for name in ("x_edges", "y_edges"):
        assert np.array_equal(np.load(tmp_path / f"{name}.npy"), g[name])
    for name in ("G", "Dxx", "Dxy", "Dyy", "inv_tau"):
        got, ref = np.load(tmp_path / f"{name}.npy"), g[name]
        assert np.max(np.abs(got - ref)) <= 4e-15 * np.max(np.abs(ref)), name
    assert max_rel(np.load(tmp_path / "f_0.npy"), g["f_0"]) < 1e-13
    from sayram2d_b200 import fields
    _, bct, lines = fields.ay_init_and_bc(g["x_edges"], g["y_edges"], lc=(case == "LC"))
    assert list(np.load(tmp_path / "bc_types.npy").astype(int)) == list(bct)
    for s in range(4):
        has = bool(np.load(tmp_path / f"bc_has{s}.npy")[0])
        assert has == (lines[s] is not None)
        if has:
            assert np.max(np.abs(np.load(tmp_path / f"bc_line{s}.npy") - lines[s])) <= 1e-15 * max(np.max(np.abs(lines[s])), 1.0)


def test_dropin_solver_compiles_against_both_header_sets():
    """sayram2d_b200/dropin/Solver.{h,cc} only use the public Mesh/Equation API: they build
    against this repo's host classes (always) and against the reference's own headers
    (oracle/_ref/*_dropin, built by `make -C oracle ref` where /root/reference exists)."""
    src = os.path.join(ROOT, "sayram2d_b200", "dropin", "Solver.cc")
    obj = os.path.join(ROOT, "tests", "host", "_solver_dropin.o")
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-I" + HOST, "-I" + os.path.join(ROOT, "sayram2d_b200", "dropin"),
                    "-I" + os.path.join(ROOT, "include"), "-c", src, "-o", obj], check=True)
    text = open(os.path.join(ROOT, "sayram2d_b200", "dropin", "Solver.h")).read()
    for decl in ("Solver(const Mesh& m_in, Equation* eqp);", "void update();", "double t() const", "const Xtensor2d& f() const;",
                 "double f(const Ind& ind) const"):
        assert decl in text, decl   # the reference's public interface, source/Solver.h:20-25
    if os.path.isdir("/root/reference/source"):
        assert os.path.exists(os.path.join(ROOT, "oracle", "_ref", "sayram-2d_AY_dropin"))
