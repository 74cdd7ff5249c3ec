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


def test_hdf5_writer_round_trip_and_structures(host_check, tmp_path):
    """h5lite::Writer (the /alpha0, /logEN, /f/<k>, /t output file of main.cc:58-89 without libhdf5):
    (1) the file is read back by the independent Python reader (oracle/h5min.py, itself validated on a
    file written by libhdf5) and by the C++ reader; (2) a "mirror" file with the names and shapes of
    data/D/AlbertYoung_chorus.h5 has byte-identical superblock parameters, group structures and dataset
    header messages as that libhdf5-written file, addresses aside."""
    import struct
    import h5min
    out = tmp_path / "o.h5"
    res = subprocess.run([host_check, "h5write", str(out)], capture_output=True, text=True)
    assert res.returncode == 0 and "datasets 16 f/12[23] 1211.5" in res.stdout, res.stdout + res.stderr
    h = h5min.H5File(str(out))
    assert sorted(h.datasets) == sorted(["/alpha0", "/logEN", "/t"] + [f"/f/{k}" for k in range(13)])
    assert np.array_equal(h.read("/alpha0"), 5.0 + 0.5 * np.arange(6)) and np.array_equal(h.read("/t"), 0.5 * np.arange(13))
    for k in range(13):
        assert np.array_equal(h.read(f"/f/{k}"), (100.0 * k + 0.5 * np.arange(24)).reshape(6, 4))
    raw = out.read_bytes()
    assert struct.unpack_from("<Q", raw, 40)[0] == len(raw)                  # end-of-file address
    assert struct.unpack_from("<H", raw, 16)[0] == 7                         # leaf K raised for the 13-member group /f

    mir = tmp_path / "m.h5"
    assert subprocess.run([host_check, "h5write", str(mir), "mirror"], capture_output=True, text=True).returncode == 0
    a = mir.read_bytes()
    b = open(os.path.join(ROOT, "data", "D", "AlbertYoung_chorus.h5"), "rb").read()
    assert a[:40] == b[:40] and a[48:64] == b[48:64]                         # superblock up to the EOF address, driver info, root name
    ra, rb = struct.unpack_from("<Q", a, 64)[0], struct.unpack_from("<Q", b, 64)[0]
    assert a[ra:ra + 24] == b[rb:rb + 24]                                    # root object header + symbol-table message head
    bta, hpa = struct.unpack_from("<QQ", a, 80)
    btb, hpb = struct.unpack_from("<QQ", b, 80)
    assert a[bta:bta + 32] == b[btb:btb + 32]                                # B-tree node head and key 0
    assert struct.unpack_from("<Q", a, bta + 40)[0] == struct.unpack_from("<Q", b, btb + 40)[0]   # key 1: offset of the largest name
    da, db = struct.unpack_from("<Q", a, hpa + 24)[0], struct.unpack_from("<Q", b, hpb + 24)[0]
    assert a[hpa:hpa + 8] == b[hpb:hpb + 8] and a[da:da + 48] == b[db:db + 48]                    # heap: names at the same offsets
    sa, sb = struct.unpack_from("<Q", a, bta + 32)[0], struct.unpack_from("<Q", b, btb + 32)[0]
    assert a[sa:sa + 8] == b[sb:sb + 8]                                      # SNOD head: version, 5 symbols
    ha, hb = h5min.H5File(str(mir)), h5min.H5File(os.path.join(ROOT, "data", "D", "AlbertYoung_chorus.h5"))
    assert sorted(ha.datasets) == sorted(hb.datasets)
    for k in range(5):                                                       # entries in the same (name) order; dataset headers
        ea, eb = sa + 8 + 40 * k, sb + 8 + 40 * k
        assert a[ea:ea + 8] == b[eb:eb + 8] and a[ea + 16:ea + 40] == b[eb + 16:eb + 40]
        oa, ob = struct.unpack_from("<Q", a, ea + 8)[0], struct.unpack_from("<Q", b, eb + 8)[0]
        hdr_a, hdr_b = bytearray(a[oa:oa + 16 + 256]), bytearray(b[ob:ob + 16 + 256])
        rank = hdr_a[16 + 8 + 1]
        lay = 16 + (8 + 8 + 16 * rank) + (8 + 24) + (8 + 8) + 8 + 2           # offset of the layout message's data address
        assert hdr_a[lay + 8:lay + 16] == hdr_b[lay + 8:lay + 16]            # same dataset size
        hdr_a[lay:lay + 8] = hdr_b[lay:lay + 8] = b"\0" * 8                  # the raw-data address is the only difference
        assert hdr_a == hdr_b, k


def test_hdf5_reader_reads_other_table_formats(host_check, tmp_path):
    """h5lite::File on D tables that are NOT laid out like data/D/AlbertYoung_chorus.h5 (Albert_Young_IO.cc:17-36 takes any
    D/<dID>.h5): files hand-packed by tests/h5craft.py - an independent writer of the HDF5 on-disk structures - with chunked
    storage (two-level chunk B-tree, ragged edge chunks) and the shuffle + deflate (+ fletcher32) filters as h5py writes them
    with compression="gzip", f32 and integer data, compact datasets, and the new-style format (superblock v2 / v3, version-2
    object headers with a continuation block, compact link messages, nested groups)."""
    import h5craft
    rng = np.random.default_rng(11)
    Daa = rng.uniform(0.0, 3e-4, (91, 49))
    Dap = rng.normal(0.0, 1e-5, (91, 49))
    E = np.linspace(0.1, 5.0, 49)
    old = {"alpha0": dict(data=np.arange(91.0), code="f4", layout="contiguous"),
           "E": dict(data=E, code="f8", layout="compact"),
           "Daa": dict(data=Daa, code="f8", layout="chunked", chunk=(32, 20), filters=[(2, [8]), (1, [4])]),
           "Dap": dict(data=Dap, code="f8", layout="chunked", chunk=(91, 49), filters=[(1, [9]), (3, [])], pipeline_version=2),
           "Dpp": dict(data=Daa.T.copy(), code="f8", layout="chunked", chunk=(7, 91)),
           "n": dict(data=np.array([[-3, 70000], [5, -2]]), code="i4", layout="contiguous"),
           "m": dict(data=np.array([1, 65535, 17]), code="u2", layout="chunked", chunk=(2,), filters=[(2, [2])])}
    f_old = tmp_path / "old.h5"
    h5craft.write_old_style(str(f_old), old)
    new = {"alpha0": dict(data=np.arange(91.0), code="f8", layout="contiguous"),
           "tables": {"Daa": dict(data=Daa, code="f8", layout="chunked", chunk=(16, 49), filters=[(2, [8]), (1, [6])]),
                      "Dap": dict(data=Dap, code="f4", layout="contiguous"),
                      "deep": {"E": dict(data=E, code="f8", layout="compact")}},
           "count": dict(data=np.array([2 ** 40, -7]), code="i8", layout="compact")}
    cases = [(f_old, {("/" + k): v for k, v in old.items()})]
    for ver, order in ((2, False), (3, True)):
        f_new = tmp_path / f"new{ver}.h5"
        h5craft.write_new_style(str(f_new), new, superblock_version=ver, track_order=order)
        flat = {"/alpha0": new["alpha0"], "/tables/Daa": new["tables"]["Daa"], "/tables/Dap": new["tables"]["Dap"],
                "/tables/deep/E": new["tables"]["deep"]["E"], "/count": new["count"]}
        cases.append((f_new, flat))
    for path, expect in cases:
        out = tmp_path / (path.name + ".d")
        out.mkdir()
        res = subprocess.run([host_check, "h5read", str(path), str(out)], capture_output=True, text=True)
        assert res.returncode == 0, res.stdout + res.stderr
        assert sorted(res.stdout.split()) == sorted(expect), res.stdout
        for name, d in expect.items():
            got = np.load(out / (name.replace("/", "_") + ".npy"))
            want = np.asarray(d["data"]).astype("<" + d["code"]).astype(np.float64)
            assert got.shape == want.shape and np.array_equal(got, want), name
    # the shipped libhdf5-written table still reads, and unsupported structures are reported, not misread
    out = tmp_path / "ay.d"
    out.mkdir()
    res = subprocess.run([host_check, "h5read", os.path.join(ROOT, "data", "D", "AlbertYoung_chorus.h5"), str(out)], capture_output=True, text=True)
    assert res.returncode == 0 and sorted(res.stdout.split()) == ["/Daa", "/Dap", "/Dpp", "/E", "/alpha0"]
    assert np.load(out / "_Daa.npy").shape == (91, 49) and abs(np.load(out / "_E.npy")[48] - 5.0) < 1e-12
    bad = dict(old)
    bad["Daa"] = dict(data=Daa, code="f8", layout="chunked", chunk=(32, 20), filters=[(32000, [1])])   # lzf
    f_bad = tmp_path / "bad.h5"
    h5craft.write_old_style(str(f_bad), bad)
    res = subprocess.run([host_check, "h5read", str(f_bad), str(out)], capture_output=True, text=True)
    assert res.returncode == 3 and "filter 32000 is not supported" in res.stdout
