"""CPU tier: problem data files (SCS(write_data) / SCS(read_data), S/src/rw.c:240-315) through the C ABI
(host-only entry points: no device needed) against
  * tests/golden/rw_ref_mixed.bin, written by the compiled reference (64-bit-int python build, so reading
    it exercises the integer-width conversion of rw.c:74-101),
  * the compiled reference itself when oracle/_ref is on this box: files written by either library are
    byte-identical at equal integer width, and each library reads the other's files,
  * a hand-packed file (the layout documented in csrc/rw.cu),
and the CSV trace header against the reference's own trace (tests/golden/rw_ref_trace.csv).
Bit-exact: this is byte / integer work."""
import ctypes as C
import os
import struct
import sys

import numpy as np
import pytest
import scipy.sparse as sp

import scs_python_b200 as scsb
from scs_python_b200 import _scs_b200 as B

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden_rw as G  # noqa: E402  (the seeded problems behind the fixtures)


def _same_problem(data, K, stg, rd, rK, rstg):
    assert (rd["A"] != sp.csc_matrix(data["A"])).nnz == 0 and rd["A"].shape == data["A"].shape
    P = sp.triu(data["P"], format="csc") if data.get("P") is not None else None
    if P is None:
        assert rd["P"] is None
    else:
        assert (rd["P"] != P).nnz == 0
    assert np.array_equal(rd["b"], data["b"]) and np.array_equal(rd["c"], data["c"])
    for key in ("z", "l", "ep", "ed"):
        assert rK.get(key, 0) == K.get(key, 0)
    for key in ("q", "s"):
        assert list(rK.get(key, [])) == list(K.get(key, []))
    for key in ("bu", "bl", "p"):
        assert np.array_equal(np.asarray(rK.get(key, []), float), np.asarray(K.get(key, []), float))
    for key, val in stg.items():
        assert rstg[key] == val, key


def test_read_reference_written_file_with_64bit_ints():
    data, K, stg = G.rw_problem()
    path = os.path.join(HERE, "golden", "rw_ref_mixed.bin")
    assert struct.unpack("<II", open(path, "rb").read(8)) == (8, 8)
    rd, rK, rstg = scsb.read_data(path)
    _same_problem(data, K, stg, rd, rK, rstg)
    assert rstg["verbose"] is False


def test_write_then_read_roundtrip_and_layout(tmp_path):
    data, K, stg = G.rw_problem()
    path = str(tmp_path / "prob.bin")
    scsb.write_data(path, data, K, **stg)
    raw = open(path, "rb").read()
    isz, fsz, vlen = struct.unpack("<III", raw[:12])
    assert (isz, fsz) == (4, 8) and raw[12:12 + vlen].decode() == scsb.__version__
    off = 12 + vlen
    assert struct.unpack("<iii", raw[off:off + 12]) == (K["z"], K["l"], len(K["bu"]) + 1)   # cone: z l bsize
    rd, rK, rstg = scsb.read_data(path)
    _same_problem(data, K, stg, rd, rK, rstg)
    # no P, no optional cones, default settings
    d2 = dict(A=data["A"], b=data["b"], c=data["c"])
    p2 = str(tmp_path / "lp.bin")
    scsb.write_data(p2, d2, dict(l=int(data["A"].shape[0])))
    rd2, rK2, rstg2 = scsb.read_data(p2)
    assert rd2["P"] is None and rK2 == dict(z=0, l=data["A"].shape[0], ep=0, ed=0)
    assert rstg2["max_iters"] == 100000 and rstg2["eps_abs"] == 1e-4 and rstg2["acceleration_lookback"] == 10


def test_hand_packed_file_with_64bit_ints(tmp_path):
    """the documented layout, written field by field with 8-byte integers"""
    m, n = 3, 2
    Ap, Ai, Ax = [0, 2, 3], [0, 2, 1], [1.5, -2.0, 4.0]
    Pp, Pi, Px = [0, 1, 2], [0, 1], [2.0, 3.0]
    b, c = [1.0, 2.0, 3.0], [-1.0, 0.5]
    ver = scsb.__version__.encode()
    q = lambda *v: struct.pack("<%dq" % len(v), *v)
    d = lambda *v: struct.pack("<%dd" % len(v), *v)
    blob = struct.pack("<III", 8, 8, len(ver)) + ver
    blob += q(1, 2, 0) + q(0) + q(0) + q(0, 0) + q(0)                                    # cone: z=1 l=2
    blob += q(m, n) + d(*b) + d(*c) + q(m, n) + q(*Ap) + d(*Ax) + q(*Ai)                # data + A
    blob += q(1) + q(n, n) + q(*Pp) + d(*Px) + q(*Pi)                                    # has_p + P
    blob += q(1) + d(0.1, 1e-6) + q(500) + d(1e-5, 1e-6, 1e-7, 1.5) + q(0, 0, 10, 10, 1) + d(1e-8, 1.0) + q(1)
    path = str(tmp_path / "hand.bin")
    open(path, "wb").write(blob)
    rd, rK, rstg = scsb.read_data(path)
    assert rd["A"].shape == (m, n) and np.array_equal(rd["A"].toarray(), [[1.5, 0], [0, 4.0], [-2.0, 0]])
    assert np.array_equal(rd["P"].toarray(), [[2.0, 0], [0, 3.0]])
    assert list(rd["b"]) == b and list(rd["c"]) == c and rK == dict(z=1, l=2, ep=0, ed=0)
    assert rstg["max_iters"] == 500 and rstg["eps_abs"] == 1e-5 and rstg["eps_rel"] == 1e-6 and rstg["adaptive_scale"] is True


def test_read_errors(tmp_path):
    with pytest.raises(ValueError):
        scsb.read_data(str(tmp_path / "does_not_exist.bin"))
    data, K, stg = G.rw_problem()
    path = str(tmp_path / "prob.bin")
    scsb.write_data(path, data, K, **stg)
    raw = open(path, "rb").read()
    open(path, "wb").write(raw[:len(raw) // 2])                       # truncated
    with pytest.raises(ValueError):
        scsb.read_data(path)
    open(path, "wb").write(struct.pack("<II", 4, 4) + raw[8:])        # single-precision file (rw.c:286-293)
    with pytest.raises(ValueError):
        scsb.read_data(path)
    with pytest.raises(ValueError):
        scsb.write_data(str(tmp_path / "no_such_dir" / "x.bin"), data, K)


def test_csv_header_matches_reference_trace():
    ref = open(os.path.join(HERE, "golden", "rw_ref_trace.csv")).readline().strip()
    ref_cols = [c for c in ref.split(",") if c]
    ours = [c for c in B.csv_header().split(",") if c]
    # the reference's USE_LAPACK build names five spectral-cone columns it never fills (rw.c:396-402, 466-472)
    assert ours == ref_cols[:len(ours)] and len(ours) == 62
    assert set(ref_cols[len(ours):]) <= {"spectral_Newton_iter", "plain_Newton_success", "res_dual_spectral",
                                        "res_pri_spectral", "comp_spectral"}


# ------------------------------------------------------------------ against the compiled reference --
def _reflib():
    p = os.path.join(ROOT, "oracle", "_ref", "libscsindir.so")
    if not os.path.exists(p):
        pytest.skip("oracle/_ref not built on this box")
    return C.CDLL(p)


def test_files_are_byte_identical_to_the_reference_library(tmp_path):
    """same integer width (the C libraries of oracle/_ref are 32-bit-int builds): same bytes"""
    lib = _reflib()
    data, K, stg = G.rw_problem()
    ours = str(tmp_path / "ours.bin")
    scsb.write_data(ours, data, K, **stg)
    # the reference reads our file, then writes it back with its own writer
    d, k, st = C.c_void_p(), C.c_void_p(), C.POINTER(B.ScsSettings)()
    lib._scs_read_data.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]
    assert lib._scs_read_data(ours.encode(), C.byref(d), C.byref(k), C.byref(st)) == 0
    theirs = str(tmp_path / "theirs.bin")
    name = theirs.encode()
    st.contents.write_data_filename = name
    lib._scs_write_data.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib._scs_write_data(d, k, st)
    assert open(ours, "rb").read() == open(theirs, "rb").read()
    rd, rK, rstg = scsb.read_data(theirs)
    _same_problem(data, K, stg, rd, rK, rstg)


def test_plain_c_client_builds_and_reports_errors():
    """tools/c/run_from_file.c compiles against include/scs_b200.h with gcc alone; without a device it reads
    the file and then refuses to solve (no CPU fallback); a missing file is an error."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    exe = os.path.join(ROOT, "tools", "c", "_bin", "run_from_file")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["gcc", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tools", "c", "run_from_file.c"), "-L" + os.path.join(ROOT, "scs_python_b200"),
                    "-lscsb200", "-Wl,-rpath,$ORIGIN/../../../scs_python_b200", "-o", exe], check=True)
    r = subprocess.run([exe, "/no/such/file"], capture_output=True, text=True)
    assert r.returncode == 3
    if B.lib.scs_b200_device_count() == 0:
        r = subprocess.run([exe, os.path.join(HERE, "golden", "rw_ref_mixed.bin")], capture_output=True, text=True)
        assert r.returncode == 4 and "no CPU fallback" in r.stderr


@pytest.mark.parametrize("name", ["random_prob", "mpc_bug1", "mpc_bug2", "mpc_bug3"])
def test_upstream_fixture_files(name):
    """The reference's own data files (S/test/problems/, written by older SCS versions with other integer
    widths): this reader must return what the reference's reader returned when tests/golden/kat.json was made
    (make_golden.py read_file_problems).  Runs where /root/reference is mounted (the build container)."""
    path = os.path.join("/root/reference/scs_source/test/problems", name)
    if not os.path.exists(path):
        pytest.skip("/root/reference not on this box")
    from tests import helpers
    rec = [p for p in helpers.golden("kat.json")["file_problems"] if p["name"] == name][0]
    rd, rK, rstg = scsb.read_data(path)
    assert rd["A"].shape == (rec["m"], rec["n"])
    assert np.array_equal(rd["A"].data, np.array(rec["Ax"], float)) and list(rd["A"].indices) == rec["Ai"]
    assert list(rd["A"].indptr) == rec["Ap"]
    assert np.array_equal(rd["b"], np.array(rec["b"], float)) and np.array_equal(rd["c"], np.array(rec["c"], float))
    if rec.get("Px") is not None:
        assert np.array_equal(rd["P"].data, np.array(rec["Px"], float)) and list(rd["P"].indices) == rec["Pi"]
    else:
        assert rd["P"] is None
    kc = rec["cone"]
    for key in ("z", "l", "ep", "ed"):
        assert rK.get(key, 0) == kc.get(key, 0)
    for key in ("q", "s"):
        assert list(rK.get(key, [])) == list(kc.get(key, []))
    assert np.array_equal(np.asarray(rK.get("p", []), float), np.asarray(kc.get("p", []), float))


def test_older_file_with_a_shorter_settings_block(tmp_path):
    """Files of older versions end before the newest settings fields (S/test/problems/random_prob, written by
    3.0.0, is 20 bytes short).  The reference keeps reading, reports the short reads and leaves those fields
    zero (rw.c:60-72, 159-180); its own tests then solve the problem.  Same here -- while a file that ends
    inside the cone or data block stays an error (test_read_errors)."""
    data, K, stg = G.rw_problem()
    path = str(tmp_path / "prob.bin")
    scsb.write_data(path, data, K, **stg)
    raw = open(path, "rb").read()
    open(path, "wb").write(raw[:-20])     # drops acceleration_regularization, _relaxation (f64) and adaptive_scale (i32)
    rd, rK, rstg = scsb.read_data(path)
    assert np.array_equal(rd["b"], data["b"]) and rK["l"] == K["l"]
    assert rstg["acceleration_regularization"] == 0.0 and rstg["acceleration_relaxation"] == 0.0
    assert rstg["adaptive_scale"] is False and rstg["acceleration_type_1"] == 1 and rstg["max_iters"] == stg["max_iters"]
