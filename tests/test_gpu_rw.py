"""GPU tier: the per-iteration CSV trace (SCS(log_data_to_csv), S/src/rw.c:317-476) and the data dump at
scs_init (scs.c:1219-1222) of the device-resident solver, against the reference's own trace of the same
seeded problem (tests/golden/rw_ref_trace.csv, CPU_INDIRECT backend, made by make_golden_rw.py).
Tolerance: every numeric column within 1e-6 * max(1, |value|) of the reference row for the same
iteration (FP64, different summation order in the CG and the reductions); the time column is skipped."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden_rw as G  # noqa: E402

pytestmark = pytest.mark.gpu


def _rows(path):
    lines = open(path).read().splitlines()
    cols = [c for c in lines[0].split(",") if c]
    rows = []
    for ln in lines[1:]:
        rows.append([float(v) for v in ln.split(",") if v != ""])
    return cols, rows


def test_csv_trace_matches_reference(gpu, tmp_path):
    import scs_python_b200 as scsb
    data, K, stg = G.trace_problem()
    out = str(tmp_path / "trace.csv")
    sol = scsb.SCS(data, K, verbose=False, log_csv_filename=out, **stg).solve()
    assert sol["info"]["status_val"] == 1
    cols, rows = _rows(out)
    rcols, rrows = _rows(os.path.join(HERE, "golden", "rw_ref_trace.csv"))
    assert cols == rcols[:62] and all(len(r) == 62 for r in rows) and all(len(r) == 62 for r in rrows)
    iters = sol["info"]["iter"]
    assert [int(r[0]) for r in rows] == list(range(iters)) + [iters]      # one row per iteration + the final row
    assert int(rrows[-1][0]) == iters                                     # same iteration count as the reference
    ours = {int(r[0]): r for r in rows[:-1]}
    ours_final, ref_final = rows[-1], rrows[-1]
    t = cols.index("time")
    checked = 0
    for ref in rrows[:-1] + [ref_final]:
        mine = ours_final if ref is ref_final else ours[int(ref[0])]
        for j, name in enumerate(cols):
            if j == t:
                continue
            a, b = mine[j], ref[j]
            if np.isnan(b):
                assert np.isnan(a), (int(ref[0]), name, a, b)
                continue
            assert abs(a - b) <= 1e-6 * max(1.0, abs(b)), (int(ref[0]), name, a, b)
            checked += 1
    assert checked > 2000
    assert all(rows[k + 1][t] >= rows[k][t] for k in range(len(rows) - 1))   # time runs forward


def test_write_data_filename_at_init(gpu, tmp_path):
    import scs_python_b200 as scsb
    data, K, stg = G.rw_problem()
    a, b = str(tmp_path / "at_init.bin"), str(tmp_path / "direct.bin")
    scsb.SCS(data, K, verbose=False, write_data_filename=a, **stg)
    scsb.write_data(b, data, K, verbose=False, **stg)
    assert open(a, "rb").read() == open(b, "rb").read()
    rd, rK, rstg = scsb.read_data(a)                 # and the file solves like the original
    s1 = scsb.SCS(rd, rK, **dict(rstg, verbose=False)).solve()
    s2 = scsb.SCS(data, K, verbose=False, **stg).solve()
    assert s1["info"]["status_val"] == s2["info"]["status_val"] and s1["info"]["iter"] == s2["info"]["iter"]
    assert np.array_equal(s1["x"], s2["x"])
