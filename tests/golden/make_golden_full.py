#!/usr/bin/env python
"""Golden fixtures at the BASELINE.json sizes: the compiled reference (oracle/_ref, QDLDL direct solver) on
the full-size instances of configs[0] (random cone LP and QP, n=2000, m=6000; eps 1e-9 and 1e-4), configs[2]
(SOCP portfolio, n=50k, 10k second-order cones) and configs[3] (MaxCut SDP, 64 PSD cones of order 200) at eps 1e-6
and 1e-4 with the reference's CPU_INDIRECT backend.  Only scalars are stored (status, iterations, objectives, residuals); the instances are
rebuilt from their seeds by scs_python_b200.problems in the tests.

    python tests/golden/make_golden_full.py            (needs oracle/_ref: make -C oracle ref)
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "scs_omp"))

import scs  # noqa: E402  (the reference)
from scs_python_b200 import problems as P  # noqa: E402

CASES = {
    "cfg1_qp": lambda: P.random_cone_qp(seed=1234, with_P=True),
    "cfg1_lp": lambda: P.random_cone_qp(seed=1234, with_P=False),
    "cfg3_socp": lambda: P.socp_portfolio(seed=0),
    "cfg4_sdp": lambda: P.maxcut_sdp(seed=0),
}


def main():
    only = sys.argv[1:]
    path = os.path.join(HERE, "full_ref.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    out["source"] = "reference bodono/scs-python (SCS %s), QDLDL, compiled by oracle/Makefile" % scs.__version__
    for name, build in CASES.items():
        if only and name not in only:
            continue
        d, K, _ = build()
        rec = dict(n=int(d["A"].shape[1]), m=int(d["A"].shape[0]), nnz=int(d["A"].nnz), runs={})
        # configs[0] to 1e-9; the two big ones to 1e-6 (QDLDL needs about an hour of one core for 1e-9 on configs[2])
        for eps in ((1e-9, 1e-4) if name.startswith("cfg1") else (1e-6, 1e-4)):
            t = time.time()
            # configs[0]: QDLDL.  The two big ones: the reference's CPU_INDIRECT (QDLDL needs more than an hour of one
            # core per run on configs[2]); the record says which
            ls = scs.LinearSolver.QDLDL if name.startswith("cfg1") else scs.LinearSolver.CPU_INDIRECT
            sol = scs.SCS(d, K, verbose=False, eps_abs=eps, eps_rel=eps, max_iters=100000, linear_solver=ls).solve()
            i = sol["info"]
            rec["runs"]["%g" % eps] = dict(linear_solver=ls.value, status=i["status"], status_val=i["status_val"], iter=i["iter"], pobj=i["pobj"],
                                           dobj=i["dobj"], res_pri=i["res_pri"], res_dual=i["res_dual"], gap=i["gap"],
                                           setup_ms=i["setup_time"], solve_ms=i["solve_time"])
            print(name, eps, i["status"], i["iter"], i["pobj"], "%.1f s" % (time.time() - t), flush=True)
        out[name] = rec
        json.dump(out, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
