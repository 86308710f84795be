#!/usr/bin/env python
"""Row-partitioned code path on ONE GPU (world = 1, SCS_B200_DIST_SELFTEST=1): column classification
(shared / private), deferred reduction finalisers, split products (A_g'z all-reduced, P p + R_x p added
after), the piggy-backed p'Gp, Anderson acceleration over gathered trapezoids -- with the collectives
degenerated to copies.  Every case is checked against the compiled reference (oracle/_ref, QDLDL) when it
travelled to the box, else against the numpy oracle, and against this library's ordinary single-GPU path.

    python tests/dist_selftest.py          (prints one JSON line per case, "dist selftest ok" at the end)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["SCS_B200_DIST_SELFTEST"] = "1"

import scs_python_b200 as scsb  # noqa: E402
from scs_python_b200 import _scs_b200 as B  # noqa: E402
from scs_python_b200 import problems as P  # noqa: E402


def reference_solver():
    p = os.path.join(ROOT, "oracle", "_ref")
    if os.path.exists(os.path.join(p, "scs", "__init__.py")):
        sys.path.insert(0, p)
        try:
            import scs
            return lambda d, K, kw: scs.SCS(d, K, verbose=False, **kw).solve(), "reference QDLDL (oracle/_ref)"
        except Exception:
            sys.path.pop(0)
    from oracle import scs_oracle as O
    return lambda d, K, kw: O.ScsOracle(d, K, **kw).solve(), "numpy oracle"


def rel(a, b):
    return abs(a - b) / max(1.0, abs(b))


def main():
    assert B.lib.scs_b200_device_count() > 0, "no CUDA device"
    B.lib.scs_b200_set_device(0)
    ref_solve, ref_kind = reference_solver()
    cases = []
    d, K, _ = P.random_cone_qp(seed=7, n=300, l=400, nq=40, q=6, ep=30, density=0.04)
    cases.append(("cone_qp", d, K, dict(eps_abs=1e-9, eps_rel=1e-9, max_iters=100000), {}))
    d, K, _ = P.lasso(3000, 6000, 20, seed=2)
    cases.append(("lasso", d, K, dict(eps_abs=1e-8, eps_rel=1e-8, eps_infeas=1e-13, max_iters=100000), {}))
    cases.append(("lasso_tiled", d, K, dict(eps_abs=1e-8, eps_rel=1e-8, eps_infeas=1e-13, max_iters=100000),
                  {"SCS_B200_TILED": "1"}))
    d, K, _ = P.socp_portfolio(seed=1, n=1500, ncones=300)
    cases.append(("socp", d, K, dict(eps_abs=1e-8, eps_rel=1e-8, max_iters=100000), {}))
    d, K, _ = P.maxcut_sdp(seed=1, nodes=20, blocks=4)
    cases.append(("sdp", d, K, dict(eps_abs=1e-8, eps_rel=1e-8, max_iters=100000), {}))
    single, refs = {}, {}
    for name, d, K, kw, env in cases:
        os.environ.update(env)
        single[name] = scsb.SCS(d, K, verbose=False, **kw).solve()
        for k in env:
            del os.environ[k]
        if name != "lasso_tiled":
            refs[name] = ref_solve(d, K, kw)
    refs["lasso_tiled"] = refs["lasso"]
    scsb.dist_init(0, 1)  # world 1 + SCS_B200_DIST_SELFTEST=1: the partitioned code path on this GPU
    ok = True
    for force in ("0", "3"):
        os.environ["SCS_B200_DIST_FORCE_SHARED"] = force
        for name, d, K, kw, env in cases:
            os.environ.update(env)
            s = scsb.SCS(d, K, verbose=False, **kw)
            r = s.solve()
            for k in env:
                del os.environ[k]
            a, g, b = single[name]["info"], refs[name]["info"], r["info"]
            errs = dict(pobj_ref=rel(b["pobj"], g["pobj"]), dobj_ref=rel(b["dobj"], g["dobj"]),
                        x_single=float(np.max(np.abs(single[name]["x"] - r["x"])) / max(1.0, np.max(np.abs(r["x"])))),
                        y_single=float(np.max(np.abs(single[name]["y"] - r["y"])) / max(1.0, np.max(np.abs(r["y"])))),
                        s_single=float(np.max(np.abs(single[name]["s"] - r["s"])) / max(1.0, np.max(np.abs(r["s"])))))
            # north_star: same status, objectives within 1e-6 relative of the reference's
            good = (b["status_val"] == g["status_val"] == 1 and errs["pobj_ref"] < 1e-6 and errs["dobj_ref"] < 1e-6
                    and errs["x_single"] < 1e-5 and errs["y_single"] < 1e-5 and errs["s_single"] < 1e-5)
            ok = ok and good
            print(json.dumps(dict(case=name, force_shared=force, ok=bool(good), ref=ref_kind,
                                  status=(b["status"], g["status"]), iters=(b["iter"], a["iter"], g["iter"]),
                                  errs=errs, accepted_aa=b["accepted_accel_steps"])), flush=True)
            # warm start from the solution + update(b, c): converges at the first check, same objective
            r2 = s.solve(warm_start=True, x=r["x"], y=r["y"], s=r["s"])
            good2 = r2["info"]["status_val"] == 1 and r2["info"]["iter"] <= 50 and rel(r2["info"]["pobj"], g["pobj"]) < 1e-6
            s.update(b=np.asarray(d["b"]) * 1.0, c=np.asarray(d["c"]) * 1.0)
            r3 = s.solve(warm_start=False)
            good3 = r3["info"]["status_val"] == 1 and rel(r3["info"]["pobj"], g["pobj"]) < 1e-6
            ok = ok and good2 and good3
            if not (good2 and good3):
                print(json.dumps(dict(case=name, warm=bool(good2), update=bool(good3), it=r2["info"]["iter"])), flush=True)
    # verbose run that stops at max_iters (the collective sequence must not depend on who prints)
    d, K, _ = P.lasso(1000, 2000, 10, seed=3)
    r = scsb.SCS(d, K, verbose=True, max_iters=60, eps_abs=1e-14, eps_rel=1e-14).solve()
    ok = ok and r["info"]["iter"] == 60 and r["info"]["status_val"] == 2
    scsb.dist_finalize()
    if not ok:
        sys.exit(1)
    print("dist selftest ok")


if __name__ == "__main__":
    main()
