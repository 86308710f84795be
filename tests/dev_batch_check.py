"""Dev check of the batch engine on a GPU box (not collected by pytest): parity against the committed
reference outputs, the streaming engine and the oracle, then a throughput sample."""
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scs_python_b200 as scsb  # noqa: E402
from scs_python_b200 import _scs_b200 as B  # noqa: E402
from scs_python_b200 import problems as bp  # noqa: E402
from tests import helpers, problems as tp  # noqa: E402

gold = helpers.golden("batch_ref.json")
REFKEY = "cpu_indirect" if os.environ.get("SCS_B200_BATCH_DIRECT") == "0" else "qdldl"
bad = 0


def rel(a, b):
    return abs(a - b) / max(1.0, abs(b))


def report(tag, sols, refs, eps):
    global bad
    for i, (s, r) in enumerate(zip(sols, refs)):
        si = s["info"]
        ok = si["status_val"] == r["status_val"] and rel(si["pobj"], r["pobj"]) < (1e-6 if eps < 1e-8 else 2e-3)
        if not ok:
            bad += 1
        if not ok or i < 4:
            print("%s[%d] %s status %s/%s iter %d/%d pobj %.9g/%.9g dobj %.9g/%.9g cg %s" % (
                tag, i, "ok " if ok else "BAD", si["status"], r["status"], si["iter"], r["iter"], si["pobj"], r["pobj"],
                si["dobj"], r["dobj"], si.get("scale_updates")))


probs = [bp.mpc_qp(seed)[:2] for seed in range(24)]
for eps in (() if "--perf-only" in sys.argv else (1e-9, 1e-4)):
    t = time.perf_counter()
    sols = scsb.solve_batch(probs, eps_abs=eps, eps_rel=eps, max_iters=100000, verbose=False)
    dt = time.perf_counter() - t
    print("mpc batch eps %g: %.1f ms wall, stats %s" % (eps, dt * 1e3, B.batch_stats()))
    report("mpc%g" % eps, sols, [g["runs"]["%s_%g" % (REFKEY, eps)] for g in gold["mpc"]], eps)
    if eps == 1e-4:
        for (d, k), s in list(zip(probs, sols))[:6]:
            try:
                helpers.verify_solution(d, k, s, eps, eps)
            except AssertionError as e:
                bad += 1
                print("verify_solution failed:", e)

# streaming engine on the same problems
for i in range(0 if '--perf-only' in sys.argv else 3):
    d, k = probs[i]
    a = scsb.SCS(d, k, eps_abs=1e-9, eps_rel=1e-9, max_iters=100000, verbose=False).solve(warm_start=False)
    b = scsb.solve_batch([probs[i]], eps_abs=1e-9, eps_rel=1e-9, max_iters=100000, verbose=False)[0]
    print("stream vs batch [%d]: iter %d/%d pobj %.10g/%.10g |dx| %.2e |dy| %.2e" % (
        i, a["info"]["iter"], b["info"]["iter"], a["info"]["pobj"], b["info"]["pobj"],
        np.max(np.abs(a["x"] - b["x"])), np.max(np.abs(a["y"] - b["y"]))))
    if rel(a["info"]["pobj"], b["info"]["pobj"]) > 1e-6:
        bad += 1

# SOC cases + heterogeneous batch
soc = []
for g in gold["soc"]:
    d, _ = tp.gen_feasible(g["cone"], g["n"], 0.3, g["seed"], with_P=bool(g["with_P"]))
    soc.append((d, g["cone"]))
for eps in (() if '--perf-only' in sys.argv else (1e-9, 1e-4)):
    sols = scsb.solve_batch(soc, eps_abs=eps, eps_rel=eps, max_iters=100000, verbose=False)
    print("soc batch eps %g stats %s" % (eps, B.batch_stats()))
    report("soc%g" % eps, sols, [g["runs"]["%s_%g" % (REFKEY, eps)] for g in gold["soc"]], eps)

# dense-ish P, settings variants, against the streaming engine
rng = np.random.RandomState(0)
K = dict(z=3, l=12, q=[4, 6])
d, _ = tp.gen_feasible(K, 18, 0.4, 9, with_P=False)
Q = sp.random(18, 18, density=0.2, random_state=rng, data_rvs=rng.randn)
d["P"] = sp.csc_matrix(sp.triu(Q @ Q.T + 0.1 * sp.eye(18)))
for kw in () if '--perf-only' in sys.argv else (dict(), dict(normalize=False), dict(acceleration_lookback=0), dict(acceleration_type_1=False),
           dict(adaptive_scale=False), dict(acceleration_lookback=5, acceleration_interval=3), dict(scale=5.0, rho_x=1e-3, alpha=1.8)):
    a = scsb.SCS(d, K, eps_abs=1e-9, eps_rel=1e-9, max_iters=100000, verbose=False, **kw).solve(warm_start=False)
    b = scsb.solve_batch([(d, K)], eps_abs=1e-9, eps_rel=1e-9, max_iters=100000, verbose=False, **kw)[0]
    ok = a["info"]["status_val"] == b["info"]["status_val"] and rel(a["info"]["pobj"], b["info"]["pobj"]) < 1e-6
    bad += 0 if ok else 1
    print("variant %s: %s status %s/%s iter %d/%d pobj %.10g/%.10g acc %d/%d rej %d/%d" % (
        kw, "ok " if ok else "BAD", a["info"]["status"], b["info"]["status"], a["info"]["iter"], b["info"]["iter"],
        a["info"]["pobj"], b["info"]["pobj"], a["info"]["accepted_accel_steps"], b["info"]["accepted_accel_steps"],
        a["info"]["rejected_accel_steps"], b["info"]["rejected_accel_steps"]))

# infeasible / unbounded records of ref_runs.json
for rec in ([] if '--perf-only' in sys.argv else helpers.golden("ref_runs.json")["solves"]["cases"]):
    if rec["name"] in ("infeasible", "unbounded"):
        d, k = helpers.problem_from_record(rec)
        b = scsb.solve_batch([(d, k)], eps_abs=1e-7, eps_rel=1e-7, verbose=False)[0]
        r = rec["runs"]["qdldl_1e-07"]
        ok = b["info"]["status_val"] == r["status_val"]
        bad += 0 if ok else 1
        print("%s: %s status %s/%s iter %d/%d" % (rec["name"], "ok " if ok else "BAD", b["info"]["status"], r["status"],
                                                  b["info"]["iter"], r["iter"]))

# throughput
if "--perf" in sys.argv:
    for cnt in (1024, 8192):
        big = [bp.mpc_qp(seed)[:2] for seed in range(cnt)]
        prepared = [scsb._prepare(d, k) for d, k in big]
        for rep in range(2):
            t = time.perf_counter()
            sols = B.solve_batch(prepared, verbose=False)
            dt = time.perf_counter() - t
            st = B.batch_stats()
            its = np.array([s["info"]["iter"] for s in sols])
            nsolved = sum(s["info"]["status_val"] == 1 for s in sols)
            print("perf cnt %d rep %d: wall %.1f ms kernel %.1f ms pack %.1f ms solved %d iters mean %.0f max %d  -> %.0f problems/s (kernel), "
                  "%.3g ADMM it/s (kernel) direct %d cg/it %.2f smem %d ctas %d" % (
                      cnt, rep, dt * 1e3, st["kernel_ms"], st["pack_ms"], nsolved, its.mean(), its.max(),
                      cnt / st["kernel_ms"] * 1e3, its.sum() / st["kernel_ms"] * 1e3, st["direct"],
                      st["cg_iters"] / max(1.0, st["admm_iters"]), st["smem_per_cta"], st["ctas"]))
            if rep == 1:
                tot = max(1.0, st["clk_total"])
                print("   cycle shares: equil %.3f factor %.3f linsys %.3f aa %.3f resid %.3f ; cycles/iter %.0f" % (
                    st["clk_equil"] / tot, st["clk_factor"] / tot, st["clk_linsys"] / tot, st["clk_aa"] / tot,
                    st["clk_resid"] / tot, tot / max(1.0, st["admm_iters"])))
print("BAD =", bad)
sys.exit(1 if bad else 0)
