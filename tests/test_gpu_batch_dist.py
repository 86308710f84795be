"""GPU tier (-m gpu), second file: the batch engine (one CTA per problem, Cfg-5 of BASELINE.json)
and the row-partitioned multi-GPU mode (Cfg-2 at N > 1), both through the C ABI.

Fixtures: tests/golden/batch_ref.json holds the compiled reference's QDLDL and CPU_INDIRECT
outputs (status, iterations, objectives) for 24 MPC QPs of the doc example
(S/docs/src/examples/python/mpc.py:12-65 with the SURVEY 8(d) Cfg-5 sizes) and a handful of
SOCPs, produced by tests/golden/make_golden_batch.py.  Tolerances: status identical; objectives
within 1e-6 relative at eps 1e-9 (north_star), within 2e-3 relative at the default eps 1e-4 where
two correct solvers only agree to the stopping tolerance.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from tests import helpers, problems as tp

pytestmark = pytest.mark.gpu

GOLD = helpers.golden("batch_ref.json")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def scsb(gpu):
    import scs_python_b200
    return scs_python_b200


@pytest.fixture(scope="module")
def mpc_probs():
    from scs_python_b200 import problems as bp
    return [bp.mpc_qp(seed)[:2] for seed in range(len(GOLD["mpc"]))]


def _rel(a, b):
    return abs(a - b) / max(1.0, abs(b))


@pytest.mark.parametrize("eps", [1e-9, 1e-4])
def test_batch_mpc_vs_reference_golden(scsb, mpc_probs, eps):
    sols = scsb.solve_batch(mpc_probs, eps_abs=eps, eps_rel=eps, max_iters=100000, verbose=False)
    assert len(sols) == len(mpc_probs)
    tol = 1e-6 if eps < 1e-8 else 2e-3
    for s, g in zip(sols, GOLD["mpc"]):
        r = g["runs"]["qdldl_%g" % eps]
        assert s["info"]["status_val"] == r["status_val"], (s["info"]["status"], r["status"])
        assert _rel(s["info"]["pobj"], r["pobj"]) < tol, (s["info"]["pobj"], r["pobj"])
        assert _rel(s["info"]["dobj"], r["dobj"]) < tol, (s["info"]["dobj"], r["dobj"])
    if eps == 1e-9:
        # the batch engine factors the reduced system exactly, like QDLDL: the iteration counts of
        # the reference's direct runs are reproduced (one convergence check = 25 iterations slack)
        its = np.array([s["info"]["iter"] for s in sols]); ref = np.array([g["runs"]["qdldl_1e-09"]["iter"] for g in GOLD["mpc"]])
        assert np.all(np.abs(its - ref) <= 25), (its, ref)
    else:
        for (d, k), s in list(zip(mpc_probs, sols))[:8]:
            helpers.verify_solution(d, k, s, eps, eps)     # S/test/problem_utils.h:107-249


@pytest.mark.parametrize("eps", [1e-9, 1e-4])
def test_batch_soc_vs_reference_golden(scsb, eps):
    soc = []
    for g in GOLD["soc"]:
        d, _ = tp.gen_feasible(g["cone"], g["n"], 0.3, g["seed"], with_P=bool(g["with_P"]))
        soc.append((d, g["cone"]))
    sols = scsb.solve_batch(soc, eps_abs=eps, eps_rel=eps, max_iters=100000, verbose=False)
    tol = 1e-6 if eps < 1e-8 else 2e-3
    for s, g in zip(sols, GOLD["soc"]):
        r = g["runs"]["qdldl_%g" % eps]
        assert s["info"]["status_val"] == r["status_val"]
        assert _rel(s["info"]["pobj"], r["pobj"]) < tol


@pytest.mark.parametrize("eps", [1e-9, 1e-4])
def test_batch_box_cone_vs_reference_golden(scsb, eps):
    """The MPC problems in the reference example's own form, bounds as ONE box cone with the t row pinned by b
    (S/docs/src/examples/python/mpc.py:49-60): the batch kernel runs the box projection (S/src/cones.c:1290-1378,
    Newton on t) itself -- every member must take the fused one-CTA path, not the streaming fallback."""
    from scs_python_b200 import problems as bp
    from scs_python_b200 import _scs_b200 as B
    probs = [bp.mpc_qp_box(g["seed"])[:2] for g in GOLD["mpc_box"]]
    sols = scsb.solve_batch(probs, eps_abs=eps, eps_rel=eps, max_iters=100000, verbose=False)
    st = B.batch_stats()
    assert st["fused"] == len(probs) and st["streamed"] == 0, st
    tol = 1e-6 if eps < 1e-8 else 2e-3
    for s, g, (d, k) in zip(sols, GOLD["mpc_box"], probs):
        r = g["runs"]["qdldl_%g" % eps]
        assert s["info"]["status_val"] == r["status_val"] == 1, (s["info"]["status"], r["status"])
        assert _rel(s["info"]["pobj"], r["pobj"]) < tol, (s["info"]["pobj"], r["pobj"])
        assert _rel(s["info"]["dobj"], r["dobj"]) < tol, (s["info"]["dobj"], r["dobj"])
        helpers.verify_solution(d, k, s, max(eps, 1e-8), max(eps, 1e-8))
    if eps == 1e-9:
        its = np.array([s["info"]["iter"] for s in sols]); ref = np.array([g["runs"]["qdldl_1e-09"]["iter"] for g in GOLD["mpc_box"]])
        print("box-cone MPC iterations b200 %s / reference QDLDL %s" % (its.tolist(), ref.tolist()))
        assert np.all(np.abs(its - ref) <= 50), (its, ref)
    # same members through the streaming engine: same optimum (two engines, one algorithm)
    for (d, k), s in list(zip(probs, sols))[:2]:
        a = scsb.SCS(d, k, eps_abs=eps, eps_rel=eps, max_iters=100000, verbose=False).solve(warm_start=False)
        assert a["info"]["status_val"] == 1 and _rel(a["info"]["pobj"], s["info"]["pobj"]) < tol


@pytest.mark.parametrize("eps", [1e-9, 1e-4])
def test_batch_exp_and_power_cones_vs_reference_golden(scsb, eps):
    """Exponential (primal / dual) and power (primal / dual) cones inside the one-CTA kernel (one thread per cone,
    the device functions the streaming engine uses: exp_cone.c, cones.c:1276-1324): every member on the fused
    path; status, objectives and iteration counts against the compiled reference's QDLDL runs, residual /
    cone-membership criteria recomputed on the host."""
    from scs_python_b200 import _scs_b200 as B
    probs = []
    for g in GOLD["tri"]:
        d, _ = tp.gen_feasible(g["cone"], g["n"], 0.3, g["seed"], with_P=bool(g["with_P"]))
        probs.append((d, g["cone"]))
    sols = scsb.solve_batch(probs, eps_abs=eps, eps_rel=eps, max_iters=100000, verbose=False)
    st = B.batch_stats()
    assert st["fused"] == len(probs) and st["streamed"] == 0, st
    tol = 1e-6 if eps < 1e-8 else 2e-3
    its, ref = [], []
    for s, g, (d, k) in zip(sols, GOLD["tri"], probs):
        r = g["runs"]["qdldl_%g" % eps]
        assert s["info"]["status_val"] == r["status_val"] == 1, (g["seed"], s["info"]["status"], r["status"])
        assert _rel(s["info"]["pobj"], r["pobj"]) < tol, (g["seed"], s["info"]["pobj"], r["pobj"])
        assert _rel(s["info"]["dobj"], r["dobj"]) < tol, (g["seed"], s["info"]["dobj"], r["dobj"])
        helpers.verify_solution(d, k, s, max(eps, 1e-8), max(eps, 1e-8))
        its.append(s["info"]["iter"]); ref.append(r["iter"])
    print("exp / power batch members, eps %g: iterations b200 %s / reference QDLDL %s" % (eps, its, ref))
    if eps == 1e-9:
        assert np.all(np.abs(np.array(its) - np.array(ref)) <= 50), (its, ref)
    # a mixed batch: members with and without three-row cones share one launch
    from scs_python_b200 import problems as bp
    mixed = [probs[0], bp.mpc_qp(0)[:2], probs[2], bp.mpc_qp_box(1)[:2]]
    ms = scsb.solve_batch(mixed, eps_abs=eps, eps_rel=eps, max_iters=100000, verbose=False)
    assert B.batch_stats()["fused"] == 4
    assert all(m["info"]["status_val"] == 1 for m in ms)
    for a, b in ((ms[0], sols[0]), (ms[2], sols[2])):
        assert _rel(a["info"]["pobj"], b["info"]["pobj"]) < tol


@pytest.mark.parametrize("section", ["psd", "cpsd"])
@pytest.mark.parametrize("eps", [1e-9, 1e-4])
def test_batch_small_psd_cones_vs_reference_golden(scsb, eps, section):
    """Positive-semidefinite cones of order 1..16 inside the one-CTA kernel (one warp per cone, two-sided Jacobi in
    shared memory; cones.c:991-1148) and complex ones of order 1..16 through their real embedding: every member on the fused path; status, objectives and iteration counts against
    the compiled reference's QDLDL runs (which project with LAPACK dsyevr), host-side residual / cone checks, and
    the same members through the streaming engine (blocked one-sided Jacobi kernels of cones.cu)."""
    from scs_python_b200 import _scs_b200 as B
    probs = []
    for g in GOLD[section]:
        d, _ = tp.gen_feasible(g["cone"], g["n"], 0.3, g["seed"], with_P=bool(g["with_P"]))
        probs.append((d, g["cone"]))
    sols = scsb.solve_batch(probs, eps_abs=eps, eps_rel=eps, max_iters=100000, verbose=False)
    st = B.batch_stats()
    assert st["fused"] == len(probs) and st["streamed"] == 0, st
    tol = 1e-6 if eps < 1e-8 else 2e-3
    its, ref = [], []
    for s, g, (d, k) in zip(sols, GOLD[section], probs):
        r = g["runs"]["qdldl_%g" % eps]
        assert s["info"]["status_val"] == r["status_val"] == 1, (g["seed"], s["info"]["status"], r["status"])
        assert _rel(s["info"]["pobj"], r["pobj"]) < tol, (g["seed"], s["info"]["pobj"], r["pobj"])
        assert _rel(s["info"]["dobj"], r["dobj"]) < tol, (g["seed"], s["info"]["dobj"], r["dobj"])
        helpers.verify_solution(d, k, s, max(eps, 1e-8), max(eps, 1e-8))
        its.append(s["info"]["iter"]); ref.append(r["iter"])
    print(section + " batch members, eps %g: iterations b200 %s / reference QDLDL %s" % (eps, its, ref))
    if eps == 1e-9:
        assert np.all(np.abs(np.array(its) - np.array(ref)) <= 50), (its, ref)
    for (d, k), s in list(zip(probs, sols))[:2]:
        a = scsb.SCS(d, k, eps_abs=eps, eps_rel=eps, max_iters=100000, verbose=False).solve(warm_start=False)
        assert a["info"]["status_val"] == 1 and _rel(a["info"]["pobj"], s["info"]["pobj"]) < tol
    # a cone above the in-kernel order limit (real 32, complex 16) sends its member to the streaming engine, the rest stay fused
    K = dict(z=1, l=2, s=[34]) if section == "psd" else dict(z=1, l=2, cs=[17])
    d, _ = tp.gen_feasible(K, 12, 0.3, 31, with_P=False)
    ms = scsb.solve_batch([probs[0], (d, K)], eps_abs=1e-6, eps_rel=1e-6, max_iters=100000, verbose=False)
    st = B.batch_stats()
    assert st["fused"] == 1 and st["streamed"] == 1, st
    assert all(m["info"]["status_val"] == 1 for m in ms)


def test_batch_agrees_with_streaming_engine(scsb, mpc_probs):
    """Same problem through the one-CTA batch kernel and through the streaming (graph-launched)
    engine: same status, objectives within 1e-6 relative, iterates within 1e-5."""
    kw = dict(eps_abs=1e-9, eps_rel=1e-9, max_iters=100000, verbose=False)
    for d, k in mpc_probs[:2]:
        a = scsb.SCS(d, k, **kw).solve(warm_start=False)
        b = scsb.solve_batch([(d, k)], **kw)[0]
        assert a["info"]["status_val"] == b["info"]["status_val"] == 1
        assert _rel(a["info"]["pobj"], b["info"]["pobj"]) < 1e-6
        assert np.max(np.abs(a["x"] - b["x"])) < 1e-5 * max(1.0, np.max(np.abs(a["x"])))


@pytest.mark.parametrize("kw", [dict(), dict(normalize=False), dict(acceleration_lookback=0),
                                dict(acceleration_type_1=False), dict(adaptive_scale=False),
                                dict(acceleration_lookback=5, acceleration_interval=3),
                                dict(scale=5.0, rho_x=1e-3, alpha=1.8)])
def test_batch_settings_variants(scsb, kw):
    """every ScsSettings field the iteration reads (S/include/scs.h:64-119) takes effect in the
    batch kernel the same way as in the streaming engine"""
    rng = np.random.RandomState(0)
    K = dict(z=3, l=12, q=[4, 6])
    d, _ = tp.gen_feasible(K, 18, 0.4, 9, with_P=False)
    Q = sp.random(18, 18, density=0.2, random_state=rng, data_rvs=rng.randn)
    d["P"] = sp.csc_matrix(sp.triu(Q @ Q.T + 0.1 * sp.eye(18)))
    base = dict(eps_abs=1e-9, eps_rel=1e-9, max_iters=100000, verbose=False)
    a = scsb.SCS(d, K, **base, **kw).solve(warm_start=False)
    b = scsb.solve_batch([(d, K)], **base, **kw)[0]
    assert a["info"]["status_val"] == b["info"]["status_val"] == 1
    assert _rel(a["info"]["pobj"], b["info"]["pobj"]) < 1e-6


def test_batch_certificates(scsb):
    """infeasible / unbounded members of a batch return the reference's status codes (-2 / -1)"""
    probs, want = [], []
    for rec in helpers.golden("ref_runs.json")["solves"]["cases"]:
        if rec["name"] in ("infeasible", "unbounded"):
            probs.append(helpers.problem_from_record(rec)); want.append(rec["runs"]["qdldl_1e-07"]["status_val"])
    assert probs
    sols = scsb.solve_batch(probs, eps_abs=1e-7, eps_rel=1e-7, verbose=False)
    assert [s["info"]["status_val"] for s in sols] == want


def test_batch_heterogeneous_and_empty(scsb, mpc_probs):
    """ragged batch (different n, m, cones per member), batch of one, empty batch"""
    assert scsb.solve_batch([], verbose=False) == []
    K = dict(z=2, l=5, q=[3])
    d, _ = tp.gen_feasible(K, 6, 0.5, 3, with_P=True)
    mixed = [mpc_probs[0], (d, K), mpc_probs[1]]
    sols = scsb.solve_batch(mixed, eps_abs=1e-8, eps_rel=1e-8, max_iters=100000, verbose=False)
    one = [scsb.solve_batch([p], eps_abs=1e-8, eps_rel=1e-8, max_iters=100000, verbose=False)[0] for p in mixed]
    for s, o, (dd, kk) in zip(sols, one, mixed):
        assert s["x"].shape == (dd["A"].shape[1],) and s["y"].shape == (dd["A"].shape[0],)
        assert s["info"]["status_val"] == 1
        # a member's result does not depend on what else is in the batch (bit-exact)
        assert np.array_equal(s["x"], o["x"]) and s["info"]["iter"] == o["info"]["iter"]


def test_batch_large_count_properties(scsb):
    """Cfg-5 scale on one GPU's share (1024 problems): all solved, residual criteria re-checked on
    the host for a sample, iteration counts in the range the reference shows for this family"""
    from scs_python_b200 import problems as bp
    probs = [bp.mpc_qp(1000 + i)[:2] for i in range(1024)]
    sols = scsb.solve_batch(probs, verbose=False)
    assert all(s["info"]["status_val"] == 1 for s in sols)
    for i in range(0, 1024, 97):
        helpers.verify_solution(probs[i][0], probs[i][1], sols[i], 1e-4, 1e-4)
