#!/usr/bin/env python
"""Generate tests/golden/batch_ref.json: outputs of the compiled reference (oracle/_ref, built from
/root/reference by `make -C oracle ref`) on the seeded small problems the batch engine is checked on.
Runs ONLY in the build container; the fixture is committed.

    make -C oracle ref && python tests/golden/make_golden_batch.py

Cases: BASELINE.json configs[4] MPC QPs (scs_python_b200.problems.mpc_qp, seeds 0..23), the same problems
with a box cone (problems.mpc_qp_box, seeds 0..11), small programs with exponential / power cones, with small PSD cones, and small
second-order-cone programs (tests/problems.gen_feasible), each solved by scs.SCS(...).solve() with
QDLDL and CPU_INDIRECT at eps 1e-4 (defaults) and 1e-9.
"""
from __future__ import annotations

import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))

import numpy as np  # noqa: E402


def rec_of(sol):
    i = sol["info"]
    f = lambda v: ("nan" if v != v else ("inf" if v == float("inf") else ("-inf" if v == float("-inf") else float(v))))
    return dict(status=i["status"], status_val=int(i["status_val"]), iter=int(i["iter"]), pobj=f(i["pobj"]), dobj=f(i["dobj"]),
                res_pri=f(i["res_pri"]), res_dual=f(i["res_dual"]), gap=f(i["gap"]), scale_updates=int(i["scale_updates"]))


def main():
    if not os.path.isdir("/root/reference"):
        sys.exit("make_golden_batch.py needs /root/reference (build container only)")
    import scs
    from scs_python_b200 import problems as bp
    from tests import problems as tp
    out = dict(source="scs.SCS(...).solve() of oracle/_ref (SCS 3.2.11), QDLDL and CPU_INDIRECT", mpc=[], soc=[], mpc_box=[], tri=[], psd=[], cpsd=[])
    for seed in range(24):
        data, cone, _ = bp.mpc_qp(seed)
        rec = dict(seed=seed, runs={})
        for eps in (1e-4, 1e-9):
            for name, ls in (("qdldl", scs.LinearSolver.QDLDL), ("cpu_indirect", scs.LinearSolver.CPU_INDIRECT)):
                sol = scs.SCS(data, cone, linear_solver=ls, verbose=False, eps_abs=eps, eps_rel=eps, max_iters=100000).solve()
                rec["runs"]["%s_%g" % (name, eps)] = rec_of(sol)
        out["mpc"].append(rec)
    for seed in range(12):  # the same MPC problems with the bounds as one box cone (the reference example's own form)
        data, cone, _ = bp.mpc_qp_box(seed)
        rec = dict(seed=seed, runs={})
        for eps in (1e-4, 1e-9):
            for name, ls in (("qdldl", scs.LinearSolver.QDLDL), ("cpu_indirect", scs.LinearSolver.CPU_INDIRECT)):
                sol = scs.SCS(data, cone, linear_solver=ls, verbose=False, eps_abs=eps, eps_rel=eps, max_iters=100000).solve()
                rec["runs"]["%s_%g" % (name, eps)] = rec_of(sol)
        out["mpc_box"].append(rec)
    for seed, K, n, withP in [(3, dict(z=4, l=10, q=[3, 5, 8]), 20, True), (4, dict(z=0, l=30, q=[4] * 10 + [1, 2]), 35, False),
                              (5, dict(z=6, l=0, q=[12, 40]), 30, True), (6, dict(z=2, l=50), 25, True)]:
        data, p_star = tp.gen_feasible(K, n, 0.3, seed, with_P=withP)
        rec = dict(seed=seed, cone=K, n=n, with_P=withP, p_star=p_star, runs={})
        for eps in (1e-4, 1e-9):
            for name, ls in (("qdldl", scs.LinearSolver.QDLDL), ("cpu_indirect", scs.LinearSolver.CPU_INDIRECT)):
                sol = scs.SCS(data, K, linear_solver=ls, verbose=False, eps_abs=eps, eps_rel=eps, max_iters=100000).solve()
                rec["runs"]["%s_%g" % (name, eps)] = rec_of(sol)
        out["soc"].append(rec)
    # exponential / power cones next to z, l, q (three-row cones, one thread each in the batch kernel)
    for seed, K, n, withP in [(11, dict(z=3, l=8, q=[4], ep=3, ed=2, p=[0.3, -0.6, 0.5]), 20, True),
                              (12, dict(z=0, l=10, ep=6), 15, False),
                              (13, dict(z=2, l=5, q=[3, 6], p=[0.25, -0.75, 0.5, 0.9]), 18, True),
                              (14, dict(z=2, l=6, ed=5), 14, True),
                              (15, dict(z=1, l=4, ep=2, ed=2, p=[-0.4, 0.7]), 12, False)]:
        data, p_star = tp.gen_feasible(K, n, 0.3, seed, with_P=withP)
        rec = dict(seed=seed, cone=K, n=n, with_P=withP, p_star=p_star, runs={})
        for eps in (1e-4, 1e-9):
            for name, ls in (("qdldl", scs.LinearSolver.QDLDL), ("cpu_indirect", scs.LinearSolver.CPU_INDIRECT)):
                sol = scs.SCS(data, K, linear_solver=ls, verbose=False, eps_abs=eps, eps_rel=eps, max_iters=100000).solve()
                rec["runs"]["%s_%g" % (name, eps)] = rec_of(sol)
        out["tri"].append(rec)
    # small positive-semidefinite cones (one warp each in the batch kernel) next to the other cones
    for seed, K, n, withP in [(21, dict(z=2, l=4, q=[3], s=[3, 5]), 14, True),
                              (22, dict(z=0, l=3, s=[8], ep=2), 12, False),
                              (23, dict(z=1, l=2, s=[1, 2, 4, 6], p=[0.4]), 16, True),
                              (24, dict(z=3, l=0, s=[12]), 20, True),
                              (25, dict(z=0, l=5, s=[16, 2]), 25, False)]:
        data, p_star = tp.gen_feasible(K, n, 0.3, seed, with_P=withP)
        rec = dict(seed=seed, cone=K, n=n, with_P=withP, p_star=p_star, runs={})
        for eps in (1e-4, 1e-9):
            for name, ls in (("qdldl", scs.LinearSolver.QDLDL), ("cpu_indirect", scs.LinearSolver.CPU_INDIRECT)):
                sol = scs.SCS(data, K, linear_solver=ls, verbose=False, eps_abs=eps, eps_rel=eps, max_iters=100000).solve()
                rec["runs"]["%s_%g" % (name, eps)] = rec_of(sol)
        out["psd"].append(rec)
    # complex positive-semidefinite cones (real embedding of order 2 cs in the batch kernel)
    for seed, K, n, withP in [(41, dict(z=2, l=3, cs=[3]), 12, True),
                              (42, dict(z=0, l=4, q=[3], s=[3], cs=[2, 5], ep=1), 18, False),
                              (43, dict(z=1, l=2, cs=[1, 8]), 20, True),
                              (44, dict(z=0, l=3, cs=[16]), 30, False)]:
        data, p_star = tp.gen_feasible(K, n, 0.3, seed, with_P=withP)
        rec = dict(seed=seed, cone=K, n=n, with_P=withP, p_star=p_star, runs={})
        for eps in (1e-4, 1e-9):
            for name, ls in (("qdldl", scs.LinearSolver.QDLDL), ("cpu_indirect", scs.LinearSolver.CPU_INDIRECT)):
                sol = scs.SCS(data, K, linear_solver=ls, verbose=False, eps_abs=eps, eps_rel=eps, max_iters=100000).solve()
                rec["runs"]["%s_%g" % (name, eps)] = rec_of(sol)
        out["cpsd"].append(rec)
    json.dump(out, open(os.path.join(HERE, "batch_ref.json"), "w"), indent=0)
    its = [r["runs"]["cpu_indirect_0.0001"]["iter"] for r in out["mpc"]]
    print("batch_ref.json: %d mpc, %d soc; mpc iters at 1e-4 (indirect): min %d median %d max %d" %
          (len(out["mpc"]), len(out["soc"]), min(its), int(np.median(its)), max(its)))
    print([(r["runs"]["qdldl_0.0001"]["iter"], r["runs"]["cpu_indirect_1e-09"]["iter"], r["runs"]["cpu_indirect_1e-09"]["status"]) for r in out["mpc"]][:8])


if __name__ == "__main__":
    main()
