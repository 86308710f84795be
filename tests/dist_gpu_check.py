#!/usr/bin/env python
"""Row-partitioned solve (>= 2 GPUs, NCCL) against the compiled reference / oracle and the single-GPU solve:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29655 tests/dist_gpu_check.py

Every rank first solves each problem alone on its GPU, rank 0 also solves it with the compiled reference
(oracle/_ref, QDLDL) or, when that did not travel, the numpy oracle; then the ranks solve it together with
the default settings (Anderson acceleration on).  Same status; objectives within 1e-6 relative of the
reference's (north_star); x, y, s agree with the single-GPU solve to 1e-5 (different summation order, so the
iterates are not bit-equal and iteration counts may differ by a convergence check).  Finally a verbose run
that stops at max_iters: every rank must issue the same collectives whoever prints.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as td  # noqa: E402

import scs_python_b200 as scsb  # noqa: E402
from scs_python_b200 import _scs_b200 as B  # noqa: E402
from scs_python_b200 import problems as P  # noqa: E402


def reference_solver():
    p = os.path.join(ROOT, "oracle", "_ref")
    if os.path.exists(os.path.join(p, "scs", "__init__.py")):
        sys.path.insert(0, p)
        try:
            import scs
            # QDLDL, except where its fill-in would explode (the random sparse LASSO matrix): CPU_INDIRECT there
            def solve(d, K, kw):
                ls = scs.LinearSolver.CPU_INDIRECT if d["A"].nnz > 300_000 else scs.LinearSolver.QDLDL
                return scs.SCS(d, K, verbose=False, linear_solver=ls, **kw).solve()
            return solve, "reference QDLDL / CPU_INDIRECT (oracle/_ref)"
        except Exception:
            sys.path.pop(0)
    from oracle import scs_oracle as O
    return lambda d, K, kw: O.ScsOracle(d, K, **kw).solve(), "numpy oracle"


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    td.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    assert B.lib.scs_b200_set_device(local) == 0
    cases = []
    d, K, aux = P.random_cone_qp(seed=7, n=400, l=600, nq=60, q=6, ep=40, density=0.03)
    cases.append(("cone_qp", d, K, dict(eps_abs=1e-9, eps_rel=1e-9, max_iters=100000)))
    d, K, aux = P.lasso(20000, 40000, 50, seed=2)
    cases.append(("lasso", d, K, dict(eps_abs=1e-8, eps_rel=1e-8, eps_infeas=1e-13, max_iters=100000)))
    d, K, aux = P.socp_portfolio(seed=1, n=3000, ncones=600)
    cases.append(("socp", d, K, dict(eps_abs=1e-8, eps_rel=1e-8, max_iters=100000)))
    d, K, aux = P.maxcut_sdp(seed=1, nodes=30, blocks=8)
    cases.append(("sdp", d, K, dict(eps_abs=1e-8, eps_rel=1e-8, max_iters=100000)))
    ref_solve, ref_kind = reference_solver()
    single, refs = {}, {}
    for name, d, K, kw in cases:
        t = time.perf_counter()
        single[name] = scsb.SCS(d, K, verbose=False, **kw).solve()
        single[name]["wall"] = time.perf_counter() - t
        if rank == 0:
            g = ref_solve(d, K, kw)["info"]
            refs[name] = dict(status_val=g["status_val"], status=g["status"], pobj=g["pobj"], dobj=g["dobj"], iter=g["iter"])
    box = [refs]
    td.broadcast_object_list(box, src=0)
    refs = box[0]
    scsb.dist_init(rank, world)
    ok = True
    for name, d, K, kw in cases:
        t = time.perf_counter()
        s = scsb.SCS(d, K, verbose=False, **kw)
        r = s.solve()
        wall = time.perf_counter() - t
        st = s._solver.stats()
        a, b, g = single[name]["info"], r["info"], refs[name]
        rel = lambda u, v: abs(u - v) / max(1.0, abs(v))
        errs = dict(pobj_ref=rel(b["pobj"], g["pobj"]), dobj_ref=rel(b["dobj"], g["dobj"]),
                    x=float(np.max(np.abs(single[name]["x"] - r["x"])) / max(1.0, np.max(np.abs(r["x"])))),
                    y=float(np.max(np.abs(single[name]["y"] - r["y"])) / max(1.0, np.max(np.abs(r["y"])))),
                    s=float(np.max(np.abs(single[name]["s"] - r["s"])) / max(1.0, np.max(np.abs(r["s"])))))
        good = (b["status_val"] == g["status_val"] == 1 and errs["pobj_ref"] < 1e-6 and errs["dobj_ref"] < 1e-6
                and errs["x"] < 1e-5 and errs["y"] < 1e-5 and errs["s"] < 1e-5)
        ok = ok and good
        if rank == 0:
            print(json.dumps(dict(case=name, ok=bool(good), world=world, ref=ref_kind, status=(b["status"], g["status"]),
                                  iters=dict(dist=b["iter"], single=a["iter"], ref=g["iter"]), errs=errs,
                                  wall_single=single[name]["wall"], wall_dist=wall, solve_ms=(a["solve_time"], b["solve_time"]),
                                  accepted_aa=b["accepted_accel_steps"], collectives=st["collectives"],
                                  collective_mb=st["collective_bytes"] / 1e6)), flush=True)
    # verbose + max_iters: rank 0 prints, every rank issues the same collectives
    d, K, aux = P.lasso(2000, 4000, 10, seed=3)
    r = scsb.SCS(d, K, verbose=True, max_iters=60, eps_abs=1e-14, eps_rel=1e-14).solve()
    ok = ok and r["info"]["iter"] == 60 and r["info"]["status_val"] == 2
    flag = torch.tensor([1 if ok else 0], device="cuda")
    td.all_reduce(flag, op=td.ReduceOp.MIN)
    scsb.dist_finalize()
    td.destroy_process_group()
    if int(flag.item()) != 1:
        sys.exit(1)
    if rank == 0:
        print("dist check ok")


if __name__ == "__main__":
    main()
