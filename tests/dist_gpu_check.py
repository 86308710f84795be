#!/usr/bin/env python
"""Row-partitioned solve vs. single-GPU solve of the same problem (run under torchrun, >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29655 tests/dist_gpu_check.py

Every rank first solves each problem alone on its GPU (Anderson acceleration off, as in the
partitioned mode), then the ranks solve it together.  Same status; objectives, x, y, s agree to
the order of the different summation order of the all-reduce (the iterates are not bit-equal,
so iteration counts may differ by one convergence check at loose tolerance).
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


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    td.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    assert B.lib.scs_b200_set_device(local) == 0
    cases = []
    d, K, aux = P.random_cone_qp(seed=7, n=400, l=600, nq=60, q=6, ep=40, density=0.03)
    cases.append(("cone_qp", d, K, dict(eps_abs=1e-9, eps_rel=1e-9, max_iters=50000)))
    d, K, aux = P.lasso(20000, 40000, 50, seed=2)
    cases.append(("lasso", d, K, dict(eps_abs=1e-6, eps_rel=1e-6, eps_infeas=1e-12, max_iters=20000)))
    d, K, aux = P.socp_portfolio(seed=1, n=3000, ncones=600)
    cases.append(("socp", d, K, dict(eps_abs=1e-7, eps_rel=1e-7, max_iters=50000)))
    d, K, aux = P.maxcut_sdp(seed=1, nodes=30, blocks=8)
    cases.append(("sdp", d, K, dict(eps_abs=1e-7, eps_rel=1e-7, max_iters=50000)))
    single = {}
    for name, d, K, kw in cases:
        t = time.perf_counter()
        single[name] = scsb.SCS(d, K, verbose=False, acceleration_lookback=0, **kw).solve()
        single[name]["wall"] = time.perf_counter() - t
    scsb.dist_init(rank, world)
    ok = True
    for name, d, K, kw in cases:
        t = time.perf_counter()
        s = scsb.SCS(d, K, verbose=False, **kw)
        r = s.solve()
        wall = time.perf_counter() - t
        st = s._solver.stats()
        a, b = single[name]["info"], r["info"]
        rel = lambda u, v: abs(u - v) / max(1.0, abs(u))
        errs = dict(pobj=rel(a["pobj"], b["pobj"]), dobj=rel(a["dobj"], b["dobj"]),
                    x=float(np.max(np.abs(single[name]["x"] - r["x"])) / max(1.0, np.max(np.abs(r["x"])))),
                    y=float(np.max(np.abs(single[name]["y"] - r["y"])) / max(1.0, np.max(np.abs(r["y"])))),
                    s=float(np.max(np.abs(single[name]["s"] - r["s"])) / max(1.0, np.max(np.abs(r["s"])))))
        good = (a["status_val"] == b["status_val"] == 1 and errs["pobj"] < 1e-6 and errs["dobj"] < 1e-6
                and errs["x"] < 1e-4 and errs["y"] < 1e-4 and errs["s"] < 1e-4)
        ok = ok and good
        if rank == 0:
            print(json.dumps(dict(case=name, ok=bool(good), world=world, status=(a["status"], b["status"]),
                                  iters=(a["iter"], b["iter"]), errs=errs, wall_single=single[name]["wall"], wall_dist=wall,
                                  solve_ms=(a["solve_time"], b["solve_time"]),
                                  collectives=st["collectives"], collective_mb=st["collective_bytes"] / 1e6)), flush=True)
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
