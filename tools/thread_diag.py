#!/usr/bin/env python
"""Diagnostic: is a solve reproducible (a) when workspaces are created / destroyed repeatedly in ONE thread,
(b) when two threads do that concurrently?  Prints the iteration counts and statuses of every solve."""
import json, os, sys, threading
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scs_python_b200 as scsb
from tests import problems

K = dict(z=5, l=40, q=[6, 5, 4], ep=4)
probs = [problems.gen_feasible(K, n=60, density=0.2, seed=s, with_P=bool(s % 2))[0] for s in (1, 2)]
kw = dict(verbose=False, eps_abs=1e-9, eps_rel=1e-9, max_iters=50000)


def pair(d):
    s = scsb.SCS(d, K, **kw)
    r = s.solve()
    s.update(b=np.asarray(d["b"]), c=np.asarray(d["c"]))
    r2 = s.solve(warm_start=False)
    return [(r["info"]["status_val"], r["info"]["iter"], float(r["x"][0])), (r2["info"]["status_val"], r2["info"]["iter"], float(r2["x"][0]))]


print("single thread, repeated:")
for t in range(2):
    print(t, [pair(probs[t]) for _ in range(4)], flush=True)
res = [[], []]


def work(t):
    for _ in range(6):
        res[t].append(pair(probs[t]))


for mode in ("graph", "nograph"):
    if mode == "nograph":
        os.environ["SCS_B200_NO_GRAPH"] = "1"
    res[0].clear(); res[1].clear()
    th = [threading.Thread(target=work, args=(t,)) for t in range(2)]
    [t.start() for t in th]; [t.join() for t in th]
    print("two threads,", mode)
    for t in range(2):
        print(t, res[t], flush=True)
