#!/usr/bin/env python
"""Per-CTA times of one launch of the tiled SpMV kernels on the bench workload (dev tool)."""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scs_python_b200 as scsb
from scs_python_b200 import problems as P

ap = argparse.ArgumentParser(); ap.add_argument("--scale", type=float, default=1.0); ap.add_argument("--reps", type=int, default=10); a = ap.parse_args()
data, cone, _ = P.lasso(int(1_000_000 * a.scale), int(2_000_000 * a.scale), 100, seed=0)
s = scsb.SCS(data, cone, verbose=False, eps_infeas=1e-12)
for which in (0, 1):
    ms, ab = s._solver.bench_spmv(which, a.reps)
    pr = s._solver.tiled_profile(which)
    print("op %d: %.3f ms/launch, %.0f GB/s; CTAs %d" % (which, ms, ab / ms / 1e6, len(pr)))
    if len(pr):
        d, st, cost, items = pr[:, 0], pr[:, 1], pr[:, 2], pr[:, 3]
        print("  duration us: min %.0f mean %.0f max %.0f | stream us: min %.0f mean %.0f max %.0f | cost max/mean %.3f"
              % (d.min(), d.mean(), d.max(), st.min(), st.mean(), st.max(), cost.max() / cost.mean()))
        order = np.argsort(-d)
        for b in list(order[:6]) + list(order[-4:]):
            print("   cta %3d: %.0f us (stream %.0f) cost %.3g items %d  us/cost %.3g" % (b, d[b], st[b], cost[b], items[b], d[b] / cost[b]))
        print("  corr(duration, cost) = %.3f" % np.corrcoef(d, cost)[0, 1])
