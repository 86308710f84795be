#!/usr/bin/env python
"""SpMV roofline sweep on the LASSO workload (dev tool, GPU box only).

    python tools/spmv_sweep.py --scale 0.25 --reps 20

Prints achieved algorithmic GB/s of the two CG SpMV kernels (z = R_y^-1 A p and
Gp = A'z + P p + R_x p) through scs_b200_bench_spmv (CUDA events on the workspace stream).
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import scs_python_b200 as scsb  # noqa: E402
from scs_python_b200 import problems  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.25)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    n0 = int(1_000_000 * a.scale)
    t = time.time()
    data, cone, _ = problems.lasso(n0, 2 * n0, 100, 0)
    t_gen = time.time() - t
    t = time.time()
    s = scsb.SCS(data, cone, verbose=False, max_iters=10)
    t_init = time.time() - t
    out = dict(scale=a.scale, nnz_A=int(data["A"].nnz), gen_s=t_gen, init_s=t_init)
    for which, name in ((0, "A"), (1, "G")):
        ms, nbytes = s._solver.bench_spmv(which, a.reps)
        out[name] = dict(ms=ms, bytes=nbytes, gbs=nbytes / (ms * 1e-3) / 1e9)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
