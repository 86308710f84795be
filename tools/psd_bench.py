#!/usr/bin/env python
"""PSD projection timing (dev tool, GPU box): s=[200]*64 (BASELINE Cfg-4 cone set) and a few others.

    python tools/psd_bench.py
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scs_python_b200 import _scs_b200 as B  # noqa: E402


def run(K, step, reps=10, warmup=2, seed=0):
    k, keep = B.make_cone(K)
    m = int(sum(s * (s + 1) // 2 for s in K.get("s", [])) + sum(c * c for c in K.get("cs", [])))
    rng = np.random.RandomState(seed)
    x0, x1 = rng.randn(m), rng.randn(m)
    w = B.lib.scs_b200_init_cone(C.byref(k), m)
    assert w
    sw = C.c_double(0.0)
    ms = B.lib.scs_b200_bench_proj_cone(w, B._dptr(x0), B._dptr(x1), step, reps, warmup, C.byref(sw))
    B.lib.scs_b200_finish_cone(w)
    return dict(cone={a: (str(b[:2]) + "x%d" % len(b) if isinstance(b, list) else b) for a, b in K.items()},
                step=step, ms_per_projection=ms, mean_sweeps=sw.value)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        print(json.dumps(run(dict(s=[200] * 64), float(sys.argv[2]) if len(sys.argv) > 2 else 1e-3, reps=3, warmup=1)))
        sys.exit(0)
    for K in (dict(s=[200] * 64), dict(s=[50] * 64), dict(s=[20] * 1000), dict(cs=[60] * 16)):
        for step in (0.0, 1e-6, 1e-3, 1e-1):
            print(json.dumps(run(K, step)), flush=True)
    # cpu reference point: numpy eigh of one 200x200
    a = np.random.RandomState(1).randn(200, 200); a = a + a.T
    t = time.perf_counter()
    for _ in range(20):
        np.linalg.eigh(a)
    print(json.dumps(dict(numpy_eigh_200_ms=(time.perf_counter() - t) / 20 * 1e3)))
