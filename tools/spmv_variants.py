#!/usr/bin/env python
"""Tiled SpMV engine variants on the bench workload, one process, one data set:
entry dealing (SCS_B200_TILED_DEAL = greedy / perm / r1), side-stream short-row pass (SCS_B200_TILED_SIDE), groups
in flight per warp (SCS_B200_TILED_U).  Prints ms per product (CUDA events, back-to-back launches) and the algorithmic
TB/s of both CG products for every combination.

    python tools/spmv_variants.py [--scale 1.0] [--reps 30]
"""
import argparse
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import scs_python_b200 as scsb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--combos", default="")
    a = ap.parse_args()
    data, cone = bench.workload(a.scale, seed=0)
    combos = list(itertools.product(("greedy", "perm"), ("0",), ("8",)))
    if a.combos:
        combos = [tuple(c.split(",")) for c in a.combos.split(";")]
    ref_obj = None
    for combo in combos:
        perm, side, u = combo[:3]
        eb = combo[3] if len(combo) > 3 else "-"   # epilogue pass: - = default, 0 / 1 = chunked kernel (batched / not), 4 / 8 = row-parallel
        os.environ.update(SCS_B200_TILED_DEAL=perm, SCS_B200_TILED_SIDE=side, SCS_B200_TILED_U=u, SCS_B200_TILED_EB=eb)
        if eb in ("", "-"):   # the library's own per-operator default
            os.environ.pop("SCS_B200_TILED_EB")
        s = scsb.SCS(data, cone, verbose=False, max_iters=50, eps_abs=0.0, eps_rel=0.0, eps_infeas=0.0)
        inner = s._solver
        a_ms, a_b = inner.bench_spmv(0, a.reps)
        g_ms, g_b = inner.bench_spmv(1, a.reps)
        r = s.solve(warm_start=False)
        obj = r["info"]["pobj"]
        if ref_obj is None:
            ref_obj = obj
        print(json.dumps(dict(lib=os.path.basename(os.environ.get('SCS_B200_LIBPATH', 'libscsb200.so')), deal=perm, side=side, U=u, EB=eb, a_ms=a_ms, g_ms=g_ms, a_tbs=a_b / a_ms / 1e9, g_tbs=g_b / g_ms / 1e9,
                              g_frac_of_6553=g_b / g_ms / 1e9 / 6.5536, pobj_50its=obj,
                              pobj_rel_diff=abs(obj - ref_obj) / max(1.0, abs(ref_obj)))), flush=True)
        inner.finish()
        del s


if __name__ == "__main__":
    main()
