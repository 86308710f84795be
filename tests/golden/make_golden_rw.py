#!/usr/bin/env python
"""Golden fixtures for the data-file format and the CSV trace (S/src/rw.c), produced by the compiled
reference in this container (oracle/_ref, `make -C oracle ref`):

  rw_ref_mixed.bin   SCS(write_data) of a seeded mixed-cone QP (box + SOC + PSD + exp + power cones, P)
                     through the reference python package (`write_data_filename=`), 32-bit-int build
  rw_ref_trace.csv   SCS(log_data_to_csv) trace of the reference solving a seeded cone QP
                     (`log_csv_filename=`, CPU_INDIRECT backend -- the algorithm this backend replaces --, eps 1e-6): header, iterations 0..40, the
                     last iteration and the final row

    python tests/golden/make_golden_rw.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))


def rw_problem():
    """the problem behind rw_ref_mixed.bin (tests rebuild it from the same seed)"""
    from tests import problems
    K = dict(z=3, l=7, bu=[1.0, 2.5, 0.5], bl=[-1.0, -0.5, -3.0], q=[4, 1, 6], s=[3, 1, 4], ep=2, ed=1, p=[0.3, -0.6])
    data, _ = problems.gen_feasible(K, n=17, density=0.4, seed=21, with_P=True)
    stg = dict(eps_abs=1e-7, eps_rel=2e-7, eps_infeas=3e-8, max_iters=1234, scale=0.3, rho_x=2e-6, alpha=1.4,
               normalize=True, adaptive_scale=False, acceleration_lookback=7, acceleration_interval=3,
               acceleration_type_1=True, acceleration_regularization=1e-9, acceleration_relaxation=0.95)
    return data, K, stg


def trace_problem():
    from tests import problems
    # m = 139 rows for n = 30 columns: a well-conditioned reduced system, so that the first iterations do
    # not depend on where a loosely converged CG happens to stop (see tests/test_gpu_rw.py)
    K = dict(z=10, l=60, q=[10, 20, 15], ep=5, ed=3)
    data, _ = problems.gen_feasible(K, n=30, density=0.5, seed=5, with_P=True)
    return data, K, dict(eps_abs=1e-6, eps_rel=1e-6)


def main():
    import scs  # the reference package
    data, K, stg = rw_problem()
    out = os.path.join(HERE, "rw_ref_mixed.bin")
    scs.SCS(data, K, verbose=False, write_data_filename=out, **stg)
    print("wrote", out, os.path.getsize(out), "bytes; reference sizeof(int) =", scs.__sizeof_int__ if hasattr(scs, "__sizeof_int__") else "?")
    data, K, stg = trace_problem()
    out = os.path.join(HERE, "rw_ref_trace.csv")
    sol = scs.SCS(data, K, verbose=False, log_csv_filename=out, linear_solver=scs.LinearSolver.CPU_INDIRECT, **stg).solve()
    lines = open(out).read().splitlines()
    keep = lines[:42] + lines[-2:]   # header, iterations 0..40, the last iteration and the final row
    open(out, "w").write("\n".join(keep) + "\n")
    print("wrote", out, os.path.getsize(out), "bytes (%d of %d rows kept);" % (len(keep) - 1, len(lines) - 1),
          sol["info"]["status"], sol["info"]["iter"], "iterations")


if __name__ == "__main__":
    main()
