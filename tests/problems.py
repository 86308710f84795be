"""Seeded synthetic problem generators shared by the tests and bench.py (own code).

gen_feasible follows the construction used by the reference's test tool
(test/gen_random_cone_prob.py:9-24: draw z, split it with the dual-cone projection into
complementary (s, y), draw A and x, set b = A x + s, c = -A' y) so that the optimum is
known; the cone projection used for the split is the oracle's (test infrastructure).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def cone_len(K):
    return (K.get("z", 0) + K.get("l", 0) + (len(K.get("bu", [])) + 1 if len(K.get("bu", [])) else 0)
            + sum(K.get("q", [])) + sum(s * (s + 1) // 2 for s in K.get("s", []))
            + sum(c * c for c in K.get("cs", [])) + 3 * (K.get("ep", 0) + K.get("ed", 0) + len(K.get("p", []))))


def gen_feasible(K, n, density, seed, with_P=False):
    from oracle import scs_oracle as O
    rng = np.random.RandomState(seed)
    m = cone_len(K)
    z = rng.randn(m)
    cw = O.ConeWork(K, m)
    y = z.copy()
    O.proj_dual_cone(y, cw, None, None)   # y = Pi_{K*}(z)
    s = y - z                              # s in K, s'y = 0
    A = sp.random(m, n, density=density, format="csc", random_state=rng, data_rvs=rng.randn)
    x = rng.randn(n)
    c = -A.T @ y
    b = A @ x + s
    data = dict(A=A.tocsc(), b=b, c=c)
    p_star = float(c @ x)
    if with_P:
        data["P"] = sp.eye(n, format="csc") * 0.1
        p_star = None
    return data, p_star


def gen_lasso(n0, m0, nnz_per_col, seed):
    from scs_python_b200 import problems as P
    data, cone, _ = P.lasso(n0, m0, nnz_per_col, seed)
    return data, cone
