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


def gen_lasso(n0, m0, nnz_per_row, seed):
    """Sparse LASSO in the form of the reference documentation example
    (S/docs/src/examples/python/lasso.py:23-45): variables (x, y, t),
    min 0.5 y'y + lam 1't  s.t. y = Ad x - b0, -t <= x <= t.
    SCS sizes: n = 2 n0 + m0, m = m0 + 2 n0, cone z = m0, l = 2 n0."""
    rng = np.random.RandomState(seed)
    nnz = int(m0 * nnz_per_row)
    rows = rng.randint(0, m0, size=nnz)
    cols = rng.randint(0, n0, size=nnz)
    vals = rng.randn(nnz)
    Ad = sp.coo_matrix((vals, (rows, cols)), shape=(m0, n0)).tocsc()
    x_true = np.where(rng.rand(n0) < 0.01, rng.randn(n0), 0.0)
    b0 = Ad @ x_true + 0.1 * rng.randn(m0)
    lam = 0.1 * np.max(np.abs(Ad.T @ b0))
    In = sp.eye(n0, format="csc")
    Im = sp.eye(m0, format="csc")
    A = sp.bmat([[Ad, -Im, None], [In, None, -In], [-In, None, -In]], format="csc")
    # sp.bmat with None blocks needs explicit zero blocks for the t / y columns:
    A = sp.bmat([[Ad, -Im, sp.csc_matrix((m0, n0))],
                 [In, sp.csc_matrix((n0, m0)), -In],
                 [-In, sp.csc_matrix((n0, m0)), -In]], format="csc")
    P = sp.block_diag([sp.csc_matrix((n0, n0)), Im, sp.csc_matrix((n0, n0))], format="csc")
    b = np.concatenate([b0, np.zeros(2 * n0)])
    c = np.concatenate([np.zeros(n0 + m0), lam * np.ones(n0)])
    A.sort_indices()
    return dict(A=A, P=P, b=b, c=c), dict(z=m0, l=2 * n0)
