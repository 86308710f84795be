"""Seeded synthetic problem builders for the BASELINE.json configurations.

Pure numpy/scipy input construction -- nothing here runs on the solve path.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def lasso(n0, m0, nnz_per_col, seed, lam_frac=0.1):
    """Sparse LASSO in the reference documentation's formulation
    (S/docs/src/examples/python/lasso.py:23-45):

        min 0.5 ||y||^2 + lam 1't   s.t.  y = Ad x - b0,  -t <= x <= t
        variables (x in R^n0, y in R^m0, t in R^n0)
        P = blkdiag(0, I_m0, 0),  A = [[Ad, -I, 0], [I, 0, -I], [-I, 0, -I]],  cone z=m0, l=2 n0

    so the SCS sizes are n = 2 n0 + m0, m = m0 + 2 n0, nnz(A) = nnz(Ad) + m0 + 4 n0.
    BASELINE.json config 2 ("n=1M, m=2M, nnz(A)=100M") is read as the DATA matrix
    Ad in R^{2M x 1M} with 100 non-zeros per column.

    Ad is built directly in CSC form: every column gets `nnz_per_col` distinct rows by
    stratified sampling (row = i*w + floor(u*w), w = m0 // k), values N(0,1) -- no sort, no COO.
    Returns (data, cone, aux) with data = dict(A=csc, P=csc, b, c).
    """
    rng = np.random.RandomState(seed)
    k = int(nnz_per_col)
    # --- Ad in CSC
    w = m0 // k  # integer stratum width -> rows strictly increasing, hence distinct
    assert w >= 1, "need m0 >= nnz_per_col"
    u = rng.random_sample((n0, k))
    rows = (np.arange(k, dtype=np.int64)[None, :] * w + np.floor(u * w).astype(np.int64)).astype(np.int32)
    del u
    vals = rng.standard_normal((n0, k))
    # --- data vectors (need Ad as an operator)
    Ad = sp.csc_matrix((vals.ravel(), rows.ravel(), np.arange(0, n0 * k + 1, k, dtype=np.int64)), shape=(m0, n0))
    x_true = np.where(rng.random_sample(n0) < 0.01, rng.standard_normal(n0), 0.0)
    b0 = Ad @ x_true + 0.1 * rng.standard_normal(m0)
    lam = lam_frac * float(np.max(np.abs(Ad.T @ b0)))
    del Ad
    # --- full A in CSC, column blocks [x | y | t]
    j = np.arange(n0, dtype=np.int32)
    xi = np.empty((n0, k + 2), dtype=np.int32)
    xi[:, :k] = rows
    xi[:, k] = m0 + j
    xi[:, k + 1] = m0 + n0 + j
    xv = np.empty((n0, k + 2))
    xv[:, :k] = vals
    xv[:, k] = 1.0
    xv[:, k + 1] = -1.0
    del rows, vals
    yi = np.arange(m0, dtype=np.int32)
    yv = -np.ones(m0)
    ti = np.empty((n0, 2), dtype=np.int32)
    ti[:, 0] = m0 + j
    ti[:, 1] = m0 + n0 + j
    tv = -np.ones((n0, 2))
    indices = np.concatenate([xi.ravel(), yi, ti.ravel()])
    dat = np.concatenate([xv.ravel(), yv, tv.ravel()])
    del xi, xv
    nx, ny = n0 * (k + 2), m0
    indptr = np.concatenate([np.arange(0, nx, k + 2, dtype=np.int64),
                             nx + np.arange(0, ny, 1, dtype=np.int64),
                             nx + ny + np.arange(0, 2 * n0 + 1, 2, dtype=np.int64)]).astype(np.int32)
    n, m = 2 * n0 + m0, m0 + 2 * n0
    A = sp.csc_matrix((dat, indices, indptr), shape=(m, n))
    A.has_sorted_indices = True
    # --- P = blkdiag(0, I, 0) (upper triangular by construction)
    pp = np.concatenate([np.zeros(n0, dtype=np.int32), np.arange(0, m0 + 1, dtype=np.int32),
                         np.full(n0, m0, dtype=np.int32)])
    P = sp.csc_matrix((np.ones(m0), np.arange(n0, n0 + m0, dtype=np.int32), pp), shape=(n, n))
    P.has_sorted_indices = True
    b = np.concatenate([b0, np.zeros(2 * n0)])
    c = np.concatenate([np.zeros(n0 + m0), lam * np.ones(n0)])
    return dict(A=A, P=P, b=b, c=c), dict(z=m0, l=2 * n0), dict(lam=lam, n0=n0, m0=m0, nnz_Ad=n0 * k)
