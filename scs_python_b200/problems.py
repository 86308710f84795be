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


# ---------------------------------------------------------------------------------------
# The other BASELINE.json configurations.  All are built around a known
# complementary pair (s in K, y in K*, s'y = 0) and a primal point x, with b = A x + s and
# c = -A'y - P x, so every instance is feasible and bounded by construction (the recipe of
# test/gen_random_cone_prob.py:9-24, without needing a cone projection to build it).
# ---------------------------------------------------------------------------------------
def _complementary_lin(rng, k):
    pick = rng.random_sample(k) < 0.5
    s = np.where(pick, rng.random_sample(k) + 0.1, 0.0)
    y = np.where(pick, 0.0, rng.random_sample(k) + 0.1)
    return s, y


def _complementary_soc(rng, q):
    """(s, y) on the boundary of the second-order cone: s = (t, u), y = a (t, -u), ||u|| = t."""
    if q == 1:
        return _complementary_lin(rng, 1)
    u = rng.standard_normal(q - 1)
    t = float(np.linalg.norm(u))
    a = rng.random_sample() + 0.1
    return np.concatenate([[t], u]), a * np.concatenate([[t], -u])


def _complementary_exp(rng):
    """Exponential cone triple: either s strictly inside K_exp with y = 0, or s = 0 with y strictly
    inside K_exp^* = {(u,v,w): u < 0, -u exp(v/u) <= e w}."""
    if rng.random_sample() < 0.5:
        yy = rng.random_sample() + 0.5
        xx = rng.standard_normal()
        zz = yy * np.exp(xx / yy) + rng.random_sample() + 0.1
        return np.array([xx, yy, zz]), np.zeros(3)
    u = -(rng.random_sample() + 0.5)
    v = rng.standard_normal()
    w = -u * np.exp(v / u) / np.e + rng.random_sample() + 0.1
    return np.zeros(3), np.array([u, v, w])


def _finish(A, P, x, s, y):
    b = A @ x + s
    c = -(A.T @ y)
    if P is not None:
        c = c - P @ x
    data = dict(A=sp.csc_matrix(A), b=b, c=c)
    if P is not None:
        data["P"] = sp.csc_matrix(sp.triu(P))
    p_star = float(c @ x + (0.5 * x @ (P @ x) if P is not None else 0.0))
    return data, p_star


def random_cone_qp(seed=1234, n=2000, l=3000, nq=240, q=10, ep=200, density=0.005, with_P=True):
    """Cfg-1: random cone program (l + q + ep cones), n = 2000, m = 6000, nnz(A) = 60 000, and the QP
    variant P = 0.1 I (pattern of test/test_mix_sd_csd_cone.py:6-12)."""
    rng = np.random.RandomState(seed)
    parts = [_complementary_lin(rng, l)] + [_complementary_soc(rng, q) for _ in range(nq)] + \
            [_complementary_exp(rng) for _ in range(ep)]
    s = np.concatenate([p[0] for p in parts]); y = np.concatenate([p[1] for p in parts])
    m = s.size
    A = sp.random(m, n, density=density, format="csc", random_state=rng, data_rvs=rng.standard_normal)
    x = rng.standard_normal(n)
    P = (0.1 * sp.eye(n, format="csc")) if with_P else None
    data, p_star = _finish(A, P, x, s, y)
    return data, dict(l=l, q=[q] * nq, ep=ep), dict(p_star=p_star)


def socp_portfolio(seed=0, n=50_000, ncones=10_000, nnz_per_row=10):
    """Cfg-3: SOCP of portfolio shape: budget row (zero cone), long-only rows (-I, nonneg cone) and
    `ncones` second-order cones of sizes 3..33 whose rows are sparse factor exposures; P = diag
    (idiosyncratic risk).  n = 50 000, m = 50 001 + sum(q) ~ 230 k, nnz(A) ~ 2 M."""
    rng = np.random.RandomState(seed)
    qs = rng.randint(3, 34, size=ncones).tolist()
    mq = int(np.sum(qs))
    budget = sp.csr_matrix(np.ones((1, n)))
    cols = rng.randint(0, n, size=(mq, nnz_per_row))
    rows = np.repeat(np.arange(mq), nnz_per_row)
    F = sp.csr_matrix((rng.standard_normal(mq * nnz_per_row) / np.sqrt(nnz_per_row), (rows, cols.ravel())), shape=(mq, n))
    A = sp.vstack([budget, -sp.eye(n, format="csr"), F], format="csc")
    A.sum_duplicates(); A.sort_indices()
    sz, yz = np.zeros(1), rng.standard_normal(1)
    sl, yl = _complementary_lin(rng, n)
    parts = [_complementary_soc(rng, q) for q in qs]
    s = np.concatenate([sz, sl] + [p[0] for p in parts]); y = np.concatenate([yz, yl] + [p[1] for p in parts])
    x = np.abs(rng.standard_normal(n)) / n
    P = sp.diags(0.01 + 0.1 * rng.random_sample(n), format="csc")
    data, p_star = _finish(A, P, x, s, y)
    return data, dict(z=1, l=n, q=qs), dict(p_star=p_star)


def maxcut_sdp(seed=0, nodes=200, blocks=64, p_edge=0.1):
    """Cfg-4: `blocks` independent MaxCut SDP relaxations (Erdos-Renyi graphs) stacked block-diagonally,
    dual form  min 1'y  s.t.  Diag(y) - L/4 >= 0 (PSD)  per block.  SCS form: A y + s = b with
    s = svec(Diag(y) - L/4): A has one -1 per variable (at its diagonal position), b = -svec(L/4);
    off-diagonal entries of svec carry the sqrt(2) scaling (S/docs/src/api/cones.rst)."""
    rng = np.random.RandomState(seed)
    k = nodes
    tri = k * (k + 1) // 2
    # position of diagonal entry (j, j) in the column-major lower-triangular packing
    diag_pos = np.array([j * k - (j - 1) * j // 2 for j in range(k)], dtype=np.int64)
    rows, cols, bvec = [], [], []
    for blk in range(blocks):
        W = np.triu((rng.random_sample((k, k)) < p_edge).astype(float), 1)
        W = W + W.T
        L = np.diag(W.sum(1)) - W
        M = -L / 4.0
        il, jl = np.tril_indices(k)           # row-major order of the lower triangle ...
        order = np.lexsort((il, jl))          # ... re-sorted to column-major packing
        il, jl = il[order], jl[order]
        vals = M[il, jl] * np.where(il == jl, 1.0, np.sqrt(2.0))
        bvec.append(vals)
        rows.append(blk * tri + diag_pos)
        cols.append(blk * k + np.arange(k))
    n, m = blocks * k, blocks * tri
    A = sp.csc_matrix((-np.ones(n), (np.concatenate(rows), np.concatenate(cols))), shape=(m, n))
    data = dict(A=A, b=np.concatenate(bvec), c=np.ones(n))
    return data, dict(s=[k] * blocks), dict()


def mpc_qp(seed=0, nx=12, nu=6, T=6):
    """Cfg-5: one small MPC QP in the shape of S/docs/src/examples/python/mpc.py:12-65 with the bounds
    written as nonneg-cone rows: variables [x_0..x_T, u_0..u_{T-1}] (n = (T+1) nx + T nu = 120),
    rows: dynamics (zero cone, (T+1) nx = 84), two-sided variable bounds (2 n = 240) and one-sided
    input-rate limits u_t - u_{t-1} <= du_max, u_{-1} = 0 (T nu = 36)  =>  m = 360."""
    rng = np.random.RandomState(seed)
    Ad = 0.95 * np.eye(nx) + 0.1 * rng.standard_normal((nx, nx))
    Bd = rng.standard_normal((nx, nu))
    x0 = 10.0 * rng.standard_normal(nx)
    n = (T + 1) * nx + T * nu
    Q, R = sp.eye(nx), 0.1 * sp.eye(nu)
    P = sp.block_diag([sp.kron(sp.eye(T + 1), Q), sp.kron(sp.eye(T), R)], format="csc")
    Ax = sp.kron(sp.eye(T + 1), -sp.eye(nx)) + sp.kron(sp.eye(T + 1, k=-1), sp.csc_matrix(Ad))
    Bu = sp.kron(sp.vstack([sp.csc_matrix((1, T)), sp.eye(T)]), sp.csc_matrix(Bd))
    Aeq = sp.hstack([Ax, Bu])
    beq = np.zeros((T + 1) * nx); beq[:nx] = -x0
    xmax, umax, dumax = 100.0, 2.0, 1.0
    ub = np.concatenate([np.full((T + 1) * nx, xmax), np.full(T * nu, umax)])
    In = sp.eye(n, format="csc")
    Du = sp.hstack([sp.csc_matrix((T * nu, (T + 1) * nx)), sp.kron(sp.eye(T) - sp.eye(T, k=-1), sp.eye(nu))])
    A = sp.vstack([Aeq, In, -In, Du], format="csc")
    A.sort_indices()
    b = np.concatenate([beq, ub, ub, np.full(T * nu, dumax)])
    c = np.zeros(n)
    return dict(A=A, P=sp.csc_matrix(sp.triu(P)), b=b, c=c), dict(z=(T + 1) * nx, l=A.shape[0] - (T + 1) * nx), dict()


def mpc_qp_box(seed=0, nx=12, nu=6, T=6):
    """Cfg-5 in the reference example's own form (S/docs/src/examples/python/mpc.py:12-65): the same dynamics,
    weights and bounds as mpc_qp, with the two-sided variable bounds carried by ONE box cone
    {(t, s) : t bl <= s <= t bu} whose t row is pinned to 1 by b (mpc.py:49-60)  =>  rows: dynamics (zero
    cone, 84), box cone (1 + n = 121), input-rate limits (nonneg cone, 36), in the reference's cone order
    z, l, box (S/include/scs.h ScsCone)  =>  m = 241."""
    rng = np.random.RandomState(seed)
    Ad = 0.95 * np.eye(nx) + 0.1 * rng.standard_normal((nx, nx))
    Bd = rng.standard_normal((nx, nu))
    x0 = 10.0 * rng.standard_normal(nx)
    n = (T + 1) * nx + T * nu
    Q, R = sp.eye(nx), 0.1 * sp.eye(nu)
    P = sp.block_diag([sp.kron(sp.eye(T + 1), Q), sp.kron(sp.eye(T), R)], format="csc")
    Ax = sp.kron(sp.eye(T + 1), -sp.eye(nx)) + sp.kron(sp.eye(T + 1, k=-1), sp.csc_matrix(Ad))
    Bu = sp.kron(sp.vstack([sp.csc_matrix((1, T)), sp.eye(T)]), sp.csc_matrix(Bd))
    Aeq = sp.hstack([Ax, Bu])
    beq = np.zeros((T + 1) * nx); beq[:nx] = -x0
    xmax, umax, dumax = 100.0, 2.0, 1.0
    ub = np.concatenate([np.full((T + 1) * nx, xmax), np.full(T * nu, umax)])
    Du = sp.hstack([sp.csc_matrix((T * nu, (T + 1) * nx)), sp.kron(sp.eye(T) - sp.eye(T, k=-1), sp.eye(nu))])
    A = sp.vstack([Aeq, Du, sp.csc_matrix((1, n)), -sp.eye(n, format="csc")], format="csc")
    A.sort_indices()
    b = np.concatenate([beq, np.full(T * nu, dumax), [1.0], np.zeros(n)])
    c = np.zeros(n)
    return dict(A=A, P=sp.csc_matrix(sp.triu(P)), b=b, c=c), dict(z=(T + 1) * nx, l=T * nu, bu=ub.copy(), bl=-ub), dict()
