"""GPU tier (-m gpu): the CUDA path, called through the C ABI of libscsb200.so, against
  * the committed golden fixtures (reference KATs + outputs of the compiled reference),
  * the numpy oracle on the same seeded inputs,
  * the compiled reference itself when oracle/_ref travelled to this box, and
  * size-independent properties (residual criteria recomputed on the host, cone membership,
    adjointness / linearity of the SpMV pair) at sizes the oracle cannot reach.
Tolerances: solver status identical; objectives within 1e-6 relative at eps 1e-9 (north_star);
kernel-level outputs within 1e-9 relative (FP64, different summation order).
"""
import ctypes as C
import math
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import scs_oracle as O
from tests import helpers, problems

pytestmark = pytest.mark.gpu

KAT = helpers.golden("kat.json")
REF = helpers.golden("ref_runs.json")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def B(gpu):
    from scs_python_b200 import _scs_b200 as mod
    return mod


@pytest.fixture(scope="module")
def scsb(gpu):
    import scs_python_b200
    return scs_python_b200


def _mat(B, M):
    M = sp.csc_matrix(M); M.sort_indices()
    x = np.ascontiguousarray(M.data, dtype=np.float64); i = M.indices.astype(np.int32); p = M.indptr.astype(np.int32)
    return B.make_matrix(x, i, p, M.shape[0], M.shape[1]), (x, i, p)


# ------------------------------------------------------------------------------- SpMV ------
@pytest.mark.parametrize("shape", [(1, 1, 1.0), (50, 30, 0.3), (3000, 2000, 0.01), (20000, 100, 0.9), (5, 40000, 0.5),
                                   (300, 300, 0.0)])
def test_accum_by_a_and_atrans(B, shape):
    m, n, dens = shape
    rng = np.random.RandomState(m + n)
    A = sp.random(m, n, density=dens, format="csc", random_state=rng, data_rvs=rng.randn)
    M, keep = _mat(B, A)
    x, y0 = rng.randn(n), rng.randn(m)
    y = y0.copy()
    assert B.lib.scs_b200_accum_by_a(C.byref(M), B._dptr(x), B._dptr(y)) == 0   # scs_matrix.c:162-177
    ref = y0 + A @ x
    assert np.max(np.abs(y - ref)) <= 1e-12 * max(1.0, np.max(np.abs(ref)))
    xt, z0 = rng.randn(m), rng.randn(n)
    z = z0.copy()
    assert B.lib.scs_b200_accum_by_atrans(C.byref(M), B._dptr(xt), B._dptr(z)) == 0  # scs_matrix.c:135-160
    ref = z0 + A.T @ xt
    assert np.max(np.abs(z - ref)) <= 1e-12 * max(1.0, np.max(np.abs(ref)))


def test_accum_by_p_symmetric_upper(B):
    rng = np.random.RandomState(5)
    n = 700
    Pf = sp.random(n, n, density=0.01, format="csc", random_state=rng, data_rvs=rng.randn)
    Pf = Pf + Pf.T + sp.eye(n)
    M, keep = _mat(B, sp.triu(Pf, format="csc"))
    x, y0 = rng.randn(n), rng.randn(n)
    y = y0.copy()
    assert B.lib.scs_b200_accum_by_p(C.byref(M), B._dptr(x), B._dptr(y)) == 0  # scs_matrix.c:180-199
    assert np.max(np.abs(y - (y0 + Pf @ x))) <= 1e-12 * np.max(np.abs(y))


def test_spmv_adjointness_and_linearity_large(B):
    """size-independent properties at a size the oracle is too slow for: y'(A x) == x'(A' y)."""
    from scs_python_b200 import problems as P
    data, cone, _ = P.lasso(100_000, 200_000, 100, seed=9)   # nnz(A) = 10.6M
    A = data["A"]
    m, n = A.shape
    M, keep = _mat(B, A)
    rng = np.random.RandomState(1)
    x1, x2, yv = rng.randn(n), rng.randn(n), rng.randn(m)
    ax1, ax2, ax12, aty = np.zeros(m), np.zeros(m), np.zeros(m), np.zeros(n)
    assert B.lib.scs_b200_accum_by_a(C.byref(M), B._dptr(x1), B._dptr(ax1)) == 0
    assert B.lib.scs_b200_accum_by_a(C.byref(M), B._dptr(x2), B._dptr(ax2)) == 0
    x12 = x1 + 2.0 * x2
    assert B.lib.scs_b200_accum_by_a(C.byref(M), B._dptr(x12), B._dptr(ax12)) == 0
    assert B.lib.scs_b200_accum_by_atrans(C.byref(M), B._dptr(yv), B._dptr(aty)) == 0
    assert np.max(np.abs(ax12 - (ax1 + 2.0 * ax2))) <= 1e-11 * np.max(np.abs(ax12))
    lhs, rhs = float(yv @ ax1), float(x1 @ aty)
    assert abs(lhs - rhs) <= 1e-10 * max(abs(lhs), np.linalg.norm(yv) * np.linalg.norm(ax1))


# ---------------------------------------------------------------- linear-system plugin ABI --
@pytest.mark.parametrize("idx", range(len(REF["solve_lin_sys"]["cases"])))
def test_lin_sys_plugin_vs_reference_golden(B, idx):
    c = REF["solve_lin_sys"]["cases"][idx]
    m, n = c["m"], c["n"]
    MA, k1 = _mat(B, sp.csc_matrix((c["Ax"], c["Ai"], c["Ap"]), shape=(m, n)))
    MP, k2 = _mat(B, sp.csc_matrix((c["Px"], c["Pi"], c["Pp"]), shape=(n, n))) if "Px" in c else (None, None)
    diag_r = np.array(c["diag_r"], float)
    w = B.lib.scs_init_lin_sys_work(C.byref(MA), C.byref(MP) if MP is not None else None, B._dptr(diag_r))
    assert w
    try:
        b, s = np.array(c["b"], float), np.array(c["s"], float)
        assert B.lib.scs_solve_lin_sys(w, B._dptr(b), B._dptr(s), c["tol"]) == 0
        assert np.max(np.abs(b - np.array(c["out_warm"]))) <= 1e-8
        b = np.array(c["b"], float)
        assert B.lib.scs_solve_lin_sys(w, B._dptr(b), None, c["tol"]) == 0
        assert np.max(np.abs(b - np.array(c["out_cold"]))) <= 1e-8
        z = np.zeros(n + m)   # ||b||_inf <= 1e-12 short-circuit (private.c:288-291)
        assert B.lib.scs_solve_lin_sys(w, B._dptr(z), None, 1e-9) == 0 and np.all(z == 0.0)
    finally:
        B.lib.scs_free_lin_sys_work(w)


def test_lin_sys_update_diag_r_and_exactness(B):
    rng = np.random.RandomState(3)
    m, n = 900, 300
    A = sp.random(m, n, density=0.03, format="csc", random_state=rng, data_rvs=rng.randn)
    Q = sp.random(n, n, density=0.01, format="csc", random_state=rng, data_rvs=rng.randn)
    Pf = Q @ Q.T + 0.1 * sp.eye(n)
    MA, k1 = _mat(B, A)
    MP, k2 = _mat(B, sp.triu(Pf, format="csc"))
    for scale in (0.1, 7.0):
        diag_r = np.concatenate([np.full(n, 1e-6), np.full(m, 1.0 / scale), [10.0]])
        diag_r[n:n + 10] = 1.0 / (1000 * scale)
        if scale == 0.1:
            w = B.lib.scs_init_lin_sys_work(C.byref(MA), C.byref(MP), B._dptr(diag_r))
            assert w
        else:
            assert B.lib.scs_update_lin_sys_diag_r(w, B._dptr(diag_r)) == 0    # linsys.h:64
        Kkt = sp.bmat([[sp.diags(diag_r[:n]) + Pf, A.T], [A, -sp.diags(diag_r[n:n + m])]], format="csc")
        b = rng.randn(n + m)
        import scipy.sparse.linalg as sla
        exact = sla.spsolve(Kkt, b)
        got = b.copy()
        assert B.lib.scs_solve_lin_sys(w, B._dptr(got), B._dptr(rng.randn(n)), 1e-12) == 0
        assert np.max(np.abs(got - exact)) <= 1e-7 * max(1.0, np.max(np.abs(exact)))
    assert B.lib.scs_b200_lin_sys_cg_its(w) > 0
    B.lib.scs_free_lin_sys_work(w)


# ------------------------------------------------------------------------ cone projections --
def _proj(B, K, x, r_y, D=None):
    m = len(x)
    k, keep = B.make_cone(K)
    w = B.lib.scs_b200_init_cone(C.byref(k), m)
    assert w
    out = np.array(x, dtype=np.float64)
    r = None if r_y is None else np.ascontiguousarray(r_y, dtype=np.float64)
    rc = B.lib.scs_b200_proj_dual_cone(B._dptr(out), w, None if D is None else B._dptr(D), None if r is None else B._dptr(r))
    B.lib.scs_b200_finish_cone(w)
    assert rc == 0
    return out


@pytest.mark.parametrize("idx", range(len(REF["proj_dual_cone"]["cases"])))
def test_proj_dual_cone_vs_reference_golden(B, idx):
    c = REF["proj_dual_cone"]["cases"][idx]
    got = _proj(B, c["cone"], np.array(c["x"], float), np.array(c["r_y"], float))
    ref = np.array(c["out"], float)
    assert np.max(np.abs(got - ref)) <= 1e-9 * max(1.0, np.max(np.abs(ref)))


def test_exp_cone_known_answers(B):
    """test_exp_cone.h:52-77 through the Moreau identity Pi_K(v) = v + Pi_{K*}(-v)."""
    k = KAT["exp_cone"]
    for v0, vp, vd in zip(k["v0"], k["vp"], k["vd"]):
        v0 = np.array(v0, float)
        # cone {ed:1}: K = Kexp*, K* = Kexp  =>  proj_dual_cone(v0) = Pi_Kexp(v0)
        assert np.linalg.norm(_proj(B, dict(ed=1), v0, None) - np.array(vp)) <= k["tol"]
        # cone {ep:1}: K = Kexp, K* = Kexp*  =>  proj_dual_cone(v0) = Pi_Kexp*(v0)
        assert np.linalg.norm(_proj(B, dict(ep=1), v0, None) - np.array(vd)) <= k["tol"]


@pytest.mark.parametrize("K", [dict(q=[3000, 2, 1, 0, 17] + [5] * 400), dict(s=[50]), dict(s=[2, 9, 0, 1, 33, 64]),
                               dict(cs=[1, 0, 6, 17]), dict(ep=500, ed=500), dict(p=[0.1, 0.5, -0.5, 0.99, -0.01] * 60),
                               dict(z=100, l=1000, bu=list(np.linspace(0.1, 3, 300)), bl=list(-np.linspace(0.2, 2, 300)))])
def test_proj_dual_cone_vs_oracle_random(B, K):
    rng = np.random.RandomState(17)
    m = problems.cone_len(K)
    for trial in range(2):
        x = rng.randn(m) * (10.0 ** rng.randint(-2, 3))
        r_y = np.full(m, 10.0 if trial == 0 else 0.37)
        r_y[:K.get("z", 0)] = 0.01
        if "bu" in K and trial == 1:
            r_y = np.abs(rng.randn(m)) + 0.5
        got = _proj(B, K, x, r_y)
        ref = x.copy(); O.proj_dual_cone(ref, O.ConeWork(K, m), None, r_y)
        assert np.max(np.abs(got - ref)) <= 1e-9 * max(1.0, np.max(np.abs(ref)))


def test_root_plus_known_answers_on_device(B):
    """S/test/problems/test_root_plus.h:68-159 on the CUDA kernel of the ADMM loop (k_rootplus + FinRootPlus)
    through its host-buffer hook; tolerances are the reference test's own."""
    for c in KAT["root_plus"]["cases"]:
        g, p_, mu, r = (np.array(c[k], float) for k in ("g", "p", "mu", "r"))
        got = B.lib.scs_b200_root_plus(B._dptr(g), B._dptr(p_), B._dptr(mu), B._dptr(r), len(g), float(c["tau_scale"]),
                                       float(c["eta"]))
        assert abs(got - c["expected"]) <= 1e-10 * max(1.0, abs(c["expected"])), (len(g), got, c["expected"])
    # a long vector: the grid reduction against numpy (different summation order: 1e-12 relative)
    rng = np.random.RandomState(5)
    nm = 300_001
    g, p_, mu = rng.randn(nm), rng.randn(nm), rng.randn(nm)
    r = np.abs(rng.randn(nm)) + 0.1
    got = B.lib.scs_b200_root_plus(B._dptr(g), B._dptr(p_), B._dptr(mu), B._dptr(r), nm, 10.0, 0.7)
    ref = O.root_plus(g, np.concatenate([r, [10.0]]), np.concatenate([p_, [0.0]]), np.concatenate([mu, [0.0]]), 0.7)
    assert abs(got - ref) <= 1e-11 * max(1.0, abs(ref))


def _psd_inputs(rng, K, kind):
    """Packed inputs for the cones of K: random, all-negative-definite (projection = 0 for the primal cone, i.e.
    proj_dual_cone returns the input's PSD part = 0 ... here the DUAL cone is the PSD cone itself, so a negative
    definite input projects to 0) and rank-deficient (few non-zero eigenvalues of both signs)."""
    parts = []
    for s in K.get("s", []):
        if kind == "random":
            M = rng.randn(s, s); M = M + M.T
        else:
            Q, _ = np.linalg.qr(rng.randn(s, s))
            if kind == "negdef":
                lam = -(np.abs(rng.randn(s)) + 0.1)
            else:  # rank deficient: 3 positive, 2 negative, the rest exactly zero
                lam = np.zeros(s); lam[:5] = [3.0, 1.0, 0.25, -2.0, -0.5][:min(5, s)]
            M = (Q * lam) @ Q.T
            M = 0.5 * (M + M.T)
        v = []
        for j in range(s):  # lower triangle, column-major, off-diagonals * sqrt2 (cones.rst "Semidefinite cones")
            col = M[j:, j].copy(); col[1:] *= math.sqrt(2.0)
            v.append(col)
        parts.append(np.concatenate(v) if v else np.zeros(0))
    for s in K.get("cs", []):
        if kind == "random":
            H = rng.randn(s, s) + 1j * rng.randn(s, s); H = H + H.conj().T
        else:
            Q, _ = np.linalg.qr(rng.randn(s, s) + 1j * rng.randn(s, s))
            if kind == "negdef":
                lam = -(np.abs(rng.randn(s)) + 0.1)
            else:
                lam = np.zeros(s); lam[:5] = [3.0, 1.0, 0.25, -2.0, -0.5][:min(5, s)]
            H = (Q * lam) @ Q.conj().T
            H = 0.5 * (H + H.conj().T)
        v = []
        for j in range(s):  # real diagonal, then (re, im) * sqrt2 of the rows below (cones.c:1088-1095)
            v.append([H[j, j].real])
            below = H[j + 1:, j] * math.sqrt(2.0)
            v.append(np.column_stack([below.real, below.imag]).ravel())
        parts.append(np.concatenate([np.asarray(a, float).ravel() for a in v]) if v else np.zeros(0))
    return np.concatenate(parts)


@pytest.mark.parametrize("K", [dict(s=[97]), dict(s=[120]), dict(s=[200]), dict(s=[257]), dict(cs=[49]), dict(cs=[60]),
                               dict(cs=[130]), dict(s=[225, 201]), dict(cs=[120, 101]), dict(s=[200, 40, 100], cs=[50, 7])])
def test_psd_cluster_path_vs_oracle(B, K):
    """PSD / complex PSD cones whose embedding dimension exceeds 96 run the 4-CTA-cluster Jacobi kernel (the path
    BASELINE.json configs[3], s = 200, takes): projections against the oracle's LAPACK eigendecomposition
    (cones.c:991-1148), cold and warm-started (one workspace projecting a drifting input), on random,
    negative-definite (cones.c:1037-1041: no positive eigenvalue) and rank-deficient inputs, and on mixed sizes
    in one launch (the shared-memory footprint is not monotone in d).  Tolerance 1e-9 relative."""
    rng = np.random.RandomState(11)
    m = problems.cone_len(K)
    k, keep = B.make_cone(K)
    w = B.lib.scs_b200_init_cone(C.byref(k), m)
    assert w
    try:
        x0 = _psd_inputs(rng, K, "random")
        dx = _psd_inputs(rng, K, "random")
        seq = [("cold", x0)] + [("warm%d" % t, x0 + 0.02 * t * dx) for t in (1, 2, 3)]
        seq += [("negdef", _psd_inputs(rng, K, "negdef")), ("rankdef", _psd_inputs(rng, K, "rankdef")),
                ("back", x0 + 0.07 * dx)]
        for name, x in seq:
            got = np.array(x, dtype=np.float64)
            assert B.lib.scs_b200_proj_dual_cone(B._dptr(got), w, None, None) == 0
            ref = x.copy(); O.proj_dual_cone(ref, O.ConeWork(K, m), None, None)
            scale = max(1.0, np.max(np.abs(x)))
            assert np.max(np.abs(got - ref)) <= 1e-9 * scale, (name, K, float(np.max(np.abs(got - ref))))
            if name == "negdef":
                assert np.max(np.abs(got)) <= 1e-9 * scale
    finally:
        B.lib.scs_b200_finish_cone(w)


def test_psd_projection_properties(B):
    """PSD result is PSD, idempotent and the Moreau residual is orthogonal (size-independent)."""
    rng = np.random.RandomState(2)
    s = 120
    n = s * (s + 1) // 2
    x = rng.randn(n)
    p1 = _proj(B, dict(s=[s]), x, None)
    p2 = _proj(B, dict(s=[s]), p1, None)
    assert np.max(np.abs(p1 - p2)) <= 1e-9 * np.max(np.abs(p1))
    M = np.zeros((s, s)); idx = 0
    for j in range(s):
        M[j:, j] = p1[idx:idx + s - j]; idx += s - j
    M = np.tril(M) + np.tril(M, -1).T
    M[np.diag_indices(s)] *= math.sqrt(2.0)
    assert np.linalg.eigvalsh(M).min() >= -1e-9 * np.abs(M).max()
    assert abs(float(p1 @ (x - p1))) <= 1e-9 * float(x @ x)


# ---------------------------------------------------------------------- Anderson acceleration --
@pytest.mark.parametrize("idx", range(len(REF["aa"]["cases"])))
def test_aa_vs_reference_golden(B, idx):
    c = REF["aa"]["cases"][idx]
    M, bv, x = np.array(c["M"]), np.array(c["b"]), np.array(c["x0"])
    w = B.lib.scs_b200_aa_init(c["dim"], c["mem"], c["mem"], c["type1"], 1e-8, c["relaxation"], 1.0, 1e10, 5)
    assert w
    norms, sgs = [], []
    for _ in range(c["iters"]):
        f = M @ x + bv
        norms.append(B.lib.scs_b200_aa_apply(B._dptr(f), B._dptr(x), w))
        fn = M @ f + bv
        xn = f.copy()
        sgs.append(B.lib.scs_b200_aa_safeguard(B._dptr(fn), B._dptr(xn), w))
        x = xn
    st = B.lib.scs_b200_aa_get_stats(w)
    B.lib.scs_b200_aa_finish(w)
    assert sgs == c["safeguards"]
    assert np.allclose(norms, c["aa_norms"], rtol=1e-4, atol=1e-8)
    assert np.max(np.abs(x - np.array(c["x_final"]))) <= 1e-7 * max(1.0, np.max(np.abs(x)))
    assert st.n_accept == sum(1 for v in c["aa_norms"] if v > 0)


def test_aa_large_dim_vs_oracle(B):
    rng = np.random.RandomState(8)
    dim, mem = 100_003, 10
    d = rng.rand(dim) * 0.95
    bv = rng.randn(dim)
    w = B.lib.scs_b200_aa_init(dim, mem, mem, 1, 1e-8, 1.0, 1.0, 1e10, 5)
    a = O.AaWork(dim, mem, mem, 1, 1e-8, 1.0)
    x1 = rng.randn(dim); x2 = x1.copy()
    for it in range(25):
        f1, f2 = d * x1 + bv, d * x2 + bv
        n1 = B.lib.scs_b200_aa_apply(B._dptr(f1), B._dptr(x1), w)
        n2 = a.apply(f2, x2)
        assert (n1 > 0) == (n2 > 0)
        assert np.max(np.abs(f1 - f2)) <= 1e-6 * max(1.0, np.max(np.abs(f2)))
        x1, x2 = f1, f2
    B.lib.scs_b200_aa_finish(w)


# -------------------------------------------------------------------------------- full solves --
def _case_data(c):
    if "Ax" in c:
        return helpers.problem_from_record(c)
    data, _ = problems.gen_feasible(c["cone"], c["n"], c["density"], c["seed"], with_P=c["with_P"])
    return data, c["cone"]


@pytest.mark.parametrize("name", [c["name"] for c in REF["solves"]["cases"]])
def test_solve_vs_reference_golden(scsb, name):
    c = [c for c in REF["solves"]["cases"] if c["name"] == name][0]
    data, K = _case_data(c)
    if name in ("infeasible", "unbounded"):
        r = c["runs"]["qdldl_1e-07"]
        got = scsb.SCS(data, K, verbose=False, eps_abs=1e-7, eps_rel=1e-7).solve()
        assert got["info"]["status_val"] == r["status_val"], (got["info"]["status"], r["status"])
        if name == "infeasible":   # certificate: b'y = -1, ||A'y|| small (scs.c:818-829)
            y = got["y"]
            assert abs(float(data["b"] @ y) + 1.0) <= 1e-6 and np.max(np.abs(data["A"].T @ y)) <= 1e-5
            assert np.all(np.isnan(got["x"])) and math.isinf(got["info"]["pobj"])
        else:
            x = got["x"]
            assert abs(float(data["c"] @ x) + 1.0) <= 1e-6 and np.all(np.isnan(got["y"]))
        return
    for eps in (1e-4, 1e-9):
        ri = c["runs"]["cpu_indirect_%g" % eps]
        rq = c["runs"]["qdldl_%g" % eps]
        if ri["status_val"] != 1:   # the reference's indirect backend itself runs out of iterations here
            continue
        got = scsb.SCS(data, K, verbose=False, eps_abs=eps, eps_rel=eps, max_iters=100000).solve()
        gi = got["info"]
        assert gi["status_val"] == ri["status_val"] == 1, (gi["status"], ri["status"])
        helpers.verify_solution(data, K, got, eps, eps)
        if eps == 1e-9:
            for key in ("pobj", "dobj"):
                assert abs(gi[key] - rq[key]) <= 1e-6 * max(1.0, abs(rq[key])), (key, gi[key], rq[key])
            if rq.get("x") is not None:
                assert np.max(np.abs(got["x"] - np.array(rq["x"], float))) <= 1e-5 * max(1.0, np.max(np.abs(got["x"])))
        else:
            assert abs(gi["pobj"] - rq["pobj"]) <= 5e-3 * max(1.0, abs(rq["pobj"]))


@pytest.mark.parametrize("name", [p["name"] for p in KAT["problems"]] + [p["name"] for p in KAT["file_problems"]])
def test_known_objectives(scsb, name):
    """objective constants of the reference's own C tests (S/test/problems/*.h), tolerance 1e-4."""
    p = [p for p in KAT["problems"] + KAT["file_problems"] if p["name"] == name][0]
    data, K = helpers.problem_from_record(p)
    st = dict(eps_abs=1e-6, eps_rel=1e-6)
    st.update({k: v for k, v in p.get("settings", {}).items() if k in ("eps_abs", "eps_rel", "eps_infeas")})
    got = scsb.SCS(data, K, verbose=False, **st).solve()
    gi = got["info"]
    assert gi["status_val"] == 1, gi["status"]
    tol = 1e-4 * max(1.0, abs(p["opt"]))
    assert abs(gi["pobj"] - p["opt"]) < tol and abs(gi["dobj"] - p["opt"]) < tol, (gi["pobj"], gi["dobj"], p["opt"])
    helpers.verify_solution(data, K, got, st["eps_abs"], st["eps_rel"])


def test_warm_start_and_update(scsb):
    p = [p for p in KAT["problems"] if p["name"] == "hs21_tiny_qp"][0]
    data, K = helpers.problem_from_record(p)
    solver = scsb.SCS(data, K, verbose=False, eps_abs=1e-6, eps_rel=1e-6, eps_infeas=0.0)
    first = solver.solve(warm_start=False)
    assert first["info"]["status_val"] == 1
    again = solver.solve()                       # warm start from the stored solution
    assert again["info"]["status_val"] == 1 and again["info"]["iter"] <= 25   # hs21_tiny_qp.h:91
    third = solver.solve(warm_start=True, x=first["x"], y=first["y"], s=first["s"])
    assert third["info"]["iter"] <= 25
    # scs_update: new b, c then solve == fresh solve on the new data (lp_update.h)
    K2 = dict(z=3, l=10)
    d2, _ = problems.gen_feasible(K2, 8, 0.5, seed=12)
    s2 = scsb.SCS(d2, K2, verbose=False, eps_abs=1e-9, eps_rel=1e-9)
    s2.solve(warm_start=False)
    d3, _ = problems.gen_feasible(K2, 8, 0.5, seed=12)
    rng = np.random.RandomState(0)
    d3["c"] = d3["c"] + 0.05 * rng.randn(8)
    d3["b"] = d3["b"] + 0.05 * np.abs(rng.randn(13))
    s2.update(b=d3["b"], c=d3["c"])
    upd = s2.solve(warm_start=False)
    fresh = scsb.SCS(d3, K2, verbose=False, eps_abs=1e-9, eps_rel=1e-9).solve(warm_start=False)
    assert upd["info"]["status_val"] == fresh["info"]["status_val"] == 1
    assert abs(upd["info"]["pobj"] - fresh["info"]["pobj"]) <= 1e-7 * max(1.0, abs(fresh["info"]["pobj"]))
    s2.update(b=None, c=d3["c"])                 # partial update keeps the other vector
    assert s2.solve()["info"]["status_val"] == 1


def test_settings_and_flow(scsb):
    K = dict(z=4, l=12, q=[3, 5], ep=2)
    data, p_star = problems.gen_feasible(K, 20, 0.3, seed=5)
    one = scsb.SCS(data, K, verbose=False, max_iters=1).solve()["info"]          # test_inaccurate.h
    assert one["iter"] == 1 and one["status_val"] in (2, -6, -7) and "max_iters" in one["status"]
    base = scsb.SCS(data, K, verbose=False, eps_abs=1e-8, eps_rel=1e-8).solve()["info"]
    assert base["status_val"] == 1 and abs(base["pobj"] - p_star) <= 1e-5 * max(1, abs(p_star))
    noaa = scsb.SCS(data, K, verbose=False, eps_abs=1e-8, eps_rel=1e-8, acceleration_lookback=0).solve()["info"]
    assert noaa["status_val"] == 1 and noaa["accepted_accel_steps"] == 0 and noaa["aa_stats"]["n_accept"] == 0
    assert math.isnan(noaa["aa_stats"]["last_aa_norm"])                          # test_solver_options.h
    for kw in (dict(acceleration_type_1=0), dict(normalize=False), dict(adaptive_scale=False), dict(alpha=1.0),
               dict(acceleration_interval=5, acceleration_lookback=5), dict(acceleration_relaxation=0.8),
               dict(rho_x=1e-3, scale=1.0)):
        r = scsb.SCS(data, K, verbose=False, eps_abs=1e-8, eps_rel=1e-8, **kw).solve()["info"]
        assert r["status_val"] == 1, (kw, r["status"])
        assert abs(r["pobj"] - p_star) <= 1e-5 * max(1, abs(p_star)), kw
    K2 = dict(z=0, l=300, q=[10] * 24, ep=20)
    d2, _ = problems.gen_feasible(K2, 200, 0.05, seed=1234)
    ad = scsb.SCS(d2, K2, verbose=False, eps_abs=1e-9, eps_rel=1e-9).solve()["info"]
    assert ad["scale_updates"] >= 1                                               # adaptive scale fires
    tl = scsb.SCS(d2, K2, verbose=False, eps_abs=1e-14, eps_rel=1e-14, time_limit_secs=0.05).solve()["info"]
    assert "time_limit" in tl["status"]
    v = scsb.SCS(d2, K2, verbose=True, eps_abs=1e-4, eps_rel=1e-4).solve()["info"]  # printing path
    assert v["status_val"] == 1 and v["lin_sys_time"] > 0 and v["cone_time"] > 0


def test_edge_cases(scsb):
    # empty cones inside the lists, an all-zero column, n = 1
    K = dict(z=1, l=2, q=[0, 3, 1], s=[0, 2, 1], cs=[0, 2], ep=1, ed=0, p=[])
    data, _ = problems.gen_feasible(K, 5, 0.6, seed=2)
    r = scsb.SCS(data, K, verbose=False, eps_abs=1e-8, eps_rel=1e-8).solve()
    o = O.ScsOracle(data, K, eps_abs=1e-8, eps_rel=1e-8).solve()
    assert r["info"]["status_val"] == o["info"]["status_val"]
    assert abs(r["info"]["pobj"] - o["info"]["pobj"]) <= 1e-6 * max(1.0, abs(o["info"]["pobj"]))
    A = sp.csc_matrix(np.array([[1.0], [-1.0]]))
    tiny = scsb.SCS(dict(A=A, b=np.array([2.0, -1.0]), c=np.array([1.0])), dict(l=2), verbose=False,
                    eps_abs=1e-9, eps_rel=1e-9).solve()
    assert tiny["info"]["status_val"] == 1 and abs(tiny["x"][0] - 1.0) <= 1e-6      # min x s.t. 1 <= x <= 2
    with pytest.raises(ValueError, match="ScsWork allocation error"):
        scsb.SCS(dict(A=A, b=np.array([2.0, -1.0]), c=np.array([1.0])), dict(l=3), verbose=False)  # cone dims mismatch


def test_lasso_large_properties(scsb):
    """BASELINE workload family at 1/20 scale: the returned point must satisfy the reference's
    convergence criteria when everything is recomputed on the host with scipy."""
    from scs_python_b200 import problems as P
    data, cone, _ = P.lasso(50_000, 100_000, 100, seed=4)     # n = m = 200k, nnz(A) = 5.3M
    solver = scsb.SCS(data, cone, verbose=False, max_iters=5000)
    sol = solver.solve(warm_start=False)
    assert sol["info"]["status_val"] == 1, sol["info"]["status"]
    helpers.verify_solution(data, cone, sol, 1e-4, 1e-4, cone_tol=1e-6)
    st = solver._solver.stats()
    assert st["kernel_launches"] > 0 and st["cg_iters"] > 0 and st["algorithmic_bytes"] > 0


def test_iteration_marks(scsb):
    from scs_python_b200 import problems as P
    data, cone, _ = P.lasso(2000, 4000, 20, seed=1)
    s = scsb.SCS(data, cone, verbose=False, eps_abs=0.0, eps_rel=0.0, eps_infeas=0.0, max_iters=100)
    s._solver.set_marks(25, 75)
    r = s.solve(warm_start=False)
    mk = s._solver.get_marks()
    assert r["info"]["iter"] == 100 and mk["iters"] == 50 and mk["ms"] > 0 and mk["kernel_launches"] > 50
    assert mk["spmv_g_launches"] == mk["cg_iters"]           # one fused G launch per real CG iteration
    assert mk["spmv_a_launches"] == mk["cg_iters"] + 50      # + the warm-start A s product of every ADMM iteration


def test_graph_and_stream_paths_agree(scsb, monkeypatch):
    """The graph-launched iteration (CG loop = WHILE node, no host sync between residual checks)
    and the stream-launched one run the same kernels in the same order: identical results."""
    K = dict(z=3, l=10, q=[4, 6], s=[3], ep=2, p=[0.4])
    data, _ = problems.gen_feasible(K, 30, 0.3, seed=11, with_P=True)
    kw = dict(eps_abs=1e-8, eps_rel=1e-8, max_iters=20000, verbose=False)
    a = scsb.SCS(data, K, **kw).solve()
    monkeypatch.setenv("SCS_B200_NO_GRAPH", "1")
    b = scsb.SCS(data, K, **kw).solve()
    monkeypatch.delenv("SCS_B200_NO_GRAPH")
    assert a["info"]["status_val"] == b["info"]["status_val"] == 1
    assert a["info"]["iter"] == b["info"]["iter"]
    assert np.array_equal(a["x"], b["x"]) and np.array_equal(a["y"], b["y"]) and np.array_equal(a["s"], b["s"])
    for k in ("lin_sys_time", "cone_time", "accel_time"):
        assert a["info"][k] > 0 and b["info"][k] > 0


# ---------------------------------------------------------------- the live compiled reference --
def test_against_live_reference(scsb, ref_scs):
    scs = ref_scs
    for seed, K, n, withP in [(31, dict(z=6, l=30, q=[5, 7], ep=3), 25, False),
                              (32, dict(z=3, l=20, s=[4, 6], ed=2, p=[0.6]), 30, True)]:
        data, _ = problems.gen_feasible(K, n, 0.3, seed, with_P=withP)
        kw = dict(eps_abs=1e-9, eps_rel=1e-9, max_iters=100000, verbose=False)
        ref = scs.SCS(data, K, linear_solver=scs.LinearSolver.CPU_INDIRECT, **kw).solve()
        got = scsb.SCS(data, K, **kw).solve()
        assert got["info"]["status_val"] == ref["info"]["status_val"] == 1
        for key in ("pobj", "dobj"):
            assert abs(got["info"][key] - ref["info"][key]) <= 1e-6 * max(1.0, abs(ref["info"][key]))
        assert set(got["info"].keys()) >= set(ref["info"].keys())        # info-dict key set, test_scs_coverage.py:352-366


def test_reference_front_end_drives_b200(ref_scs, gpu):
    """The reference's UNMODIFIED scspy.c front end, linked against libscsb200.so (oracle/Makefile
    ref_frontend, module name _scs_gpu), selected by the reference's own LinearSolver enum."""
    scs = ref_scs
    try:
        from scs import _scs_gpu  # noqa: F401
    except ImportError:
        pytest.skip("reference front end not linked on this box")
    K = dict(z=4, l=12, q=[3, 5], s=[3], ep=2)
    data, p_star = problems.gen_feasible(K, 20, 0.3, seed=6)
    sol = scs.SCS(data, K, linear_solver=scs.LinearSolver.GPU_INDIRECT, verbose=False, eps_abs=1e-9, eps_rel=1e-9).solve()
    assert sol["info"]["status"] == "solved"
    assert abs(sol["info"]["pobj"] - p_star) <= 1e-6 * max(1.0, abs(p_star))
