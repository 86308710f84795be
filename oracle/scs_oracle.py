"""CPU restatement (numpy) of the SCS ADMM hot path -- TEST INFRASTRUCTURE ONLY.

This module is the *oracle* for the B200 backend: a plain numpy/scipy restatement of the
reference algorithm (bodono/scs-python @ d7626bf, scs_source = cvxgrp/scs v3.2.11).  It is
imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg; the
product path (scs_python_b200/) never imports it and has no CPU fallback.

Parity status: PINNED.  tests/test_oracle.py checks this restatement against
  * the reference's own golden vectors (S/test/problems/test_exp_cone.h:52-77,
    S/test/problems/test_root_plus.h:68-159, objective constants of hs21_tiny_qp.h:39,
    test_soc_sizes.h, test_zero_cone.h, test_box_cone.h, test_power_cone.h ...), committed
    as tests/golden/*.json by tests/golden/make_golden.py, and
  * the compiled reference itself (oracle/_ref, built by oracle/Makefile) on seeded problems.

Every function cites the reference file:line it follows ("S/" = scs_source/).
All arithmetic is float64 (scs_types.h:27-33); indices int32 in the GPU build.
"""
from __future__ import annotations

import math
import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp

# ----------------------------------------------------------------------------------------
# constants  (S/include/glbopts.h:35-50,184-257 )
# ----------------------------------------------------------------------------------------
MAX_ITERS, EPS_REL, EPS_ABS, EPS_INFEAS = 100000, 1e-4, 1e-4, 1e-7
ALPHA, RHO_X, SCALE = 1.5, 1e-6, 0.1
FEASIBLE_ITERS, RESCALING_MIN_ITERS = 1, 100
CONVERGED_INTERVAL, PRINT_INTERVAL = 25, 250
DIV_EPS_TOL = 1e-18
ITERATE_NORM, TAU_FACTOR = 1.0, 10.0
INFEAS_NEGATIVITY_TOL = 1e-9
AA_SAFEGUARD_FACTOR, AA_MAX_WEIGHT_NORM, AA_IR_MAX_STEPS = 1.0, 1e10, 5
MAX_SCALE_VALUE, MIN_SCALE_VALUE = 1e6, 1e-6
CG_BEST_TOL, CG_TOL_FACTOR, CG_RATE = 1e-12, 0.2, 1.5
MIN_NORMALIZATION_FACTOR, MAX_NORMALIZATION_FACTOR = 1e-4, 1e4
NUM_RUIZ_PASSES, NUM_L2_PASSES = 25, 1
BOX_CONE_MAX_ITERS, POW_CONE_TOL, POW_CONE_MAX_ITERS, MAX_BOX_VAL = 25, 1e-9, 20, 1e15
EXP_CONE_INFINITY_VALUE = 1e15

SCS_INFEASIBLE_INACCURATE, SCS_UNBOUNDED_INACCURATE, SCS_SIGINT, SCS_FAILED = -7, -6, -5, -4
SCS_INDETERMINATE, SCS_INFEASIBLE, SCS_UNBOUNDED, SCS_UNFINISHED = -3, -2, -1, 0
SCS_SOLVED, SCS_SOLVED_INACCURATE = 1, 2

DEFAULT_SETTINGS = dict(
    normalize=1, scale=SCALE, adaptive_scale=1, rho_x=RHO_X, max_iters=MAX_ITERS,
    eps_abs=EPS_ABS, eps_rel=EPS_REL, eps_infeas=EPS_INFEAS, alpha=ALPHA,
    time_limit_secs=0.0, verbose=0, warm_start=0, acceleration_lookback=10,
    acceleration_interval=10, acceleration_type_1=1, acceleration_regularization=1e-8,
    acceleration_relaxation=1.0)


def safediv_pos(x, y):
    """SAFEDIV_POS, glbopts.h:194-196."""
    return x / DIV_EPS_TOL if y < DIV_EPS_TOL else x / y


def norm_inf(x):
    return float(np.max(np.abs(x))) if len(x) else 0.0


# ----------------------------------------------------------------------------------------
# cone bookkeeping  (S/src/cones.c:342-424)
# ----------------------------------------------------------------------------------------
def cone_dict(k):
    """Normalise a cone dict to the ScsCone field set (scs.h:120-173)."""
    def lst(v):
        if v is None:
            return []
        if np.isscalar(v):
            return [int(v)]
        return list(v)
    out = dict(z=int(k.get("z", k.get("f", 0)) or 0), l=int(k.get("l", 0) or 0),
               bu=np.asarray(k.get("bu", []) if k.get("bu", None) is not None else [], dtype=float).copy(),
               bl=np.asarray(k.get("bl", []) if k.get("bl", None) is not None else [], dtype=float).copy(),
               q=[int(v) for v in lst(k.get("q"))], s=[int(v) for v in lst(k.get("s"))],
               cs=[int(v) for v in lst(k.get("cs"))],
               ep=int(k.get("ep", 0) or 0), ed=int(k.get("ed", 0) or 0),
               p=[float(v) for v in (k.get("p") if k.get("p", None) is not None else [])])
    out["bsize"] = (len(out["bu"]) + 1) if len(out["bu"]) > 0 else int(k.get("bsize", 0) or 0)
    return out


def cone_boundaries(k):
    """set_cone_boundaries, cones.c:386-424: [z+l+bsize, q..., s(s+1)/2..., cs^2..., 3...]."""
    b = [k["z"] + k["l"] + k["bsize"]]
    b += list(k["q"])
    b += [s * (s + 1) // 2 for s in k["s"]]
    b += [c * c for c in k["cs"]]
    b += [3] * (k["ep"] + k["ed"] + len(k["p"]))
    return b


def cone_dims(k):
    return sum(cone_boundaries(k))


def enforce_cone_boundaries(bounds, vec, f):
    """cones.c:366-379: aggregate `vec` inside every cone of size > 1 with `f`."""
    count = bounds[0]
    for delta in bounds[1:]:
        if delta > 0:
            vec[count:count + delta] = f(vec[count:count + delta])
        count += delta


def set_r_y(k, m, scale):
    """cones.c:349-363."""
    r_y = np.full(m, 1.0 / scale)
    r_y[:k["z"]] = 1.0 / (1000.0 * scale)
    return r_y


# ----------------------------------------------------------------------------------------
# equilibration  (S/linsys/scs_matrix.c:203-470, S/src/normalize.c:33-90)
# ----------------------------------------------------------------------------------------
def _apply_limit(x):
    x = np.where(x < MIN_NORMALIZATION_FACTOR, 1.0, x)
    return np.where(x > MAX_NORMALIZATION_FACTOR, MAX_NORMALIZATION_FACTOR, x)


def _inv_sqrt_limited(x):
    x = np.sqrt(_apply_limit(x))
    return np.where(x < DIV_EPS_TOL, 1.0 / DIV_EPS_TOL, 1.0 / x)


def _col_reduce(M, fn_abs, n):
    """per-column reduction over a CSC matrix (helper)."""
    out = np.zeros(n)
    nz = np.diff(M.indptr) > 0
    if M.nnz:
        red = fn_abs(M.data, M.indptr[:-1][nz])
        out[nz] = red
    return out


def normalize_a_p(A, P, k):
    """normalize_a_p, scs_matrix.c:407-470.  A: csc (m x n), P: upper-tri csc or None.
    Returns (A_scaled, P_scaled, D, E); inputs are not modified."""
    A = A.tocsc().copy().astype(float)
    A.sort_indices()
    m, n = A.shape
    P = None if P is None else sp.triu(P, format="csc").astype(float)
    bounds = cone_boundaries(k)
    D = np.ones(m)
    E = np.ones(n)
    rows = A.indices
    cols = np.repeat(np.arange(n), np.diff(A.indptr))
    if P is not None:
        P.sort_indices()
        prow = P.indices
        pcol = np.repeat(np.arange(n), np.diff(P.indptr))

    def rescale(Dt, Et):  # scs_matrix.c:344-381
        nonlocal D, E
        A.data *= Dt[rows] * Et[cols]
        if P is not None:
            P.data *= Et[prow] * Et[pcol]
        D = D * Dt
        E = E * Et

    for _ in range(NUM_RUIZ_PASSES):  # compute_ruiz_mats, scs_matrix.c:210-277
        Dt = np.zeros(m)
        np.maximum.at(Dt, rows, np.abs(A.data))
        enforce_cone_boundaries(bounds, Dt, lambda v: norm_inf(v))
        Dt = _inv_sqrt_limited(Dt)
        Et = np.zeros(n)
        if P is not None:
            np.maximum.at(Et, pcol, np.abs(P.data))
            np.maximum.at(Et, prow, np.abs(P.data))
        np.maximum.at(Et, cols, np.abs(A.data))
        Et = _inv_sqrt_limited(Et)
        rescale(Dt, Et)
    for _ in range(NUM_L2_PASSES):  # compute_l2_mats, scs_matrix.c:279-342
        Dt = np.zeros(m)
        np.add.at(Dt, rows, A.data * A.data)
        Dt = np.sqrt(Dt)
        enforce_cone_boundaries(bounds, Dt, lambda v: float(np.sum(v)) / len(v))  # SCS(mean)
        Dt = _inv_sqrt_limited(Dt)
        Et = np.zeros(n)
        if P is not None:
            sq = P.data * P.data
            np.add.at(Et, pcol, sq)
            off = prow != pcol
            np.add.at(Et, prow[off], sq[off])
        np.add.at(Et, cols, A.data * A.data)
        Et = _inv_sqrt_limited(np.sqrt(Et))
        rescale(Dt, Et)
    return A, P, D, E


def normalize_b_c(D, E, b, c):
    """normalize.c:33-61. returns (b, c, sigma) with primal_scale = dual_scale = sigma."""
    c = c * E
    b = b * D
    sigma = max(norm_inf(c), norm_inf(b))
    sigma = 1.0 if sigma < MIN_NORMALIZATION_FACTOR else sigma
    sigma = MAX_NORMALIZATION_FACTOR if sigma > MAX_NORMALIZATION_FACTOR else sigma
    sigma = safediv_pos(1.0, sigma)
    return b * sigma, c * sigma, sigma


# ----------------------------------------------------------------------------------------
# cone projections  (S/src/cones.c:986-1588, S/src/exp_cone.c)
# ----------------------------------------------------------------------------------------
def proj_soc(x):
    """cones.c:1242-1271 (in place)."""
    q = len(x)
    if q <= 0:
        return
    if q == 1:
        x[0] = max(x[0], 0.0)
        return
    v1 = x[0]
    s = float(np.sqrt(np.dot(x[1:], x[1:]))) if q > 2 else abs(x[1])
    alpha = (s + v1) / 2.0
    if s <= v1:
        return
    if s <= -v1:
        x[:] = 0.0
    else:
        x[0] = alpha
        x[1:] *= alpha / s


def proj_psd(X, n):
    """proj_semi_definite_cone, cones.c:991-1059.  X: packed lower-tri col-major, off-diag*sqrt2."""
    if n == 0:
        return
    if n == 1:
        X[0] = max(X[0], 0.0)
        return
    M = np.zeros((n, n))
    idx = 0
    for j in range(n):
        M[j:, j] = X[idx:idx + n - j]
        idx += n - j
    M[np.diag_indices(n)] *= math.sqrt(2.0)          # cones.c:1016-1017
    M = np.tril(M) + np.tril(M, -1).T
    e, Z = np.linalg.eigh(M)                          # dsyevr 'V','A','L'
    pos = e > 0
    if not pos.any():
        X[:] = 0.0
        return
    Zp = Z[:, pos] * np.sqrt(e[pos])
    M = Zp @ Zp.T                                     # dsyrk
    M[np.diag_indices(n)] /= math.sqrt(2.0)
    idx = 0
    for j in range(n):
        X[idx:idx + n - j] = M[j:, j]
        idx += n - j


def proj_cpsd(X, n):
    """proj_complex_semi_definite_cone, cones.c:1064-1148. X has n*n reals: per column j:
    real diagonal, then (re,im) pairs of the strictly-lower part."""
    if n == 0:
        return
    if n == 1:
        X[0] = max(X[0], 0.0)
        return
    M = np.zeros((n, n), dtype=complex)
    for i in range(n):
        base = i * (2 * n - i)
        M[i, i] = X[base]
        rest = X[base + 1: base + 1 + 2 * (n - i - 1)]
        M[i + 1:, i] = rest[0::2] + 1j * rest[1::2]
    M[np.diag_indices(n)] *= math.sqrt(2.0)
    M = np.tril(M) + np.tril(M, -1).conj().T
    e, Z = np.linalg.eigh(M)
    pos = e > 0
    if not pos.any():
        X[:] = 0.0
        return
    Zp = Z[:, pos] * np.sqrt(e[pos])
    M = Zp @ Zp.conj().T
    M[np.diag_indices(n)] /= math.sqrt(2.0)
    for i in range(n):
        base = i * (2 * n - i)
        X[base] = M[i, i].real
        col = M[i + 1:, i]
        X[base + 1: base + 1 + 2 * (n - i - 1)][0::2] = col.real
        X[base + 1: base + 1 + 2 * (n - i - 1)][1::2] = col.imag


def proj_box_cone(tx, bl, bu, t_wm, r_box):
    """proj_box_cone, cones.c:1174-1237 (in place on tx=[t;s]); returns t."""
    bsize = len(tx)
    if bsize == 1:
        tx[0] = max(tx[0], 0.0)
        return tx[0]
    x = tx[1:]
    t = t_wm
    if r_box is not None:
        rho_t = 1.0 / r_box[0]
        rinv = 1.0 / r_box[1:]
    else:
        rho_t = 1.0
        rinv = np.ones(bsize - 1)
    with np.errstate(invalid="ignore"):
        for _ in range(BOX_CONE_MAX_ITERS):
            t_prev = t
            gt = rho_t * (t - tx[0])
            ht = rho_t
            up = x > t * bu
            lo = (~up) & (x < t * bl)
            gt += float(np.sum(rinv[up] * (t * bu[up] - x[up]) * bu[up]))
            ht += float(np.sum(rinv[up] * bu[up] * bu[up]))
            gt += float(np.sum(rinv[lo] * (t * bl[lo] - x[lo]) * bl[lo]))
            ht += float(np.sum(rinv[lo] * bl[lo] * bl[lo]))
            t = max(t - gt / max(ht, 1e-8), 0.0)
            if abs(gt / max(ht, 1e-6)) < 1e-12 * max(t, 1.0) or abs(t - t_prev) < 1e-11 * max(t, 1.0):
                break
        up = x > t * bu
        lo = (~up) & (x < t * bl)
        x[up] = t * bu[up]
        x[lo] = t * bl[lo]
    tx[0] = t
    return t


def _pow_calc_x(r, xh, rh, a):
    x = 0.5 * (xh + math.sqrt(xh * xh + 4 * a * (rh - r) * r))
    return max(x, 1e-12)


def proj_power_cone(v, a):
    """cones.c:1282-1324 (in place)."""
    xh, yh, rh = v[0], v[1], abs(v[2])
    if xh >= 0 and yh >= 0 and POW_CONE_TOL + xh ** a * yh ** (1 - a) >= rh:
        return
    if xh <= 0 and yh <= 0 and POW_CONE_TOL + (-xh) ** a * (-yh) ** (1 - a) >= rh * a ** a * (1 - a) ** (1 - a):
        v[:] = 0.0
        return
    r = rh / 2
    x = y = 0.0
    for _ in range(POW_CONE_MAX_ITERS):
        x = _pow_calc_x(r, xh, rh, a)
        y = _pow_calc_x(r, yh, rh, 1 - a)
        xa = x ** a
        y1a = y ** (1 - a)
        f = xa * y1a - r
        if abs(f) < POW_CONE_TOL:
            break
        dxdr = a * (rh - 2 * r) / (2 * x - xh)
        dydr = (1 - a) * (rh - 2 * r) / (2 * y - yh)
        fp = xa * y1a * (a * dxdr / x + (1 - a) * dydr / y) - 1
        r = max(r - f / fp, 0)
        r = min(r, rh)
    v[0], v[1] = x, y
    v[2] = -r if v[2] < 0 else r


# ---- exponential cone, Friberg 2021 (exp_cone.c) ----
def _exp(x):
    try:
        return math.exp(x)
    except OverflowError:
        return math.inf


def _isfinite(x):
    return abs(x) < EXP_CONE_INFINITY_VALUE


def _hfun_f(v0, rho):  # exp_cone.c:41-48
    r0, s0, t0 = v0
    e = _exp(rho)
    en = 1.0 / e if e != 0 else math.inf
    return ((rho - 1) * r0 + s0) * e - (r0 - rho * s0) * en - (rho * (rho - 1) + 1) * t0


def _hfun_fd(v0, rho):  # exp_cone.c:50-62
    r0, s0, t0 = v0
    e = _exp(rho)
    en = 1.0 / e if e != 0 else math.inf
    f = ((rho - 1) * r0 + s0) * e - (r0 - rho * s0) * en - (rho * (rho - 1) + 1) * t0
    df = (rho * r0 + s0) * e + (r0 - (rho - 1) * s0) * en - (2 * rho - 1) * t0
    return f, df


def _root_search_binary(v0, xl, xu, x):  # exp_cone.c:65-95
    x_plus = x
    for _ in range(40):
        f = _hfun_f(v0, x)
        if f < 0.0:
            xl = x
        else:
            xu = x
        x_plus = 0.5 * (xl + xu)
        if abs(x_plus - x) <= 1e-12 * max(1.0, abs(x_plus)) or x_plus == xl or x_plus == xu:
            break
        x = x_plus
    return x_plus


def _root_search_newton(v0, xl, xu, x):  # exp_cone.c:98-162
    EPS, DFTOL, MAXITER, LODAMP, HIDAMP = 1e-15, 1e-13, 20, 0.05, 0.95
    i = 0
    converged = False
    while i < MAXITER:
        f, df = _hfun_fd(v0, x)
        if abs(f) <= EPS:
            converged = True
            break
        if f < 0.0:
            xl = x
        else:
            xu = x
        if xu <= xl:
            xu = 0.5 * (xu + xl)
            xl = xu
            converged = True
            break
        if (not _isfinite(f)) or df < DFTOL:
            converged = True
            break
        x_plus = x - f / df
        if abs(x_plus - x) <= EPS * max(1.0, abs(x_plus)):
            converged = True
            break
        if x_plus >= xu:
            x = min(LODAMP * x + HIDAMP * xu, xu)
        elif x_plus <= xl:
            x = max(LODAMP * x + HIDAMP * xl, xl)
        else:
            x = x_plus
        i += 1
    if converged:
        return max(xl, min(xu, x))
    return _root_search_binary(v0, xl, xu, x)


def _nds3(a, b):
    d0, d1, d2 = a[0] - b[0], a[1] - b[1], a[2] - b[2]
    return d0 * d0 + d1 * d1 + d2 * d2


def _primal_heur(v0):  # exp_cone.c:165-188
    r0, s0, t0 = v0
    vp = [min(r0, 0.0), 0.0, max(t0, 0.0)]
    d = _nds3(v0, vp)
    if s0 > 0.0:
        tp = max(t0, s0 * _exp(r0 / s0))
        nd = (tp - t0) * (tp - t0)
        if nd < d:
            vp = [r0, s0, tp]
            d = nd
    return d, vp


def _polar_heur(v0):  # exp_cone.c:191-214
    r0, s0, t0 = v0
    vd = [0.0, min(s0, 0.0), min(t0, 0.0)]
    d = _nds3(v0, vd)
    if r0 > 0.0:
        td = min(t0, -r0 * _exp(s0 / r0 - 1.0))
        nd = (t0 - td) * (t0 - td)
        if nd < d:
            vd = [r0, s0, td]
            d = nd
    return d, vd


def _ppsi(v0):  # exp_cone.c:216-227
    r0, s0 = v0[0], v0[1]
    if r0 > s0:
        psi = (r0 - s0 + math.sqrt(r0 * r0 + s0 * s0 - r0 * s0)) / r0
    else:
        psi = -s0 / (r0 - s0 - math.sqrt(r0 * r0 + s0 * s0 - r0 * s0))
    return ((psi - 1.0) * r0 + s0) / (psi * (psi - 1.0) + 1.0)


def _pomega(rho):
    val = _exp(rho) / (rho * (rho - 1.0) + 1.0)
    if rho < 2.0:
        val = min(val, math.exp(2.0) / 3.0)
    return val


def _dpsi(v0):  # exp_cone.c:238-249
    r0, s0 = v0[0], v0[1]
    if s0 > r0:
        psi = (r0 - math.sqrt(r0 * r0 + s0 * s0 - r0 * s0)) / s0
    else:
        psi = (r0 - s0) / (r0 + math.sqrt(r0 * r0 + s0 * s0 - r0 * s0))
    return (r0 - psi * s0) / (psi * (psi - 1.0) + 1.0)


def _domega(rho):
    val = -_exp(-rho) / (rho * (rho - 1.0) + 1.0)
    if rho > -1.0:
        val = max(val, -math.exp(1.0) / 3.0)
    return val


def _clip(x, l, u):
    return max(l, min(u, x))


def _exp_search_bracket(v0, pdist_sq, ddist_sq):  # exp_cone.c:261-323
    r0, s0, t0 = v0
    baselow, baseupr = -EXP_CONE_INFINITY_VALUE, EXP_CONE_INFINITY_VALUE
    low, upr = -EXP_CONE_INFINITY_VALUE, EXP_CONE_INFINITY_VALUE
    Dp = math.sqrt(max(pdist_sq - min(s0, 0.0) * min(s0, 0.0), 0.0))
    Dd = math.sqrt(max(ddist_sq - min(r0, 0.0) * min(r0, 0.0), 0.0))
    if t0 > 0.0:
        low = max(low, math.log(t0 / _ppsi(v0)))
    elif t0 < 0.0:
        upr = min(upr, -math.log(-t0 / _dpsi(v0)))
    if r0 > 0.0:
        baselow = 1.0 - s0 / r0
        low = max(low, baselow)
        tpu = max(1e-12, min(Dd, Dp + t0))
        val = r0 * _pomega(low)
        sgn = -1 if val < 0 else 1
        upr = min(upr, max(low, baselow + safediv_pos(tpu, abs(val)) * sgn))
    if s0 > 0.0:
        baseupr = r0 / s0
        upr = min(upr, baseupr)
        tdl = -max(1e-12, min(Dp, Dd - t0))
        val = s0 * _domega(upr)
        sgn = -1 if val < 0 else 1
        low = max(low, min(upr, baseupr - safediv_pos(tdl, abs(val)) * sgn))
    low = _clip(min(low, upr), baselow, baseupr)
    upr = _clip(max(low, upr), baselow, baseupr)
    if low != upr:
        fl, fu = _hfun_f(v0, low), _hfun_f(v0, upr)
        if fl * fu > 0.0:
            if abs(fl) < abs(fu):
                upr = low
            else:
                low = upr
    return low, upr


def proj_pd_exp_cone(v0, primal):
    """SCS(proj_pd_exp_cone), exp_cone.c:373-441 (in place on a length-3 array)."""
    TOL = 1e-8
    if not primal:
        v0[:] = -v0
    v = [float(v0[0]), float(v0[1]), float(v0[2])]
    pdist_sq, vp = _primal_heur(v)
    ddist_sq, vd = _polar_heur(v)
    err = max(abs(vp[0] + vd[0] - v[0]), abs(vp[1] + vd[1] - v[1]), abs(vp[2] + vd[2] - v[2]))
    opt = (v[1] <= 0.0 and v[0] <= 0.0)
    opt = opt or (min(pdist_sq, ddist_sq) <= TOL * TOL)
    opt = opt or (err <= TOL and (vp[0] * vd[0] + vp[1] * vd[1] + vp[2] * vd[2]) <= TOL)
    if not opt:
        xl, xh = _exp_search_bracket(v, pdist_sq, ddist_sq)
        rho = _root_search_newton(v, xl, xh, 0.5 * (xl + xh))
        if primal:  # proj_sol_primal_exp_cone, exp_cone.c:326-345
            linrho = (rho - 1.0) * v[0] + v[1]
            exprho = _exp(rho)
            if linrho > 0.0 and _isfinite(exprho):
                quad = rho * (rho - 1.0) + 1.0
                vh = [rho * linrho / quad, linrho / quad, exprho * linrho / quad]
                dh = _nds3(vh, v)
            else:
                vh, dh = [0.0, 0.0, EXP_CONE_INFINITY_VALUE], EXP_CONE_INFINITY_VALUE
            if dh <= pdist_sq:
                vp, pdist_sq = vh, dh
        else:       # proj_sol_polar_exp_cone, exp_cone.c:348-367
            linrho = v[0] - rho * v[1]
            exprho = _exp(-rho)
            if linrho > 0.0 and _isfinite(exprho):
                quad = rho * (rho - 1.0) + 1.0
                vh = [linrho / quad, (1.0 - rho) * linrho / quad, -exprho * linrho / quad]
                dh = _nds3(v, vh)
            else:
                vh, dh = [0.0, 0.0, -EXP_CONE_INFINITY_VALUE], EXP_CONE_INFINITY_VALUE
            if dh <= ddist_sq:
                vd, ddist_sq = vh, dh
    if primal:
        v0[:] = vp
        return math.sqrt(pdist_sq)
    v0[:] = [-vd[0], -vd[1], -vd[2]]
    return math.sqrt(ddist_sq)


class ConeWork:
    """ScsConeWork (cones.h:24-78): cone dict + box warm start + one-time box normalisation."""

    def __init__(self, k, m):
        self.k = cone_dict(k)
        self.m = m
        self.scaled_cones = False
        self.box_t_warm_start = 1.0
        self.bounds = cone_boundaries(self.k)


def proj_cone(x, c: ConeWork, r_y):
    """proj_cone, cones.c:1332-1486: project x onto the PRIMAL cone K, in place."""
    k = c.k
    count = 0
    if k["z"]:
        x[:k["z"]] = 0.0
        count += k["z"]
    if k["l"]:
        np.maximum(x[count:count + k["l"]], 0.0, out=x[count:count + k["l"]])
        count += k["l"]
    if k["bsize"]:
        r_box = r_y[count:count + k["bsize"]] if r_y is not None else None
        c.box_t_warm_start = proj_box_cone(x[count:count + k["bsize"]], k["bl"], k["bu"],
                                           c.box_t_warm_start, r_box)
        count += k["bsize"]
    for q in k["q"]:
        proj_soc(x[count:count + q])
        count += q
    for s in k["s"]:
        sz = s * (s + 1) // 2
        proj_psd(x[count:count + sz], s)
        count += sz
    for cs in k["cs"]:
        proj_cpsd(x[count:count + cs * cs], cs)
        count += cs * cs
    for i in range(k["ep"] + k["ed"]):
        proj_pd_exp_cone(x[count + 3 * i: count + 3 * i + 3], i < k["ep"])
    count += 3 * (k["ep"] + k["ed"])
    for i, p in enumerate(k["p"]):
        idx = count + 3 * i
        if p >= 0:
            proj_power_cone(x[idx:idx + 3], p)
        else:
            v = -x[idx:idx + 3]
            proj_power_cone(v, -p)
            x[idx:idx + 3] += v
    count += 3 * len(k["p"])
    return 0


def proj_dual_cone(x, c: ConeWork, D, r_y):
    """SCS(proj_dual_cone), cones.c:1544-1588: x <- Pi_{K*}^{R}(x) via Moreau, in place.
    D is scal->D (or None when normalize=0); used once to normalise the box bounds."""
    k = c.k
    if not c.scaled_cones:
        if k["bsize"] and len(k["bu"]) and len(k["bl"]):
            c.box_t_warm_start = 1.0
            if D is not None:  # only with a scaling struct (cones.c:1553-1554)
                Db = D[k["z"] + k["l"]:]
                for j in range(k["bsize"] - 1):  # normalize_box_cone, cones.c:1153-1169
                    factor = Db[j + 1] / Db[0]
                    k["bu"][j] = math.inf if k["bu"][j] >= MAX_BOX_VAL else k["bu"][j] * factor
                    k["bl"][j] = -math.inf if k["bl"][j] <= -MAX_BOX_VAL else k["bl"][j] * factor
        c.scaled_cones = True
    s = x.copy()
    if r_y is not None:
        x *= -r_y
    else:
        x *= -1.0
    status = proj_cone(x, c, r_y)
    if r_y is not None:
        x[:] = x / r_y + s
    else:
        x += s
    return status


# ----------------------------------------------------------------------------------------
# indirect linear system  (S/linsys/cpu/indirect/private.c)
# ----------------------------------------------------------------------------------------
class LinSys:
    """ScsLinSysWork of the CPU indirect backend (private.h:16-31)."""

    def __init__(self, A, P, diag_r):
        self.A = A.tocsc()
        self.At = self.A.T.tocsc()   # CSC(A') == CSR(A); transpose(), private.c:7-46
        self.P = P                   # upper-tri csc or None
        self.Pfull = None if P is None else (P + sp.triu(P, 1).T).tocsr()
        self.m, self.n = A.shape
        self.tot_cg_its = 0
        self.update_diag_r(diag_r)

    def update_diag_r(self, diag_r):
        """set_preconditioner, private.c:50-84."""
        n, m = self.n, self.m
        self.diag_r = diag_r
        A = self.A
        M = diag_r[:n].copy()
        cols = np.repeat(np.arange(n), np.diff(A.indptr))
        np.add.at(M, cols, A.data * A.data / diag_r[n + A.indices])
        if self.P is not None:
            M += self.P.diagonal()
        self.M = 1.0 / M

    def mat_vec(self, x):
        """y = (R_x + P + A' R_y^{-1} A) x, private.c:108-121."""
        n = self.n
        y = np.zeros(n)
        if self.Pfull is not None:
            y += self.Pfull @ x
        z = (self.A @ x) / self.diag_r[n:n + self.m]
        y += self.A.T @ z
        y += self.diag_r[:n] * x
        return y

    def pcg(self, s, b, max_its, tol):
        """pcg, private.c:135-219.  b is rhs (length n) and is overwritten with the solution."""
        n = self.n
        if s is None:
            r = b.copy()
            b[:] = 0.0
        else:
            r = b - self.mat_vec(s)
            b[:] = s
        if norm_inf(r) < max(tol, 1e-12):
            return 0
        z = r * self.M
        ztr = float(z @ r)
        p = z.copy()
        i = 0
        while i < max_its:
            Gp = self.mat_vec(p)
            alpha = ztr / float(p @ Gp)
            b += alpha * p
            r -= alpha * Gp
            ztr_prev = ztr
            z = r * self.M
            ztr = float(z @ r)
            norm_r = norm_inf(r)
            if norm_r < tol:
                return i + 1
            if ztr_prev == 0.0:
                break
            beta = ztr / ztr_prev
            p = z + beta * p
            i += 1
        return i

    def solve(self, b, s, tol):
        """scs_solve_lin_sys, private.c:276-316: b=[rx;ry] in, [x;y] out (in place)."""
        n, m = self.n, self.m
        if norm_inf(b) <= 1e-12:
            b[:] = 0.0
            return 0
        tmp = b[n:] / self.diag_r[n:n + m]
        b[:n] += self.A.T @ tmp
        its = self.pcg(s, b[:n], 10 * n, tol)
        b[n:] = (-b[n:] + self.A @ b[:n]) / self.diag_r[n:n + m]
        self.tot_cg_its += its
        return 0


# ----------------------------------------------------------------------------------------
# Anderson acceleration  (S/src/aa.c)
# ----------------------------------------------------------------------------------------
class AaWork:
    def __init__(self, dim, mem, min_len, type1, regularization, relaxation,
                 safeguard_factor=AA_SAFEGUARD_FACTOR, max_weight_norm=AA_MAX_WEIGHT_NORM,
                 ir_max_steps=AA_IR_MAX_STEPS):
        """aa_init, aa.c:657-820."""
        self.type1, self.dim = int(type1), dim
        self.mem = min(mem, dim)
        self.min_len = min(min_len, self.mem) if self.mem > 0 else 0
        self.regularization, self.relaxation = regularization, relaxation
        self.safeguard_factor, self.max_weight_norm = safeguard_factor, max_weight_norm
        self.ir_max_steps = ir_max_steps
        self.iter, self.success, self.norm_g = 0, 0, 0.0
        self.x = np.zeros(dim); self.f = np.zeros(dim); self.g = np.zeros(dim)
        self.g_prev = np.zeros(dim)
        self.Y = np.zeros((dim, max(self.mem, 1)), order="F")
        self.S = np.zeros_like(self.Y); self.D = np.zeros_like(self.Y)
        self.nrm_s_col = np.zeros(max(self.mem, 1)); self.nrm_y_col = np.zeros(max(self.mem, 1))
        self.x_work = np.zeros(dim) if relaxation != 1.0 else None
        self.stats = dict(n_accept=0, n_reject_lapack=0, n_reject_rank0=0, n_reject_nonfinite=0,
                          n_reject_weight_cap=0, n_safeguard_reject=0, last_rank=0,
                          last_aa_norm=math.nan, last_regularization=0.0)

    @staticmethod
    def _frob(nrm_col):  # frob_from_col_norms, aa.c:257-270
        mx = float(np.max(nrm_col))
        if mx == 0:
            return 0.0
        return mx * math.sqrt(float(np.sum((nrm_col / mx) ** 2)))

    def reset(self):  # aa_reset, aa.c:934-964
        self.iter, self.success, self.norm_g = 0, 0, 0.0
        self.nrm_s_col[:] = 0.0
        self.nrm_y_col[:] = 0.0

    def _update(self, x, f):  # update_accel_params, aa.c:340-390
        idx = (self.iter - 1) % self.mem
        self.S[:, idx] = x - self.x
        self.D[:, idx] = f - self.f
        self.g = x - f
        self.Y[:, idx] = self.g - self.g_prev
        self.nrm_s_col[idx] = float(np.linalg.norm(self.S[:, idx]))
        self.nrm_y_col[idx] = float(np.linalg.norm(self.Y[:, idx]))
        self.x = x.copy(); self.f = f.copy(); self.g_prev = self.g.copy()
        if self.x_work is not None:
            self.x_work = x.copy()
        self.norm_g = float(np.linalg.norm(self.g))

    def _solve(self, f, length):  # solve, aa.c:422-652
        dim, mem = self.dim, self.mem
        A_src = self.S if self.type1 else self.Y
        if self.regularization > 0:
            nrm_y = self._frob(self.nrm_y_col)
            nrm_a = self._frob(self.nrm_s_col) if self.type1 else nrm_y
            r = self.regularization * nrm_a * nrm_y
        elif self.regularization < 0:
            r = -self.regularization
        else:
            r = 0.0
        sqrt_r = math.sqrt(r) if r > 0 else 0.0
        aug = dim + mem
        A_aug = np.zeros((aug, length), order="F")
        A_aug[:dim, :] = A_src[:, :length]
        A_aug[dim + np.arange(length), np.arange(length)] = sqrt_r
        info = 0
        rank = 0
        gamma = np.zeros(length)
        try:
            Q, R, piv = sla.qr(A_aug, mode="economic", pivoting=True)  # dgeqp3
        except Exception:
            info = -1
        lapack_info = info
        if info == 0:
            r11 = abs(R[0, 0])
            if r11 > 0:
                tol = r11 * length * np.finfo(float).eps
                while rank < length and abs(R[rank, rank]) >= tol:
                    rank += 1
            if rank == 0:
                info = 1
        if info == 0:
            c_aug = np.concatenate([self.g, np.zeros(mem)])
            c_top = (Q[:, :rank].T @ c_aug)                      # dormqr on c
            if self.type1:
                B_aug = np.zeros((aug, rank), order="F")
                for i in range(rank):
                    B_aug[:dim, i] = self.Y[:, piv[i]]
                    B_aug[dim + piv[i], i] = sqrt_r
                W = Q[:, :rank].T @ B_aug                        # dormqr on B, top block
                try:
                    lu, ipiv = sla.lu_factor(W, check_finite=False)   # dgesv
                    if np.any(np.diag(lu) == 0):
                        raise np.linalg.LinAlgError
                    gred = sla.lu_solve((lu, ipiv), c_top, check_finite=False)
                    prev = 0.0
                    for kk in range(self.ir_max_steps):
                        res = c_top - W @ gred
                        dlt = sla.lu_solve((lu, ipiv), res, check_finite=False)
                        dn = float(np.linalg.norm(dlt))
                        gred = gred + dlt
                        if kk > 0 and dn >= 0.5 * prev:
                            break
                        prev = dn
                except Exception:
                    info = 2
            else:
                Rr = R[:rank, :rank]
                gred = sla.solve_triangular(Rr, c_top, check_finite=False)
                prev = 0.0
                for kk in range(self.ir_max_steps):
                    res = c_top - Rr @ gred
                    dlt = sla.solve_triangular(Rr, res, check_finite=False)
                    dn = float(np.linalg.norm(dlt))
                    gred = gred + dlt
                    if kk > 0 and dn >= 0.5 * prev:
                        break
                    prev = dn
            if info == 0:
                gamma[piv[:rank]] = gred
        aa_norm = float(np.linalg.norm(gamma)) if info == 0 else -1.0
        st = self.stats
        st["last_rank"], st["last_regularization"] = rank, r
        st["last_aa_norm"] = aa_norm if (info == 0 and math.isfinite(aa_norm)) else math.nan
        if info != 0 or not math.isfinite(aa_norm) or aa_norm >= self.max_weight_norm:
            if lapack_info != 0:
                st["n_reject_lapack"] += 1
            elif rank == 0:
                st["n_reject_rank0"] += 1
            elif not math.isfinite(aa_norm):
                st["n_reject_nonfinite"] += 1
            else:
                st["n_reject_weight_cap"] += 1
            self.success = 0
            self.reset()
            if not math.isfinite(aa_norm):
                aa_norm = -1.0
            return aa_norm if aa_norm < 0 else -aa_norm
        f -= self.D[:, :length] @ gamma
        if self.relaxation != 1.0:  # relax, aa.c:393-408
            self.x_work -= self.S[:, :length] @ gamma
            f *= self.relaxation
            f += (1.0 - self.relaxation) * self.x_work
        self.success = 1
        return aa_norm

    def apply(self, f, x):  # aa_apply, aa.c:822-854  (f overwritten in place)
        aa_norm = 0.0
        length = min(self.iter, self.mem)
        self.success = 0
        if self.mem <= 0:
            return 0.0
        if self.iter == 0:
            self.x = x.copy(); self.f = f.copy(); self.g_prev = x - f
            self.iter += 1
            return 0.0
        self._update(x, f)
        if self.iter >= self.min_len:
            aa_norm = self._solve(f, length)
            if aa_norm > 0:
                self.stats["n_accept"] += 1
        self.iter += 1
        return aa_norm

    def safeguard(self, f_new, x_new):  # aa_safeguard, aa.c:856-901
        if self.mem <= 0 or not self.success:
            return 0
        self.success = 0
        norm_diff = float(np.linalg.norm(x_new - f_new))
        if norm_diff > self.safeguard_factor * self.norm_g:
            f_new[:] = self.f
            x_new[:] = self.x
            self.stats["n_safeguard_reject"] += 1
            self.reset()
            return -1
        return 0


# ----------------------------------------------------------------------------------------
# root_plus  (S/src/scs.c:667-688)
# ----------------------------------------------------------------------------------------
def root_plus(g, diag_r, p, mu, eta):
    nm = len(g)
    r = diag_r[:nm]
    gg = float(np.sum(g * g * r)); mug = float(np.sum(mu[:nm] * g * r))
    pg = float(np.sum(p[:nm] * g * r)); pp = float(np.sum(p[:nm] * p[:nm] * r))
    pmu = float(np.sum(p[:nm] * mu[:nm] * r))
    tau_scale = diag_r[nm]
    a = tau_scale + gg
    b = mug - 2 * pg - eta * tau_scale
    c = pp - pmu
    rad = b * b - 4 * a * c
    return (-b + math.sqrt(max(rad, 0.0))) / (2 * a)


# ----------------------------------------------------------------------------------------
# the solver  (S/src/scs.c)
# ----------------------------------------------------------------------------------------
class Residuals:
    def __init__(self):
        self.last_iter = -1
        for f in ("xt_p_x", "xt_p_x_tau", "ctx", "ctx_tau", "bty", "bty_tau", "pobj", "dobj", "gap",
                  "tau", "kap", "res_pri", "res_dual"):
            setattr(self, f, 0.0)
        self.res_infeas = self.res_unbdd_p = self.res_unbdd_a = math.nan
        self.ax = self.ax_s = self.px = self.aty = self.ax_s_btau = self.px_aty_ctau = None


def _compute_residuals(r: Residuals, pd):  # scs.c:441-463
    tol = INFEAS_NEGATIVITY_TOL / pd
    r.res_pri = safediv_pos(norm_inf(r.ax_s_btau), r.tau)
    r.res_dual = safediv_pos(norm_inf(r.px_aty_ctau), r.tau)
    r.res_unbdd_a = r.res_unbdd_p = r.res_infeas = math.nan
    if r.ctx_tau < -tol:
        r.res_unbdd_a = safediv_pos(norm_inf(r.ax_s), -r.ctx_tau)
        r.res_unbdd_p = safediv_pos(norm_inf(r.px), -r.ctx_tau)
    if r.bty_tau < -tol:
        r.res_infeas = safediv_pos(norm_inf(r.aty), -r.bty_tau)


class ScsOracle:
    """scs_init / scs_solve / scs_update restated (scs.c:1193-1430)."""

    def __init__(self, data, cone, **settings):
        self.stgs = dict(DEFAULT_SETTINGS)
        self.stgs.update(settings)
        A = sp.csc_matrix(data["A"]).astype(float)
        A.sort_indices()
        P = data.get("P", None)
        if P is not None:
            P = sp.triu(sp.csc_matrix(P), format="csc").astype(float)
            P.sort_indices()
        self.m, self.n = A.shape
        m, n = self.m, self.n
        self.l = n + m + 1
        self.k = cone_dict(cone)
        if cone_dims(self.k) != m:
            raise ValueError("cone dims != rows of A")
        self.cone_work = ConeWork(self.k, m)
        self.k = self.cone_work.k
        self.diag_r = np.zeros(self.l)
        self._set_diag_r()
        if self.stgs["normalize"]:
            self.A, self.P, self.D, self.E = normalize_a_p(A, P, self.k)
        else:
            self.A, self.P, self.D, self.E = A, P, None, None
        self.primal_scale = self.dual_scale = 1.0
        self.update(np.asarray(data["b"], dtype=float), np.asarray(data["c"], dtype=float))
        self.p = LinSys(self.A, self.P, self.diag_r)
        st = self.stgs
        self.accel = None
        if st["acceleration_lookback"]:
            self.accel = AaWork(self.l, st["acceleration_lookback"], st["acceleration_lookback"],
                                st["acceleration_type_1"], st["acceleration_regularization"],
                                st["acceleration_relaxation"])
        l = self.l
        self.u = np.zeros(l); self.u_t = np.zeros(l); self.v = np.zeros(l)
        self.v_prev = np.zeros(l); self.rsk = np.zeros(l); self.g = np.zeros(l - 1)
        self.r_n = Residuals(); self.r_o = Residuals() if st["normalize"] else self.r_n
        for r in {self.r_n, self.r_o}:
            r.ax_s_btau = np.zeros(m); r.px_aty_ctau = np.zeros(n)
        self.sol = dict(x=np.zeros(n), y=np.zeros(m), s=np.zeros(m))
        self.trace = []

    # -- set_diag_r, scs.c:929-938
    def _set_diag_r(self):
        n, m = self.n, self.m
        self.diag_r[:n] = self.stgs["rho_x"]
        self.diag_r[n:n + m] = set_r_y(self.k, m, self.stgs["scale"])
        self.diag_r[n + m] = TAU_FACTOR

    # -- scs_update, scs.c:1235-1273
    def update(self, b=None, c=None):
        if b is not None:
            self.b_orig = np.array(b, dtype=float)
            self.nm_b_orig = norm_inf(self.b_orig)
        if c is not None:
            self.c_orig = np.array(c, dtype=float)
            self.nm_c_orig = norm_inf(self.c_orig)
        self.b, self.c = self.b_orig.copy(), self.c_orig.copy()
        if self.D is not None:
            self.b, self.c, sig = normalize_b_c(self.D, self.E, self.b, self.c)
            self.primal_scale = self.dual_scale = sig

    def _un_normalize_sol(self, x, y, s):  # normalize.c:78-90
        return (x * (self.E / self.dual_scale), y * (self.D / self.primal_scale),
                s / (self.D * self.dual_scale))

    def _normalize_sol(self, x, y, s):  # normalize.c:64-76
        return (x / (self.E / self.dual_scale), y / (self.D / self.primal_scale),
                s * (self.D * self.dual_scale))

    # -- populate_residual_struct, scs.c:513-585
    def _populate_residuals(self, it):
        n, m = self.n, self.m
        r = self.r_n
        if r.last_iter == it:
            return
        r.last_iter = it
        x, y, s = self.u[:n].copy(), self.u[n:n + m].copy(), self.rsk[n:n + m].copy()
        self.xys_n = (x, y, s)
        r.tau, r.kap = abs(self.u[n + m]), abs(self.rsk[n + m])
        r.ax = self.A @ x
        r.ax_s = r.ax + s
        r.ax_s_btau = r.ax_s - r.tau * self.b
        if self.P is not None:
            r.px = self.p.Pfull @ x
            r.xt_p_x_tau = float(r.px @ x)
        else:
            r.px = np.zeros(n)
            r.xt_p_x_tau = 0.0
        r.aty = self.A.T @ y
        r.px_aty_ctau = r.px + r.aty + r.tau * self.c
        r.bty_tau = float(y @ self.b)
        r.ctx_tau = float(x @ self.c)
        r.bty = safediv_pos(r.bty_tau, r.tau)
        r.ctx = safediv_pos(r.ctx_tau, r.tau)
        r.xt_p_x = safediv_pos(r.xt_p_x_tau, r.tau * r.tau)
        r.gap = abs(r.xt_p_x + r.ctx + r.bty)
        r.pobj = r.xt_p_x / 2.0 + r.ctx
        r.dobj = -r.xt_p_x / 2.0 - r.bty
        _compute_residuals(r, 1.0)
        if self.stgs["normalize"]:
            self.xys_o = self._un_normalize_sol(x, y, s)
            ro = self.r_o
            pd = self.primal_scale * self.dual_scale   # unnormalize_residuals, scs.c:465-509
            ro.last_iter, ro.tau = r.last_iter, r.tau
            for f in ("kap", "bty_tau", "ctx_tau", "xt_p_x_tau", "xt_p_x", "ctx", "bty", "pobj", "dobj", "gap"):
                setattr(ro, f, getattr(r, f) / pd)
            fD = (1.0 / self.dual_scale) / self.D
            fE = (1.0 / self.primal_scale) / self.E
            ro.ax, ro.ax_s, ro.ax_s_btau = r.ax * fD, r.ax_s * fD, r.ax_s_btau * fD
            ro.aty, ro.px, ro.px_aty_ctau = r.aty * fE, r.px * fE, r.px_aty_ctau * fE
            _compute_residuals(ro, pd)
        else:
            self.xys_o = self.xys_n

    # -- has_converged, scs.c:589-627
    def _has_converged(self):
        r = self.r_o
        st = self.stgs
        lt = lambda a, b: (not math.isnan(a)) and (not math.isnan(b)) and a < b   # isless()
        if r.tau > 0.0:
            grl = max(abs(r.xt_p_x), abs(r.ctx), abs(r.bty))
            nm_s = norm_inf(self.xys_o[2])
            prl = max(self.nm_b_orig * r.tau, nm_s, norm_inf(r.ax)) / r.tau
            drl = max(self.nm_c_orig * r.tau, norm_inf(r.px), norm_inf(r.aty)) / r.tau
            if (lt(r.res_pri, st["eps_abs"] + st["eps_rel"] * prl)
                    and lt(r.res_dual, st["eps_abs"] + st["eps_rel"] * drl)
                    and lt(r.gap, st["eps_abs"] + st["eps_rel"] * grl)):
                return SCS_SOLVED
        if lt(r.res_unbdd_a, st["eps_infeas"]) and lt(r.res_unbdd_p, st["eps_infeas"]):
            return SCS_UNBOUNDED
        if lt(r.res_infeas, st["eps_infeas"]):
            return SCS_INFEASIBLE
        return 0

    # -- update_work_cache, scs.c:1066-1076
    def _update_work_cache(self):
        self.g[:self.n] = self.c
        self.g[self.n:] = -self.b
        self.p.solve(self.g, None, CG_BEST_TOL)

    # -- update_scale, scs.c:1112-1189
    def _update_scale(self, it):
        r = self.r_o
        st = self.stgs
        nm_ax, nm_s = norm_inf(r.ax), norm_inf(self.xys_o[2])
        denom_pri = max(nm_ax, nm_s, self.nm_b_orig * r.tau)
        rel_pri = safediv_pos(norm_inf(r.ax_s_btau), denom_pri)
        denom_dual = max(norm_inf(r.px), norm_inf(r.aty), self.nm_c_orig * r.tau)
        rel_dual = safediv_pos(norm_inf(r.px_aty_ctau), denom_dual)
        rel_pri, rel_dual = max(rel_pri, DIV_EPS_TOL), max(rel_dual, DIV_EPS_TOL)
        self.sum_log_scale_factor += math.log(rel_pri) - math.log(rel_dual)
        self.n_log_scale_factor += 1
        factor = math.sqrt(math.exp(self.sum_log_scale_factor / self.n_log_scale_factor))
        if it - self.last_scale_update_iter < RESCALING_MIN_ITERS:
            return
        new_scale = min(max(st["scale"] * factor, MIN_SCALE_VALUE), MAX_SCALE_VALUE)
        if new_scale == st["scale"]:
            return
        if factor > math.sqrt(10.0) or factor < 1.0 / math.sqrt(10.0):
            self.scale_updates += 1
            self.sum_log_scale_factor, self.n_log_scale_factor = 0.0, 0
            self.last_scale_update_iter = it
            st["scale"] = new_scale
            self._set_diag_r()
            self.p.update_diag_r(self.diag_r)
            self._update_work_cache()
            if self.accel:
                self.accel.reset()
            self.v = self.rsk / self.diag_r + 2 * self.u_t - self.u

    # -- project_lin_sys, scs.c:691-729
    def _project_lin_sys(self, it):
        n, m, l = self.n, self.m, self.l
        self.u_t[:n] = self.v[:n] * self.diag_r[:n]
        self.u_t[n:l - 1] = -self.v[n:l - 1] * self.diag_r[n:l - 1]
        self.u_t[l - 1] = self.v[l - 1]
        ws = self.u[:n] + self.u[l - 1] * self.g[:n]
        tol = min(norm_inf(self.r_n.ax_s_btau), norm_inf(self.r_n.px_aty_ctau))
        nm_ws = norm_inf(ws) / (it + 1.0) ** CG_RATE
        tol = max(CG_BEST_TOL, CG_TOL_FACTOR * min(tol, nm_ws))
        self.last_cg_tol = tol
        self.p.solve(self.u_t[:l - 1], ws, tol)
        if it < FEASIBLE_ITERS:
            self.u_t[l - 1] = 1.0
        else:
            self.u_t[l - 1] = root_plus(self.g, self.diag_r, self.u_t, self.v, self.v[l - 1])
        self.u_t[:l - 1] -= self.u_t[l - 1] * self.g

    # -- scs_solve, scs.c:1275-1430
    def solve(self, warm_start=False, x=None, y=None, s=None):
        st = self.stgs
        n, m, l = self.n, self.m, self.l
        # reset_tracking, scs.c:1079-1092
        self.last_scale_update_iter, self.sum_log_scale_factor = 0, 0.0
        self.n_log_scale_factor, self.scale_updates = 0, 0
        self.rejected_accel_steps = self.accepted_accel_steps = 0
        self.aa_norm = 0.0
        self.r_n.last_iter = self.r_o.last_iter = -1
        if warm_start:  # warm_start_vars, scs.c:638-657
            sx = np.array(self.sol["x"] if x is None else x, dtype=float)
            sy = np.array(self.sol["y"] if y is None else y, dtype=float)
            ss = np.array(self.sol["s"] if s is None else s, dtype=float)
            if st["normalize"]:
                sx, sy, ss = self._normalize_sol(sx, sy, ss)
            self.v[:n] = np.where(np.isnan(sx), 0.0, sx)
            vy = sy + ss / self.diag_r[n:n + m]
            self.v[n:n + m] = np.where(np.isnan(vy), 0.0, vy)
            self.v[l - 1] = 1.0
        else:
            self.v[:] = 0.0
            self.v[l - 1] = 1.0
        self._update_work_cache()
        status = SCS_UNFINISHED
        it = 0
        self.trace = []
        for it in range(st["max_iters"]):
            i = it
            if self.accel is not None and i > 0 and i % st["acceleration_interval"] == 0:
                self.aa_norm = self.accel.apply(self.v, self.v_prev)
            if i >= FEASIBLE_ITERS:  # normalize_v, scs.c:771-779
                vn = float(np.linalg.norm(self.v))
                if vn != 0.0:
                    self.v *= math.sqrt(l) * ITERATE_NORM / vn
            if self.accel is not None:
                self.v_prev[:] = self.v
            self._project_lin_sys(i)
            # project_cones, scs.c:754-768
            self.u[:] = 2 * self.u_t - self.v
            proj_dual_cone(self.u[n:n + m], self.cone_work, self.D, self.diag_r[n:n + m])
            self.u[l - 1] = 1.0 if i < FEASIBLE_ITERS else max(self.u[l - 1], 0.0)
            # compute_rsk, scs.c:739-744
            self.rsk[:] = (self.v + self.u - 2 * self.u_t) * self.diag_r
            if i % CONVERGED_INTERVAL == 0:
                self._populate_residuals(i)
                status = self._has_converged()
                if status != 0:
                    break
            if st["adaptive_scale"] and i == self.r_o.last_iter:
                self._update_scale(i)
            self.v += st["alpha"] * (self.u - self.u_t)       # update_dual_vars, scs.c:746-751
            if self.accel is not None and i % st["acceleration_interval"] == 0 and self.aa_norm > 0:
                if self.accel.safeguard(self.v, self.v_prev) < 0:
                    self.rejected_accel_steps += 1
                else:
                    self.accepted_accel_steps += 1
        else:
            it = st["max_iters"]
        return self._finalize(status, it)

    # -- finalize + set_* , scs.c:805-924
    def _finalize(self, status, it):
        n, m = self.n, self.m
        x, y, s = self.u[:n].copy(), self.u[n:n + m].copy(), self.rsk[n:n + m].copy()
        if self.stgs["normalize"]:
            x, y, s = self._un_normalize_sol(x, y, s)
        self._populate_residuals(it)
        r = self.r_o
        info = dict(iter=it, res_infeas=r.res_infeas, res_unbdd_a=r.res_unbdd_a,
                    res_unbdd_p=r.res_unbdd_p, scale=self.stgs["scale"],
                    scale_updates=self.scale_updates, comp_slack=abs(float(s @ y)),
                    rejected_accel_steps=self.rejected_accel_steps,
                    accepted_accel_steps=self.accepted_accel_steps,
                    tot_cg_its=self.p.tot_cg_its)

        def set_solved():
            nonlocal x, y, s
            f = safediv_pos(1.0, r.tau)
            x, y, s = x * f, y * f, s * f
            info.update(gap=r.gap, res_pri=r.res_pri, res_dual=r.res_dual,
                        pobj=r.xt_p_x / 2.0 + r.ctx, dobj=-r.xt_p_x / 2.0 - r.bty,
                        status="solved", status_val=SCS_SOLVED)

        def set_infeasible():
            nonlocal x, y, s
            y = y * (-1.0 / r.bty_tau)
            x = np.full(n, math.nan); s = np.full(m, math.nan)
            info.update(gap=math.nan, res_pri=math.nan, res_dual=math.nan, pobj=math.inf,
                        dobj=math.inf, status="infeasible", status_val=SCS_INFEASIBLE)

        def set_unbounded():
            nonlocal x, y, s
            x = x * (-1.0 / r.ctx_tau); s = s * (-1.0 / r.ctx_tau)
            y = np.full(m, math.nan)
            info.update(gap=math.nan, res_pri=math.nan, res_dual=math.nan, pobj=-math.inf,
                        dobj=-math.inf, status="unbounded", status_val=SCS_UNBOUNDED)

        if status == SCS_SOLVED:
            set_solved()
        elif status == SCS_INFEASIBLE:
            set_infeasible()
        elif status == SCS_UNBOUNDED:
            set_unbounded()
        else:  # set_unfinished, scs.c:845-871
            if r.kap > r.tau and (r.bty_tau < 0 or r.ctx_tau < 0):
                if r.bty_tau < 0 and r.bty_tau < r.ctx_tau:
                    set_infeasible(); info["status_val"] = SCS_INFEASIBLE_INACCURATE
                else:
                    set_unbounded(); info["status_val"] = SCS_UNBOUNDED_INACCURATE
            elif r.tau > 0:
                set_solved(); info["status_val"] = SCS_SOLVED_INACCURATE
            else:
                info.update(status="failed", status_val=SCS_FAILED)
            info["status"] = info.get("status", "") + " (inaccurate - reached max_iters)"
        self.sol = dict(x=x, y=y, s=s)
        return dict(x=x, y=y, s=s, info=info)


def solve(data, cone, **settings):
    return ScsOracle(data, cone, **settings).solve()
