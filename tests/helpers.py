"""Shared test helpers: golden fixture loading and the reference's universal post-solve checker."""
from __future__ import annotations

import json
import math
import os

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))


def _unjf(x):
    if isinstance(x, dict):
        return {k: _unjf(v) for k, v in x.items()}
    if isinstance(x, list):
        return [_unjf(v) for v in x]
    if x == "nan":
        return math.nan
    if x == "inf":
        return math.inf
    if x == "-inf":
        return -math.inf
    return x


def golden(name):
    return _unjf(json.load(open(os.path.join(HERE, "golden", name))))


def problem_from_record(rec):
    """kat.json / ref_runs.json record -> (data dict, cone dict)."""
    m, n = rec["m"], rec["n"]
    A = sp.csc_matrix((np.array(rec["Ax"], dtype=float), np.array(rec["Ai"], dtype=np.int32),
                       np.array(rec["Ap"], dtype=np.int32)), shape=(m, n))
    data = dict(A=A, b=np.array(rec["b"], dtype=float), c=np.array(rec["c"], dtype=float))
    if rec.get("Px") is not None:
        data["P"] = sp.csc_matrix((np.array(rec["Px"], dtype=float), np.array(rec["Pi"], dtype=np.int32),
                                   np.array(rec["Pp"], dtype=np.int32)), shape=(n, n))
    cone = {k: v for k, v in rec["cone"].items() if v not in (None, [], 0)}
    return data, cone


def full_P(data):
    P = data.get("P")
    if P is None:
        return None
    P = sp.csc_matrix(P)
    return sp.triu(P) + sp.triu(P, 1).T


def verify_solution(data, cone, sol, eps_abs, eps_rel, cone_tol=1e-5):
    """Restatement of verify_solution_correct (S/test/problems/problem_utils.h:107-249) for a
    'solved' result: recompute residuals / objectives from (x, y, s) with scipy, compare with the
    returned info, check the eps criteria, complementary slackness and cone membership."""
    from oracle import scs_oracle as O
    A = sp.csc_matrix(data["A"])
    b, c = np.asarray(data["b"], float), np.asarray(data["c"], float)
    x, y, s, info = sol["x"], sol["y"], sol["s"], sol["info"]
    P = full_P(data)
    px = P @ x if P is not None else np.zeros_like(x)
    ax, aty = A @ x, A.T @ y
    res_pri = np.max(np.abs(ax + s - b)) if len(b) else 0.0
    res_dual = np.max(np.abs(px + aty + c)) if len(c) else 0.0
    xpx = float(x @ px)
    pobj = 0.5 * xpx + float(c @ x)
    dobj = -0.5 * xpx - float(b @ y)
    gap = abs(xpx + float(c @ x) + float(b @ y))
    scale = max(1.0, abs(pobj))
    assert abs(info["pobj"] - pobj) <= 1e-7 * scale, ("pobj", info["pobj"], pobj)
    assert abs(info["dobj"] - dobj) <= 1e-7 * max(1.0, abs(dobj)), ("dobj", info["dobj"], dobj)
    assert abs(info["res_pri"] - res_pri) <= 1e-8 * max(1.0, res_pri) + 1e-9, ("res_pri", info["res_pri"], res_pri)
    assert abs(info["res_dual"] - res_dual) <= 1e-8 * max(1.0, res_dual) + 1e-9, ("res_dual", info["res_dual"], res_dual)
    assert abs(info["gap"] - gap) <= 1e-7 * max(1.0, gap) + 1e-9, ("gap", info["gap"], gap)
    ninf = lambda v: float(np.max(np.abs(v))) if len(v) else 0.0
    prl = max(ninf(b), ninf(s), ninf(ax))
    drl = max(ninf(c), ninf(px), ninf(aty))
    grl = max(abs(xpx), abs(float(c @ x)), abs(float(b @ y)))
    assert res_pri <= (eps_abs + eps_rel * prl) * (1 + 1e-6) + 1e-12
    assert res_dual <= (eps_abs + eps_rel * drl) * (1 + 1e-6) + 1e-12
    assert gap <= (eps_abs + eps_rel * grl) * (1 + 1e-6) + 1e-12
    # cone membership: s in K  <=>  proj_{K*}(-s) = 0 ;  y in K*  <=>  proj_{K*}(y) = y
    m = len(b)
    cw = O.ConeWork(cone, m)
    t = -s.copy(); O.proj_dual_cone(t, cw, None, None)
    assert ninf(t) <= cone_tol * max(1.0, ninf(s)), ("s not in K", ninf(t))
    t = y.copy(); O.proj_dual_cone(t, O.ConeWork(cone, m), None, None)
    assert ninf(t - y) <= cone_tol * max(1.0, ninf(y)), ("y not in K*", ninf(t - y))
    assert abs(float(s @ y)) <= 1e-4 * max(1.0, ninf(s), ninf(y)) * max(1.0, math.sqrt(m)), ("comp slack", float(s @ y))
