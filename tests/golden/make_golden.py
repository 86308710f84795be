#!/usr/bin/env python
"""Generate tests/golden/*.json from the reference tree.  Runs ONLY in the build container
(needs /root/reference and the compiled reference in oracle/_ref); the fixtures it writes are
committed so that the GPU box and the CPU test tier never read /root/reference.

    make -C oracle ref && python tests/golden/make_golden.py

Outputs
  kat.json        known-answer data transcribed from the reference's own C tests
                  (S/test/problems/*.h): exponential-cone projection vectors, root_plus cases,
                  tiny problems with their published optimal objectives, data-file problems
                  (random_prob, mpc_bug1-3) converted from the reference binary format.
  ref_runs.json   outputs of the compiled reference itself on seeded inputs: full solves
                  (QDLDL and CPU_INDIRECT), SCS(proj_dual_cone), scs_solve_lin_sys,
                  aa_apply/aa_safeguard sequences, SCS(accum_by_*).
"""
from __future__ import annotations

import ctypes as C
import json
import math
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
REF = "/root/reference"
PROB = os.path.join(REF, "scs_source", "test", "problems")

import numpy as np  # noqa: E402
import scipy.sparse as sp  # noqa: E402


def jf(x):
    """json-safe floats"""
    if isinstance(x, dict):
        return {k: jf(v) for k, v in x.items()}
    if isinstance(x, (list, tuple, np.ndarray)):
        return [jf(v) for v in x]
    if isinstance(x, (float, np.floating)):
        x = float(x)
        if math.isnan(x):
            return "nan"
        if math.isinf(x):
            return "inf" if x > 0 else "-inf"
        return x
    if isinstance(x, (int, np.integer)):
        return int(x)
    return x


# ------------------------------------------------------------------ C header mini-parser --
def c_functions(text):
    """{name: body} for every `static const char *name(void) { ... }`."""
    out = {}
    for m in re.finditer(r"static const char \*(\w+)\(void\)\s*\{", text):
        i = m.end()
        depth = 1
        while depth and i < len(text):
            depth += {"{": 1, "}": -1}.get(text[i], 0)
            i += 1
        out[m.group(1)] = text[m.end():i - 1]
    return out


def strip_comments(s):
    s = re.sub(r"/\*.*?\*/", " ", s, flags=re.S)
    return re.sub(r"//[^\n]*", " ", s)


def c_eval(expr, ns):
    expr = expr.strip()
    expr = re.sub(r"\bINFINITY\b", "float('inf')", expr)
    expr = re.sub(r"\bSQRTF\b|\bsqrt\b", "math.sqrt", expr)
    expr = re.sub(r"(\d)\.(?=[^\d]|$)", r"\1.0", expr)
    return eval(expr, {"math": math, "float": float}, ns)


def parse_problem(body):
    """Extract arrays, scalars, cone and settings assignments made before the first scs( call."""
    body = strip_comments(body)
    cut = re.search(r"\bscs\(\s*d\s*,", body)
    pre = body[:cut.start()] if cut else body
    ns = {}
    for m in re.finditer(r"scs_(float|int)\s+(\w+)\s*\[\s*\d*\s*\]\s*=\s*\{(.*?)\}\s*;", pre, flags=re.S):
        vals = [v for v in m.group(3).replace("\n", " ").split(",") if v.strip()]
        ns[m.group(2)] = [c_eval(v, ns) for v in vals]
    for m in re.finditer(r"scs_(float|int)\s+((?:\w+\s*=\s*[^,;\[\]]+,?\s*)+);", pre):
        for part in m.group(2).split(","):
            if "=" in part:
                name, val = part.split("=", 1)
                try:
                    ns[name.strip()] = c_eval(val, ns)
                except Exception:
                    pass
    cone, stgs = {}, {}
    for m in re.finditer(r"k->(\w+)\s*=\s*([^;]+);", pre):
        try:
            cone[m.group(1)] = c_eval(m.group(2), ns)
        except Exception:
            pass
    for m in re.finditer(r"stgs->(\w+)\s*=\s*([^;]+);", pre):
        try:
            stgs[m.group(1)] = c_eval(m.group(2), ns)
        except Exception:
            pass
    return ns, cone, stgs


def problem_from_header(fname, func, names=None):
    text = open(os.path.join(PROB, fname)).read()
    ns, cone, stgs = parse_problem(c_functions(text)[func])
    nm = dict(Ax="Ax", Ai="Ai", Ap="Ap", Px="Px", Pi="Pi", Pp="Pp", b="b", c="c", m="m", n="n", opt="opt")
    nm.update(names or {})
    g = lambda k: ns.get(nm[k])
    m, n = int(g("m")), int(g("n"))
    out = dict(name=func, source="S/test/problems/%s" % fname, m=m, n=n, Ax=g("Ax"), Ai=g("Ai"), Ap=g("Ap"),
               b=g("b"), c=g("c"), opt=g("opt"))
    if g("Px") is not None:
        out.update(Px=g("Px"), Pi=g("Pi"), Pp=g("Pp"))
    k = {}
    for f in ("z", "l", "ep", "ed"):
        if f in cone:
            k[f] = int(cone[f])
    for f in ("q", "s", "cs", "p", "bu", "bl"):
        if f in cone and isinstance(cone[f], list):
            k[f] = cone[f]
    out["cone"] = k
    out["settings"] = {kk: vv for kk, vv in stgs.items() if isinstance(vv, (int, float))}
    assert len(out["Ap"]) == n + 1 and len(out["b"]) == m and len(out["c"]) == n, (func, "shape mismatch")
    return out


def build_kat():
    kat = {}
    # --- exponential cone vectors, test_exp_cone.h:52-77
    t = strip_comments(open(os.path.join(PROB, "test_exp_cone.h")).read())
    arrs = {}
    for m in re.finditer(r"scs_float (\w+)\[6\]\[3\]\s*=\s*\{(.*?)\};", t, flags=re.S):
        rows = re.findall(r"\{([^{}]*)\}", m.group(2))
        arrs[m.group(1)] = [[float(v) for v in r.split(",") if v.strip()] for r in rows]
    kat["exp_cone"] = dict(source="S/test/problems/test_exp_cone.h:52-77", tol=1e-6, v0=arrs["v0"], vp=arrs["vp"],
                           vd=arrs["vd"])
    # --- root_plus cases, test_root_plus.h:68-159 (expected value by exact rational arithmetic)
    t = strip_comments(open(os.path.join(PROB, "test_root_plus.h")).read())
    cases = []
    for blk in re.findall(r"\{\s*(scs_float g\d.*?)old_val\s*=", t, flags=re.S):
        a = {m.group(1)[0:2].rstrip("0123456789"): [float(v) for v in m.group(2).split(",")]
             for m in re.finditer(r"scs_float (\w+)\[\]\s*=\s*\{(.*?)\};", blk, flags=re.S)}
        ts = float(re.search(r"tau_scale\s*=\s*([^;]+);", blk).group(1))
        eta = float(re.search(r"eta\s*=\s*([^;]+);", blk).group(1))
        from fractions import Fraction as F
        g, p, mu, r = (list(map(F, a[k])) for k in ("g", "p", "mu", "r"))
        dot = lambda x, y: sum(xi * yi * ri for xi, yi, ri in zip(x, y, r))
        aa = F(ts) + dot(g, g); bb = dot(mu, g) - 2 * dot(p, g) - F(eta) * F(ts); cc = dot(p, p) - dot(p, mu)
        rad = max(bb * bb - 4 * aa * cc, 0)
        val = (-float(bb) + math.sqrt(float(rad))) / (2 * float(aa))
        cases.append(dict(g=a["g"], p=a["p"], mu=a["mu"], r=a["r"], tau_scale=ts, eta=eta, expected=val))
    assert len(cases) == 5
    kat["root_plus"] = dict(source="S/test/problems/test_root_plus.h:68-159", cases=cases)
    # --- tiny problems with published objectives
    plist = [("test_zero_cone.h", "test_zero_cone", None), ("hs21_tiny_qp.h", "hs21_tiny_qp", None),
             ("test_box_cone.h", "test_box_cone", None), ("test_psd_n1.h", "test_psd_n1", None),
             ("test_dual_exp_cone.h", "test_dual_exp_cone", None), ("test_mixed_cones.h", "test_mixed_cones", None),
             ("qafiro_tiny_qp.h", "qafiro_tiny_qp", None), ("complex_PSD.h", "complex_PSD", None)]
    t = open(os.path.join(PROB, "test_soc_sizes.h")).read()
    for fn in c_functions(t):
        plist.append(("test_soc_sizes.h", fn, None))
    t = open(os.path.join(PROB, "test_power_cone.h")).read()
    for fn in c_functions(t):
        plist.append(("test_power_cone.h", fn, None))
    probs = []
    for fname, func, names in plist:
        try:
            pr = problem_from_header(fname, func, names)
            if pr["opt"] is None:
                raise ValueError("no opt")
            probs.append(pr)
        except Exception as e:
            print("skip %s:%s (%r)" % (fname, func, e))
    kat["problems"] = probs
    # --- data-file problems through the reference's own reader
    kat["file_problems"] = read_file_problems()
    # --- struct layout of the reference's public headers (non-DLONG build), from its own compiler view
    kat["abi"] = reference_abi()
    return kat


def reference_abi():
    import subprocess, tempfile
    src = r"""
#include <stdio.h>
#include <stddef.h>
#include "scs.h"
#define SZ(T) printf("\"sizeof_%s\": %zu,\n", #T, sizeof(T))
#define OFF(T, f) printf("\"offsetof_%s_%s\": %zu,\n", #T, #f, offsetof(T, f))
int main(void) {
  printf("{\n");
  SZ(ScsMatrix); SZ(ScsSettings); SZ(ScsData); SZ(ScsCone); SZ(ScsSolution); SZ(ScsInfo); SZ(AaStats);
  OFF(ScsSettings, scale); OFF(ScsSettings, max_iters); OFF(ScsSettings, alpha); OFF(ScsSettings, verbose);
  OFF(ScsSettings, acceleration_regularization); OFF(ScsSettings, write_data_filename);
  OFF(ScsCone, bu); OFF(ScsCone, bsize); OFF(ScsCone, q); OFF(ScsCone, s); OFF(ScsCone, cs); OFF(ScsCone, ep); OFF(ScsCone, p);
  OFF(ScsCone, psize);
  OFF(ScsInfo, status); OFF(ScsInfo, lin_sys_solver); OFF(ScsInfo, status_val); OFF(ScsInfo, pobj); OFF(ScsInfo, setup_time);
  OFF(ScsInfo, comp_slack); OFF(ScsInfo, rejected_accel_steps); OFF(ScsInfo, aa_stats); OFF(ScsInfo, lin_sys_time);
  OFF(ScsInfo, accel_time); OFF(AaStats, last_aa_norm);
  printf("\"sizeof_scs_int\": %zu, \"sizeof_scs_float\": %zu\n}\n", sizeof(scs_int), sizeof(scs_float));
  return 0;
}
"""
    with tempfile.TemporaryDirectory() as td:
        cf = os.path.join(td, "abi.c")
        open(cf, "w").write(src)
        exe = os.path.join(td, "abi")
        subprocess.check_call(["gcc", "-I", os.path.join(REF, "scs_source", "include"), cf, "-o", exe])
        return json.loads(subprocess.check_output([exe]).decode())


# ------------------------------------------------------------------ reference C library ---
class Mat(C.Structure):
    _fields_ = [("x", C.POINTER(C.c_double)), ("i", C.POINTER(C.c_int)), ("p", C.POINTER(C.c_int)), ("m", C.c_int),
                ("n", C.c_int)]


class Data(C.Structure):
    _fields_ = [("m", C.c_int), ("n", C.c_int), ("A", C.POINTER(Mat)), ("P", C.POINTER(Mat)),
                ("b", C.POINTER(C.c_double)), ("c", C.POINTER(C.c_double))]


class Cone(C.Structure):
    _fields_ = [("z", C.c_int), ("l", C.c_int), ("bu", C.POINTER(C.c_double)), ("bl", C.POINTER(C.c_double)),
                ("bsize", C.c_int), ("q", C.POINTER(C.c_int)), ("qsize", C.c_int), ("s", C.POINTER(C.c_int)),
                ("ssize", C.c_int), ("cs", C.POINTER(C.c_int)), ("cssize", C.c_int), ("ep", C.c_int), ("ed", C.c_int),
                ("p", C.POINTER(C.c_double)), ("psize", C.c_int)]


def reflib():
    return C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libscsindir.so"))


def read_file_problems():
    lib = reflib()
    opts = dict(random_prob=5.751458006385587, mpc_bug1=-0.473957794500, mpc_bug2=-0.029336830816,
                mpc_bug3=-0.002215217478)
    out = []
    for name, opt in opts.items():
        d, k, st = C.POINTER(Data)(), C.POINTER(Cone)(), C.c_void_p()
        rc = lib._scs_read_data(os.path.join(PROB, name).encode(), C.byref(d), C.byref(k), C.byref(st))
        assert rc == 0
        D, K = d.contents, k.contents
        A = D.A.contents
        nnz = A.p[A.n]
        pr = dict(name=name, source="S/test/problems/%s (+ %s.h)" % (name, name.rstrip("123")), m=D.m, n=D.n, opt=opt,
                  Ax=[A.x[i] for i in range(nnz)], Ai=[A.i[i] for i in range(nnz)], Ap=[A.p[i] for i in range(A.n + 1)],
                  b=[D.b[i] for i in range(D.m)], c=[D.c[i] for i in range(D.n)], settings=dict(eps_abs=1e-6, eps_rel=1e-6))
        if D.P:
            P = D.P.contents
            pn = P.p[P.n]
            pr.update(Px=[P.x[i] for i in range(pn)], Pi=[P.i[i] for i in range(pn)], Pp=[P.p[i] for i in range(P.n + 1)])
        kk = dict(z=K.z, l=K.l, ep=K.ep, ed=K.ed)
        if K.bsize > 1:
            kk["bu"] = [K.bu[i] for i in range(K.bsize - 1)]
            kk["bl"] = [K.bl[i] for i in range(K.bsize - 1)]
        for f, sz in (("q", "qsize"), ("s", "ssize"), ("cs", "cssize"), ("p", "psize")):
            nn = getattr(K, sz)
            if nn > 0:
                kk[f] = [getattr(K, f)[i] for i in range(nn)]
        pr["cone"] = kk
        out.append(jf(pr))
    return out


def cmat(A):
    A = sp.csc_matrix(A); A.sort_indices()
    x = np.ascontiguousarray(A.data, dtype=np.float64); i = A.indices.astype(np.int32); p = A.indptr.astype(np.int32)
    return Mat(x.ctypes.data_as(C.POINTER(C.c_double)), i.ctypes.data_as(C.POINTER(C.c_int)),
               p.ctypes.data_as(C.POINTER(C.c_int)), A.shape[0], A.shape[1]), (x, i, p)


def ccone(K):
    keep = []
    k = Cone()
    k.z, k.l, k.ep, k.ed = K.get("z", 0), K.get("l", 0), K.get("ep", 0), K.get("ed", 0)
    def fa(v):
        a = np.asarray(v, dtype=np.float64); keep.append(a); return a.ctypes.data_as(C.POINTER(C.c_double))
    def ia(v):
        a = np.asarray(v, dtype=np.int32); keep.append(a); return a.ctypes.data_as(C.POINTER(C.c_int))
    if K.get("bu"):
        k.bu, k.bl, k.bsize = fa(K["bu"]), fa(K["bl"]), len(K["bu"]) + 1
    for f, sz in (("q", "qsize"), ("s", "ssize"), ("cs", "cssize")):
        if K.get(f):
            setattr(k, f, ia(K[f])); setattr(k, sz, len(K[f]))
    if K.get("p"):
        k.p, k.psize = fa(K["p"]), len(K["p"])
    return k, keep


CONE_CASES = [dict(z=3, l=5), dict(q=[1, 2, 3, 5, 40]), dict(ep=6, ed=6), dict(p=[0.3, -0.6, 0.5, 0.9, -0.1]),
              dict(s=[1, 2, 3, 6]), dict(s=[12]), dict(cs=[1, 2, 3, 5]),
              dict(bu=[1.5, 0.3, 2.0, 1e20, 0.7], bl=[-1.0, -0.2, 0.5, -1e20, -3.0]),
              dict(z=2, l=3, bu=[1.0, 2.0], bl=[-1.0, 0.5], q=[3, 4], s=[3], cs=[2], ep=2, ed=2, p=[0.4, -0.7])]


def build_ref_runs():
    import scs
    from tests import problems
    lib = reflib()
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    out = {}
    # ---- SCS(proj_dual_cone) on seeded inputs
    lib._scs_init_cone.restype = C.c_void_p
    lib._scs_init_cone.argtypes = [C.POINTER(Cone), C.c_int]
    lib._scs_proj_dual_cone.argtypes = [C.POINTER(C.c_double), C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
    lib._scs_finish_cone.argtypes = [C.c_void_p]
    cones = []
    rng = np.random.RandomState(11)
    for K in CONE_CASES:
        m = problems.cone_len(K)
        k, keep = ccone(K)
        w = lib._scs_init_cone(C.byref(k), m)
        for trial in range(3):
            x = rng.randn(m) * (10.0 ** rng.randint(-1, 2))
            r_y = np.full(m, 10.0); r_y[:K.get("z", 0)] = 0.01
            if trial == 2 and "bu" in K:
                r_y = np.abs(rng.randn(m)) + 0.5
            y = x.copy()
            lib._scs_proj_dual_cone(dp(y), w, None, dp(r_y))
            cones.append(dict(cone=K, x=jf(x), r_y=jf(r_y), out=jf(y)))
        lib._scs_finish_cone(w)
    out["proj_dual_cone"] = dict(source="SCS(proj_dual_cone) of oracle/_ref/libscsindir.so (cones.c:1544-1588)", cases=cones)
    # ---- scs_solve_lin_sys on seeded inputs
    lib.scs_init_lin_sys_work.restype = C.c_void_p
    lib.scs_init_lin_sys_work.argtypes = [C.POINTER(Mat), C.POINTER(Mat), C.POINTER(C.c_double)]
    lib.scs_solve_lin_sys.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double]
    lib.scs_free_lin_sys_work.argtypes = [C.c_void_p]
    ls = []
    for (m, n, dens, withP, seed) in [(40, 15, 0.3, False, 1), (60, 25, 0.2, True, 2)]:
        rng = np.random.RandomState(seed)
        A = sp.random(m, n, density=dens, format="csc", random_state=rng, data_rvs=rng.randn)
        P = None
        if withP:
            Q = sp.random(n, n, density=0.1, format="csc", random_state=rng, data_rvs=rng.randn)
            P = sp.triu(Q @ Q.T + 0.1 * sp.eye(n), format="csc")
        diag_r = np.concatenate([np.full(n, 1e-6), np.full(m, 10.0), [10.0]]); diag_r[n:n + 4] = 100.0
        MA, k1 = cmat(A)
        if P is not None:
            MP, k2 = cmat(P)
        w = lib.scs_init_lin_sys_work(C.byref(MA), C.byref(MP) if P is not None else None, dp(diag_r))
        b = rng.randn(n + m); s = rng.randn(n)
        bw = b.copy(); lib.scs_solve_lin_sys(w, dp(bw), dp(s), 1e-12)
        bc = b.copy(); lib.scs_solve_lin_sys(w, dp(bc), None, 1e-12)
        A2 = sp.csc_matrix(A); A2.sort_indices()
        rec = dict(m=m, n=n, Ax=jf(A2.data), Ai=jf(A2.indices), Ap=jf(A2.indptr), diag_r=jf(diag_r), b=jf(b), s=jf(s),
                   tol=1e-12, out_warm=jf(bw), out_cold=jf(bc))
        if P is not None:
            P2 = sp.csc_matrix(P); P2.sort_indices()
            rec.update(Px=jf(P2.data), Pi=jf(P2.indices), Pp=jf(P2.indptr))
        ls.append(rec)
        lib.scs_free_lin_sys_work(w)
    out["solve_lin_sys"] = dict(source="scs_solve_lin_sys of the CPU indirect backend (cpu/indirect/private.c:276-316)", cases=ls)
    # ---- aa_apply / aa_safeguard sequences on a fixed contraction
    lib.aa_init.restype = C.c_void_p
    lib.aa_init.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]
    lib.aa_apply.restype = C.c_double
    lib.aa_apply.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p]
    lib.aa_safeguard.restype = C.c_int
    lib.aa_safeguard.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p]
    lib.aa_finish.argtypes = [C.c_void_p]
    aas = []
    for (dim, mem, type1, relax, seed) in [(30, 5, 1, 1.0, 5), (30, 5, 0, 1.0, 6), (40, 10, 1, 0.8, 7)]:
        rng = np.random.RandomState(seed)
        Mx = rng.randn(dim, dim) / np.sqrt(dim) * 0.5
        bv = rng.randn(dim)
        x0 = rng.randn(dim)
        a = lib.aa_init(dim, mem, mem, type1, 1e-8, relax, 1.0, 1e10, 5, 0)
        x = x0.copy()
        norms, sgs = [], []
        for it in range(25):
            f = Mx @ x + bv
            fa = f.copy()
            norms.append(lib.aa_apply(dp(fa), dp(x), a))
            xn = fa
            fn = Mx @ xn + bv
            sgs.append(lib.aa_safeguard(dp(fn), dp(xn), a))
            x = xn
        lib.aa_finish(a)
        aas.append(dict(dim=dim, mem=mem, type1=type1, relaxation=relax, M=jf(Mx), b=jf(bv), x0=jf(x0), iters=25,
                        aa_norms=jf(norms), safeguards=sgs, x_final=jf(x)))
    out["aa"] = dict(source="aa_apply / aa_safeguard of oracle/_ref/libscsindir.so (aa.c:822-901)", cases=aas)
    # ---- full solves of the compiled reference on seeded problems (tests/problems.py generator)
    runs = []
    specs = [
        ("lp", dict(z=5, l=20), 15, 0.3, 1, False),
        ("qp", dict(z=5, l=20), 15, 0.3, 1, True),
        ("mixed", dict(z=5, l=20, q=[3, 4, 9], ep=4, ed=3, p=[0.3, -0.6]), 30, 0.3, 1, False),
        ("all_cones_P", dict(z=5, l=20, q=[3, 4, 0, 1, 9], s=[1, 2, 4, 0, 7], cs=[2, 3], ep=4, ed=3, p=[0.3, -0.6]), 40, 0.3, 2, True),
        ("box", dict(z=3, l=4, bu=[1.0, 2.0, 0.5, 3.0], bl=[-1.0, 0.5, -0.5, -2.0], q=[4]), 12, 0.4, 3, True),
        ("cfg1_small", dict(z=0, l=300, q=[10] * 24, ep=20), 200, 0.05, 1234, False),
    ]
    for name, K, n, dens, seed, withP in specs:
        data, p_star = problems.gen_feasible(K, n, dens, seed, with_P=withP)
        rec = dict(name=name, cone=K, n=n, density=dens, seed=seed, with_P=withP, p_star=p_star, runs={})
        for eps in (1e-4, 1e-9):
            for lsname, lsv in (("qdldl", scs.LinearSolver.QDLDL), ("cpu_indirect", scs.LinearSolver.CPU_INDIRECT)):
                sol = scs.SCS(data, K, linear_solver=lsv, verbose=False, eps_abs=eps, eps_rel=eps, max_iters=100000).solve()
                i = sol["info"]
                rec["runs"]["%s_%g" % (lsname, eps)] = jf(dict(status=i["status"], status_val=i["status_val"], iter=i["iter"],
                                                               pobj=i["pobj"], dobj=i["dobj"], res_pri=i["res_pri"],
                                                               res_dual=i["res_dual"], gap=i["gap"],
                                                               x=sol["x"] if n <= 40 else None))
        runs.append(rec)
    # infeasible / unbounded seeded LPs (constructions of test/gen_random_cone_prob.py:27-61)
    for kind in ("infeasible", "unbounded"):
        rng = np.random.RandomState(21)
        K = dict(z=3, l=15, q=[4])
        m, n = problems.cone_len(K), 10
        from oracle import scs_oracle as O
        z = rng.randn(m)
        y = z.copy(); O.proj_dual_cone(y, O.ConeWork(K, m), None, None)
        if kind == "infeasible":   # A'y = 0, b'y = -1
            A = sp.random(m, n, density=0.5, format="csc", random_state=rng, data_rvs=rng.randn).toarray()
            A = A - np.outer(y, A.T @ y) / float(y @ y)
            b = rng.randn(m); b = -b / float(b @ y)
            data = dict(A=sp.csc_matrix(A), b=b, c=rng.randn(n))
        else:                      # A x + s = 0, c'x = -1
            s = y - z
            A = sp.random(m, n, density=0.5, format="csc", random_state=rng, data_rvs=rng.randn).toarray()
            x = rng.randn(n)
            A = A - np.outer(s + A @ x, x) / float(x @ x)
            c = rng.randn(n); c = -c / float(c @ x)
            data = dict(A=sp.csc_matrix(A), b=rng.randn(m), c=c)
        sol = scs.SCS(data, K, linear_solver=scs.LinearSolver.QDLDL, verbose=False, eps_abs=1e-7, eps_rel=1e-7).solve()
        i = sol["info"]
        A2 = sp.csc_matrix(data["A"]); A2.sort_indices()
        runs.append(dict(name=kind, cone=K, m=m, n=n, Ax=jf(A2.data), Ai=jf(A2.indices), Ap=jf(A2.indptr), b=jf(data["b"]),
                         c=jf(data["c"]), runs={"qdldl_1e-07": jf(dict(status=i["status"], status_val=i["status_val"],
                                                                         iter=i["iter"], pobj=i["pobj"], dobj=i["dobj"]))}))
    out["solves"] = dict(source="scs.SCS(...).solve() of oracle/_ref (QDLDL and CPU_INDIRECT), problems from tests/problems.py",
                         cases=runs)
    return out


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("make_golden.py needs /root/reference (build container only)")
    kat = build_kat()
    json.dump(jf(kat), open(os.path.join(HERE, "kat.json"), "w"), indent=0)
    print("kat.json: %d header problems, %d file problems" % (len(kat["problems"]), len(kat["file_problems"])))
    rr = build_ref_runs()
    json.dump(jf(rr), open(os.path.join(HERE, "ref_runs.json"), "w"), indent=0)
    print("ref_runs.json:", {k: len(v["cases"]) for k, v in rr.items()})
