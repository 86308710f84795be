"""Developer diagnostics run on the GPU box (not a pytest file): prints component errors."""
import sys, os, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, scipy.sparse as sp
from oracle import scs_oracle as O
from tests import problems
import scs_python_b200 as scsb
from scs_python_b200 import _scs_b200 as B
lib = B.lib

def mat(Acsc):
    A = sp.csc_matrix(Acsc); A.sort_indices()
    x = np.ascontiguousarray(A.data, dtype=np.float64); i = A.indices.astype(np.int32); p = A.indptr.astype(np.int32)
    return B.make_matrix(x, i, p, A.shape[0], A.shape[1]), (x, i, p)

def t_spmv():
    rng = np.random.RandomState(0)
    for (m, n, dens) in [(50, 30, 0.3), (3000, 2000, 0.01), (20000, 100, 0.9), (5, 40000, 0.5)]:
        A = sp.random(m, n, density=dens, format="csc", random_state=rng, data_rvs=rng.randn)
        M, keep = mat(A)
        x = rng.randn(n); y0 = rng.randn(m); y = y0.copy()
        rc = lib.scs_b200_accum_by_a(C.byref(M), B._dptr(x), B._dptr(y))
        e1 = np.max(np.abs(y - (y0 + A @ x)))
        xt = rng.randn(m); z0 = rng.randn(n); z = z0.copy()
        rc2 = lib.scs_b200_accum_by_atrans(C.byref(M), B._dptr(xt), B._dptr(z))
        e2 = np.max(np.abs(z - (z0 + A.T @ xt)))
        print("spmv", (m, n, dens), rc, rc2, "errA %.2e errAt %.2e" % (e1, e2))
    n = 500
    Pf = sp.random(n, n, density=0.02, format="csc", random_state=rng, data_rvs=rng.randn); Pf = Pf + Pf.T + sp.eye(n)
    Pu = sp.triu(Pf, format="csc"); M, keep = mat(Pu)
    x = rng.randn(n); y0 = rng.randn(n); y = y0.copy()
    rc = lib.scs_b200_accum_by_p(C.byref(M), B._dptr(x), B._dptr(y))
    print("spmv P", rc, "err %.2e" % np.max(np.abs(y - (y0 + Pf @ x))))

def t_linsys():
    rng = np.random.RandomState(1)
    for (m, n, dens, withP) in [(68, 30, 0.3, False), (600, 200, 0.05, True), (6000, 2000, 0.005, True)]:
        A = sp.random(m, n, density=dens, format="csc", random_state=rng, data_rvs=rng.randn)
        P = None
        if withP:
            Q = sp.random(n, n, density=0.01, format="csc", random_state=rng, data_rvs=rng.randn)
            P = sp.triu(Q @ Q.T + 0.1 * sp.eye(n), format="csc")
        diag_r = np.concatenate([np.full(n, 1e-6), np.full(m, 10.0), [10.0]]); diag_r[n:n + 5] = 1.0 / (1000 * 0.1)
        MA, k1 = mat(A)
        MP, k2 = mat(P) if P is not None else (None, None)
        w = lib.scs_init_lin_sys_work(C.byref(MA), C.byref(MP) if MP is not None else None, B._dptr(diag_r))
        assert w
        ls = O.LinSys(A, P, diag_r)
        for tol in [1e-12, 1e-9, 1e-4]:
            b = rng.randn(n + m); s = rng.randn(n)
            b1 = b.copy(); b2 = b.copy()
            t = time.time()
            rc = lib.scs_solve_lin_sys(w, B._dptr(b1), B._dptr(s), tol)
            t = time.time() - t
            ls.solve(b2, s.copy(), tol)
            # true solve
            K = sp.bmat([[sp.diags(diag_r[:n]) + (ls.Pfull if P is not None else 0 * sp.eye(n)), A.T], [A, -sp.diags(diag_r[n:n + m])]], format="csc")
            import scipy.sparse.linalg as sla
            xs = sla.spsolve(K, b)
            print("linsys", (m, n), "tol", tol, "rc", rc, "vs oracle %.2e vs exact %.2e (oracle vs exact %.2e) cg_its %d %.1fms" % (
                np.max(np.abs(b1 - b2)), np.max(np.abs(b1 - xs)), np.max(np.abs(b2 - xs)), lib.scs_b200_lin_sys_cg_its(w), t * 1e3))
        b = rng.randn(n + m); b1 = b.copy(); b2 = b.copy()
        lib.scs_solve_lin_sys(w, B._dptr(b1), None, 1e-12); ls.solve(b2, None, 1e-12)
        print("linsys cold", np.max(np.abs(b1 - b2)))
        bz = np.zeros(n + m); lib.scs_solve_lin_sys(w, B._dptr(bz), None, 1e-12); print("zero rhs ->", np.max(np.abs(bz)))
        lib.scs_free_lin_sys_work(w)

def t_cones():
    rng = np.random.RandomState(2)
    Ks = [dict(z=3, l=5), dict(q=[1, 2, 3, 5, 40, 3000]), dict(ep=50, ed=50), dict(p=[0.3, -0.6, 0.5, 0.9, -0.1] * 10),
          dict(s=[1, 2, 3, 6, 20]), dict(s=[50]), dict(cs=[1, 2, 3, 5]), dict(bu=list(np.abs(rng.randn(20)) + 0.1), bl=list(-np.abs(rng.randn(20)) - 0.1)),
          dict(z=2, l=3, bu=[1.0, 2.0], bl=[-1.0, 0.5], q=[3, 4], s=[3], cs=[2], ep=2, ed=2, p=[0.4, -0.7])]
    for K in Ks:
        m = problems.cone_len(K)
        k, keep = B.make_cone(K)
        w = lib.scs_b200_init_cone(C.byref(k), m); assert w
        cw = O.ConeWork(K, m)
        for trial in range(3):
            x = rng.randn(m) * (10.0 ** rng.randint(-2, 3))
            r_y = np.full(m, 10.0); r_y[:K.get("z", 0)] = 0.01
            if trial == 2: r_y = np.abs(rng.randn(m)) + 0.5 if "bu" in K else r_y
            x1 = x.copy(); x2 = x.copy()
            t = time.time(); rc = lib.scs_b200_proj_dual_cone(B._dptr(x1), w, None, B._dptr(r_y)); t = time.time() - t
            O.proj_dual_cone(x2, cw, None, r_y)
            print("cone", {kk: (vv if not isinstance(vv, list) or len(vv) < 6 else "[%d]" % len(vv)) for kk, vv in K.items()}, "rc", rc,
                  "err %.2e rel %.2e  %.1fms" % (np.max(np.abs(x1 - x2)), np.max(np.abs(x1 - x2)) / max(1e-300, np.max(np.abs(x2))), t * 1e3))
        lib.scs_b200_finish_cone(w)

def t_aa():
    rng = np.random.RandomState(3)
    for (dim, mem, type1, relax) in [(50, 5, 1, 1.0), (5000, 10, 1, 1.0), (5000, 10, 0, 1.0), (700, 10, 1, 0.7), (3, 10, 1, 1.0)]:
        w = lib.scs_b200_aa_init(dim, mem, mem, type1, 1e-8, relax, 1.0, 1e10, 5); assert w
        a = O.AaWork(dim, mem, mem, type1, 1e-8, relax)
        Mx = rng.randn(dim, dim) / np.sqrt(dim) * 0.5 if dim <= 700 else None
        d = rng.rand(dim) * 0.9
        bvec = rng.randn(dim)
        fmap = (lambda x: Mx @ x + bvec) if Mx is not None else (lambda x: d * x + bvec)
        x1 = rng.randn(dim); x2 = x1.copy()
        errs = []
        for it in range(40):
            f1 = fmap(x1); f2 = fmap(x2)
            f1a = f1.copy(); f2a = f2.copy()
            n1 = lib.scs_b200_aa_apply(B._dptr(f1a), B._dptr(x1), w)
            n2 = a.apply(f2a, x2)
            errs.append((np.max(np.abs(f1a - f2a)), n1, n2))
            xn1 = f1a; xn2 = f2a
            fn1 = fmap(xn1); fn2 = fmap(xn2)
            r1 = lib.scs_b200_aa_safeguard(B._dptr(fn1), B._dptr(xn1), w)
            r2 = a.safeguard(fn2, xn2)
            if r1 != r2: print("  safeguard mismatch", it, r1, r2)
            x1, x2 = xn1, xn2
        st = lib.scs_b200_aa_get_stats(w)
        print("aa", (dim, mem, type1, relax), "max err %.2e" % max(e[0] for e in errs), "final norms", errs[-1][1:], "resid %.2e %.2e" % (np.linalg.norm(x1 - fmap(x1)), np.linalg.norm(x2 - fmap(x2))),
              "stats", st.n_accept, st.n_reject_rank0, st.n_reject_weight_cap, st.n_safeguard_reject, "| oracle", a.stats["n_accept"], a.stats["n_reject_rank0"], a.stats["n_reject_weight_cap"], a.stats["n_safeguard_reject"])
        lib.scs_b200_aa_finish(w)

def t_solve():
    cases = [
        (dict(z=5, l=20), 15, 0.3, 1, False, {}),
        (dict(z=5, l=20), 15, 0.3, 1, True, {}),
        (dict(z=5, l=20, q=[3, 4, 9], ep=4, ed=3, p=[0.3, -0.6]), 30, 0.3, 1, False, {}),
        (dict(z=5, l=20, q=[3, 4, 0, 1, 9], s=[1, 2, 4, 0, 7], cs=[2, 3], ep=4, ed=3, p=[0.3, -0.6]), 40, 0.3, 2, True, {}),
        (dict(z=0, l=300, q=[10] * 24, ep=20), 200, 0.05, 1234, False, {}),
        (dict(z=0, l=3000, q=[10] * 240, ep=200), 2000, 0.005, 1234, False, {}),
    ]
    for K, n, dens, seed, withP, kw in cases:
        data, p_star = problems.gen_feasible(K, n, dens, seed, with_P=withP)
        for eps in [1e-4, 1e-9]:
            kws = dict(eps_abs=eps, eps_rel=eps, verbose=False, max_iters=20000)
            t = time.time(); sol = scsb.SCS(data, K, **kws).solve(); t = time.time() - t
            i = sol["info"]
            t2 = time.time(); o = O.ScsOracle(data, K, **{k: v for k, v in kws.items() if k != "verbose"}).solve() if n <= 200 else None; t2 = time.time() - t2
            print("solve n=%d eps=%g P=%s:" % (n, eps, withP), i["status"], i["iter"], "pobj %.9f dobj %.9f" % (i["pobj"], i["dobj"]), "p* %s" % p_star,
                  "res %.1e %.1e gap %.1e" % (i["res_pri"], i["res_dual"], i["gap"]), "acc/rej %d/%d su %d" % (i["accepted_accel_steps"], i["rejected_accel_steps"], i["scale_updates"]),
                  "t=%.2fs (lin %.0f cone %.0f aa %.0f ms)" % (t, i["lin_sys_time"], i["cone_time"], i["accel_time"]),
                  ("| oracle %s %d %.9f" % (o["info"]["status"], o["info"]["iter"], o["info"]["pobj"])) if o else "")

if __name__ == "__main__":
    which = sys.argv[1:] or ["spmv", "linsys", "cones", "aa", "solve"]
    for w in which:
        t0 = time.time()
        try:
            globals()["t_" + w]()
        except Exception as e:
            import traceback; traceback.print_exc()
        print("== %s done in %.1fs" % (w, time.time() - t0), flush=True)
