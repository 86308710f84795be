"""CPU tier: host-side logic -- workload builders, bench.py plumbing and the world_size-2
rank sharding used for N > 1 (gloo)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lasso_builder_matches_scipy_assembly():
    from scs_python_b200 import problems
    n0, m0, k = 50, 100, 7
    data, cone, aux = problems.lasso(n0, m0, k, seed=3)
    A = data["A"]
    assert A.shape == (m0 + 2 * n0, 2 * n0 + m0) and cone == dict(z=m0, l=2 * n0)
    assert A.nnz == n0 * k + m0 + 4 * n0
    B = A.copy(); B.sort_indices()
    assert np.array_equal(A.indices, B.indices) and np.array_equal(A.indptr, B.indptr)  # already sorted CSC
    Ad = A[:m0, :n0]
    assert np.all(np.diff(Ad.indptr) == k)
    for j in range(n0):
        rows = Ad.indices[Ad.indptr[j]:Ad.indptr[j + 1]]
        assert len(set(rows.tolist())) == k
    I_n, I_m = sp.eye(n0, format="csc"), sp.eye(m0, format="csc")
    ref = sp.bmat([[Ad, -I_m, sp.csc_matrix((m0, n0))], [I_n, sp.csc_matrix((n0, m0)), -I_n],
                   [-I_n, sp.csc_matrix((n0, m0)), -I_n]], format="csc")
    assert abs(A - ref).max() == 0.0
    P = data["P"]
    assert abs(P - sp.block_diag([sp.csc_matrix((n0, n0)), I_m, sp.csc_matrix((n0, n0))])).max() == 0.0
    assert np.allclose(data["c"][n0 + m0:], aux["lam"]) and np.all(data["c"][:n0 + m0] == 0)


def test_lasso_builder_is_seeded():
    from scs_python_b200 import problems
    a, _, _ = problems.lasso(20, 40, 5, seed=1)
    b, _, _ = problems.lasso(20, 40, 5, seed=1)
    c, _, _ = problems.lasso(20, 40, 5, seed=2)
    assert abs(a["A"] - b["A"]).max() == 0 and abs(a["A"] - c["A"]).max() > 0


def test_gen_feasible_is_feasible():
    from oracle import scs_oracle as O
    from tests import problems
    K = dict(z=2, l=3, q=[3], s=[2], ep=1, ed=1, p=[0.4])
    data, p_star = problems.gen_feasible(K, 6, 0.5, seed=4)
    sol = O.ScsOracle(data, K, eps_abs=1e-8, eps_rel=1e-8).solve()
    assert sol["info"]["status_val"] == 1 and abs(sol["info"]["pobj"] - p_star) < 1e-5


_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np
import torch, torch.distributed as td
import bench
td.init_process_group(backend="gloo")
rank, world = td.get_rank(), td.get_world_size()
dev = torch.device("cpu")
# N > 1 is strong scaling: every rank holds the SAME problem and keeps its block of rows / columns
data, cone = bench.workload(0.001, seed=0)
from scs_python_b200 import _scs_b200 as B
A, P = data["A"], data["P"]
L = B.dist_local(A.shape, A.data, A.indices.astype(np.int32), A.indptr.astype(np.int32), P.data,
                 P.indices.astype(np.int32), P.indptr.astype(np.int32), data["b"], data["c"], cone, rank, world)
rows = bench.sum_over_ranks(td, dev, L["m"])
priv = bench.sum_over_ranks(td, dev, len(L["loc2glob"]) - L["n_sh"])
iters, ms = 25 * 4, 10.0 * (rank + 1)
mx = bench.max_over_ranks(td, dev, ms)
if rank == 0:
    import json
    print(json.dumps(dict(world=world, rows=rows, m=int(A.shape[0]), n=int(A.shape[1]), n_sh=int(L["n_sh"]), priv=priv,
                          max_ms=mx, value=iters / (mx * 1e-3))))
td.barrier()
td.destroy_process_group()
'''


def test_rank_sharding_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % dict(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29631", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["world"] == 2 and out["max_ms"] == 20.0
    assert out["rows"] == out["m"] and out["n_sh"] + out["priv"] == out["n"]  # the ranks' blocks tile the problem
    assert abs(out["value"] - 100 / 0.020) < 1e-6   # strong scaling: the job's iterations / slowest rank


def test_bench_reference_arm_runs_on_cpu():
    """`bench.py --impl reference` is CPU-only and must print one JSON line (tiny scale here)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup",
                        "1", "--scale", "0.002", "--iters-per-step", "5"], capture_output=True, text=True, timeout=600,
                       cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert out["impl"] == "reference" and out["metric"] == "admm_iters_per_sec" and out["value"] > 0
    assert out["e2e"]["h2d_bytes_per_step"] == 0 and out["cpu_baseline"]["kind"] in ("reference", "port")


# ------------------------------------------------------------------ row partition (dist) --
def _partitions(data, K, world):
    from scs_python_b200 import _scs_b200 as B
    A = data["A"]
    return [B.dist_partition(A.shape, A.data, A.indices, A.indptr, data["b"], data["c"], K, r, world) for r in range(world)]


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_row_partition_is_cone_aligned_and_balanced(world):
    """The blocks tile [0, m), never split a second-order / exponential cone, carry every cone
    exactly once and hold about the same number of non-zeros."""
    from scs_python_b200 import problems as P
    data, K, _ = P.random_cone_qp(seed=3, n=300, l=400, nq=40, q=7, ep=30, density=0.05)
    parts = _partitions(data, K, world)
    m, nnz = data["A"].shape[0], data["A"].nnz
    assert parts[0]["row0"] == 0 and sum(p["m"] for p in parts) == m
    for a, b in zip(parts[:-1], parts[1:]):
        assert a["row0"] + a["m"] == b["row0"]
    assert sum(p["nnz"] for p in parts) == nnz
    assert sum(p["l"] for p in parts) == K["l"] and sum(p["qsize"] for p in parts) == len(K["q"])
    assert sum(p["ep"] for p in parts) == K["ep"]
    for p in parts:  # rows of a block = rows of the cones it holds
        assert p["m"] == p["z"] + p["l"] + 7 * p["qsize"] + 3 * p["ep"]
        assert p["nnz"] <= nnz / world * 1.25 + 50


def test_row_partition_lasso_and_errors():
    from scs_python_b200 import problems as P
    data, K, _ = P.lasso(500, 1000, 10, seed=0)
    parts = _partitions(data, K, 4)
    assert sum(p["z"] for p in parts) == K["z"] and sum(p["l"] for p in parts) == K["l"]
    # a single PSD cone cannot be cut
    d4, K4, _ = P.maxcut_sdp(nodes=6, blocks=1)
    with pytest.raises(ValueError):
        _partitions(d4, K4, 2)


_DIST_WORKER = r'''
import os, sys, json
sys.path.insert(0, %(root)r)
import torch.distributed as td
td.init_process_group(backend="gloo")
rank, world = td.get_rank(), td.get_world_size()
from scs_python_b200 import _scs_b200 as B, problems as P
data, K, _ = P.random_cone_qp(seed=5, n=120, l=200, nq=20, q=5, ep=10, density=0.1)
A = data["A"]
mine = B.dist_partition(A.shape, A.data, A.indices, A.indptr, data["b"], data["c"], K, rank, world)
box = [None] * world
td.all_gather_object(box, mine)
if rank == 0:
    print(json.dumps(dict(world=world, parts=box, m=int(A.shape[0]))))
td.barrier()
td.destroy_process_group()
'''


def test_row_partition_world_size_2_gloo(tmp_path):
    """Two processes (gloo) each derive their own block from the full problem, as the ranks of a
    row-partitioned solve do; together the blocks tile the rows."""
    script = tmp_path / "dist_worker.py"
    script.write_text(_DIST_WORKER % dict(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29641", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    p0, p1 = out["parts"]
    assert out["world"] == 2 and p0["row0"] == 0 and p0["m"] == p1["row0"] and p0["m"] + p1["m"] == out["m"]


def test_solve_batch_wrapper_fills_the_c_structures(monkeypatch):
    """The batch wrapper fills five contiguous arrays of C structures through integer addresses; read them
    back through the typed-pointer view the library gets, without a device (the C entry point is replaced
    by a checker that also writes into the solution block)."""
    import scipy.sparse as sp
    import scs_python_b200 as scsb
    from scs_python_b200 import _scs_b200 as B
    rng = np.random.RandomState(0)
    probs = []
    for i in range(5):
        m, n = 7 + i, 3 + i
        A = sp.random(m, n, density=0.6, format="csc", random_state=rng, data_rvs=rng.randn)
        d = dict(A=A, b=rng.randn(m), c=rng.randn(n))
        if i % 2 == 0:
            d["P"] = sp.eye(n, format="csc") * (1.0 + i)
        cone = dict(z=1, l=m - 4, q=[3]) if i != 3 else dict(l=m - 3, ep=1)   # i == 3: general cone path
        probs.append((d, cone))
    prepared = [scsb._prepare(d, k) for d, k in probs]
    seen = {}

    def fake(cnt, pd, pk, st, ps, infos, z):
        assert cnt == 5 and st._obj.max_iters == 77      # C.byref(settings)
        for i in range(cnt):
            d, k, s = pd[i].contents, pk[i].contents, ps[i].contents
            A = d.A.contents
            nnz = A.p[A.n]
            seen[i] = dict(m=d.m, n=d.n, Am=A.m, An=A.n, Ax=[A.x[j] for j in range(nnz)], Ai=[A.i[j] for j in range(nnz)],
                           b=[d.b[j] for j in range(d.m)], c=[d.c[j] for j in range(d.n)],
                           P=[d.P.contents.x[j] for j in range(d.P.contents.p[d.n])] if d.P else None,
                           z=k.z, l=k.l, q=[k.q[j] for j in range(k.qsize)], ep=k.ep, bsize=k.bsize, ssize=k.ssize)
            s.x[0], s.y[d.m - 1], s.s[0] = 1.5 + i, 2.5 + i, 3.5 + i
            infos[i].iter = 10 + i
        return 0
    monkeypatch.setattr(B.lib, "scs_b200_solve_batch", fake)
    res = B.solve_batch(prepared, verbose=False, max_iters=77)
    for i, ((d, cone), pr) in enumerate(zip(probs, prepared)):
        shape, Ax, Ai, Ap, Px, Pi, Pp, b, c, _ = pr
        e = seen[i]
        assert (e["m"], e["n"], e["Am"], e["An"]) == (shape[0], shape[1], shape[0], shape[1])
        assert e["Ax"] == list(Ax) and e["Ai"] == list(Ai) and e["b"] == list(b) and e["c"] == list(c)
        assert e["P"] == (list(Px) if Px is not None else None)
        assert (e["z"], e["l"], e["q"], e["ep"]) == (cone.get("z", 0), cone.get("l", 0), list(cone.get("q", [])), cone.get("ep", 0))
        assert e["bsize"] == 0 and e["ssize"] == 0
        assert res[i]["x"][0] == 1.5 + i and res[i]["y"][-1] == 2.5 + i and res[i]["s"][0] == 3.5 + i
        assert res[i]["x"].shape == (shape[1],) and res[i]["y"].shape == (shape[0],) and res[i]["info"]["iter"] == 10 + i
    with pytest.raises(ValueError):
        B.solve_batch([prepared[0][:8] + (np.zeros(2), prepared[0][9])], verbose=False)   # c of the wrong length
    with pytest.raises(TypeError):
        B.solve_batch([prepared[0][:1] + (np.zeros(3, dtype=np.int32),) + prepared[0][2:]], verbose=False)  # Ax not float


# ------------------------------------------------------------ tiled SpMV engine: the host plan ------
def _tile_geometry():
    import ctypes as C
    from scs_python_b200 import _scs_b200 as B
    g = (B.c_int * 4)()
    B.lib.scs_b200_tiled_geometry(C.byref(g))
    return int(g[0]), int(g[1])


TR, TC = _tile_geometry()   # csrc/tiled.cuh: rows per row bin, columns per column bin of this build


def _tiled_plan(nrows, hg, p1, p2, sms):
    from scs_python_b200 import _scs_b200 as B
    nrb, ncb = hg.shape
    hg = np.ascontiguousarray(hg, dtype=np.int32)
    p1 = np.ascontiguousarray(p1, dtype=np.int32)
    p2c = np.ascontiguousarray(p2, dtype=np.int32) if p2 is not None else None
    cap = int(nrb * ncb + nrb + sms + 8)
    items, seq = np.zeros(5 * cap, dtype=np.int32), np.zeros(cap, dtype=np.int32)
    binfo, cost = np.zeros(2 * nrb, dtype=np.int32), np.zeros(sms)
    ncta, nseq = B.c_int(0), B.c_int(0)
    import ctypes as C
    n = B.lib.scs_b200_tiled_plan(nrows, ncb, B._iptr(hg), B._iptr(p1), B._iptr(p2c) if p2c is not None else None, sms,
                                  B._iptr(items), cap, B._iptr(seq), cap, B._iptr(binfo), B._dptr(cost),
                                  C.byref(ncta), C.byref(nseq))
    assert n >= 0
    return items[:5 * n].reshape(-1, 5), seq[:nseq.value], binfo.reshape(-1, 2), cost[:ncta.value]


@pytest.mark.parametrize("seed,nrb,ncb,sms", [(0, 9, 40, 148), (1, 3, 7, 148), (2, 30, 25, 16), (3, 1, 1, 148), (4, 6, 300, 148)])
def test_tiled_plan_covers_every_cell_once_and_balances(seed, nrb, ncb, sms):
    """Invariants of the partition the streaming kernel relies on (csrc/tiled.cu: tiled_plan_host): every active
    cell of every tiled row bin is visited exactly once, in ascending column order inside a piece; the pieces
    of a row bin own consecutive scratch slots in piece order (the epilogue adds them in that order); short-row
    and empty bins get no item; CTA ranges are contiguous in (row bin, column) order; the modelled cost of the
    busiest CTA is within one cell (+ the snapping slack of two cells) of the mean."""
    rng = np.random.RandomState(seed)
    nrows = nrb * TR - (rng.randint(1, TR) if nrb > 1 else TR - 5000)
    kind = rng.randint(0, 3, size=nrb)                 # 0: heavy rows, 1: short rows (direct bin), 2: empty
    if nrb > 1:
        kind[0] = 0
    rowlen = np.zeros(nrows, dtype=np.int64)
    hg = np.zeros((nrb, ncb), dtype=np.int32)
    for rb in range(nrb):
        r0, r1 = rb * TR, min(nrows, (rb + 1) * TR)
        if kind[rb] == 0:
            rowlen[r0:r1] = rng.randint(20, 60, size=r1 - r0)
            active = rng.rand(ncb) < 0.7
            active[rng.randint(ncb)] = True
            hg[rb, active] = rng.randint(16, 200, size=int(active.sum()))
        elif kind[rb] == 1:
            rowlen[r0:r1] = rng.randint(0, 3, size=r1 - r0)
            hg[rb, rng.randint(ncb)] = 16          # cells exist in the format but must not be scheduled
    p1 = np.concatenate([[0], np.cumsum(rowlen)])
    items, seq, binfo, cost = _tiled_plan(nrows, hg, p1, None, sms)
    tiled_bins = [rb for rb in range(nrb) if kind[rb] == 0]
    # coverage, order, slots
    seen = {rb: [] for rb in tiled_bins}
    pieces = {rb: [] for rb in tiled_bins}
    prev_key = (-1, -1)
    assert list(items[:, 0]) == sorted(items[:, 0]) and (len(items) == 0 or items[:, 0].max() < len(cost) <= sms)
    for cta, rb, slot, a0, a1 in items:
        assert rb in seen, "a short-row / empty bin was scheduled"
        cells = list(seq[a0:a1])
        assert cells == sorted(cells) and len(set(cells)) == len(cells) and len(cells) > 0
        assert (rb, cells[0]) > prev_key                 # contiguous ranges in (row bin, column) order
        prev_key = (rb, cells[-1])
        seen[rb] += cells
        pieces[rb].append(slot)
    for rb in tiled_bins:
        assert seen[rb] == [cb for cb in range(ncb) if hg[rb, cb] > 0]
        assert binfo[rb, 1] == len(pieces[rb]) and pieces[rb] == list(range(binfo[rb, 0], binfo[rb, 0] + binfo[rb, 1]))
    for rb in range(nrb):
        if kind[rb] != 0:
            assert binfo[rb, 1] == 0
    slots = sorted(s for rb in tiled_bins for s in pieces[rb])
    assert slots == list(range(len(slots)))              # scratch slots are dense and unique
    # balance (cost model of tiled_plan_host: max(384 g + 32768, 70000) per cell, 262144 per item)
    if len(items):
        cell = np.maximum(384.0 * hg[hg > 0] + 8.0 * TC, 70000.0)
        assert abs(cost.sum() - (cell[np.isin(np.nonzero(hg)[0], tiled_bins)].sum() + 16.0 * TR * len(items))) < 1e-6 * cost.sum()
        assert cost.max() <= cost.sum() / len(cost) + 3 * cell.max() + 2 * 16.0 * TR * max(1, items.shape[0] // len(cost) + 1)


def _locals(data, K, world):
    import scipy.sparse as sp
    from scs_python_b200 import _scs_b200 as B
    A = sp.csc_matrix(data["A"]); A.sort_indices()
    P = data.get("P")
    args = [A.shape, A.data, A.indices.astype(np.int32), A.indptr.astype(np.int32)]
    if P is not None:
        P = sp.triu(sp.csc_matrix(P), format="csc"); P.sort_indices()
        args += [P.data, P.indices.astype(np.int32), P.indptr.astype(np.int32)]
    else:
        args += [None, None, None]
    return A, P, [B.dist_local(*args, data["b"], data["c"], K, r, world) for r in range(world)]


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("kind", ["lasso", "cone_qp", "coupled_P"])
def test_column_classification_reassembles_the_problem(world, kind):
    """Row-partitioned mode, host side: every rank keeps its rows and the columns they touch, in the local
    order [shared | private].  Shared columns are the same on every rank, every private column belongs to
    exactly one rank and has no entry outside that rank's rows, P restricted to a rank's columns loses
    nothing, and scattering the local blocks back gives A, P and c exactly."""
    import scipy.sparse as sp
    from scs_python_b200 import problems as Pm
    if kind == "lasso":
        data, K, _ = Pm.lasso(400, 800, 10, seed=1)
    elif kind == "cone_qp":
        data, K, _ = Pm.random_cone_qp(seed=3, n=200, l=300, nq=30, q=7, ep=20, density=0.04)
    else:  # P with off-diagonal couplings: coupled columns must stay shared
        data, K, _ = Pm.lasso(300, 600, 8, seed=2)
        n = data["A"].shape[1]
        rng = np.random.RandomState(0)
        i, j = rng.randint(0, n, 40), rng.randint(0, n, 40)
        C_ = sp.csc_matrix((np.ones(40), (np.minimum(i, j), np.maximum(i, j))), shape=(n, n))
        data["P"] = sp.triu(sp.csc_matrix(data["P"]) + C_ + sp.eye(n), format="csc")
    A, P, loc = _locals(data, K, world)
    m, n = A.shape
    shared = loc[0]["loc2glob"][:loc[0]["n_sh"]]
    seen = np.zeros(n, dtype=int)
    seen[shared] += 1
    Afull = sp.lil_matrix((m, n))
    row = 0
    for r, L in enumerate(loc):
        l2g, nsh = L["loc2glob"], L["n_sh"]
        assert np.array_equal(l2g[:nsh], shared) and np.all(np.diff(l2g[:nsh]) > 0) and np.all(np.diff(l2g[nsh:]) > 0)
        seen[l2g[nsh:]] += 1
        assert L["row0"] == row
        row += L["m"]
        Al = sp.coo_matrix(L["A"])
        Afull[Al.row + L["row0"], l2g[Al.col]] = Al.data
        assert np.array_equal(L["c"], data["c"][l2g])
        # a private column has no entry outside this rank's rows
        priv = l2g[nsh:]
        if len(priv):
            assert A[:, priv].nnz == L["A"][:, nsh:].nnz
        if P is not None:
            Pl = sp.coo_matrix(L["P"])
            assert np.all(Pl.row <= Pl.col)  # still upper triangular in the local order
            ref = sp.coo_matrix(P[:, l2g][l2g, :])
            assert abs(sp.csc_matrix(L["P"]) - sp.csc_matrix(ref)).sum() == 0
            assert P[:, l2g].nnz == L["P"].nnz  # no entry of these columns of P falls outside the local set
    assert row == m and np.all(seen == 1)
    assert abs(sp.csc_matrix(Afull) - A).sum() == 0
    if world > 1 and kind == "lasso":  # the y block of LASSO (identity columns) is private: less than half is shared
        assert len(shared) < 0.6 * n


def test_batch_plan_eligibility_and_footprint():
    """scs_b200_batch_plan (host only): which members the one-CTA batch kernel takes.  Every cone of the default
    build is eligible (PSD up to order 32, complex PSD up to 16); warm starts, time limits, lookback > 10 and
    oversized footprints go to the streaming engine; invalid members are reported, not planned; the shared
    carve-up is the maximum over the members and the PSD workspaces shrink before a member is turned away."""
    import scs_python_b200 as scsb
    from scs_python_b200 import _scs_b200 as B, problems as P
    from tests import problems as tp
    prep = lambda d, k: scsb._prepare(d, k)
    mpc, box = prep(*P.mpc_qp(0)[:2]), prep(*P.mpc_qp_box(1)[:2])
    fused, plan = B.batch_plan([mpc, box], verbose=False)
    assert fused == [1, 1] and plan["direct"] == 1 and plan["extended_cones"] == 0 and plan["box_bounds"] == 120
    assert 0 < plan["smem_bytes"] <= 227 * 1024
    mk = lambda K, n=12, seed=31: prep(tp.gen_feasible(K, n, 0.3, seed, with_P=False)[0], K)
    for K, want in ((dict(z=1, l=2, s=[32]), 1), (dict(z=1, l=2, s=[33]), 0), (dict(z=1, l=2, cs=[16]), 1),
                    (dict(z=1, l=2, cs=[17]), 0), (dict(z=1, l=4, ep=2, ed=2, p=[-0.4, 0.7]), 1),
                    (dict(z=2, l=4, q=[3], s=[3, 5], cs=[2]), 1)):
        f, pl = B.batch_plan([mk(K)], verbose=False)
        assert f == [want], (K, f, pl)
        if want:
            assert pl["extended_cones"] == 1
            assert pl["psd_order"] == max([0] + list(K.get("s", [])) + [2 * c for c in K.get("cs", [])])
    # settings the kernel does not implement
    for kw in (dict(time_limit_secs=1.0), dict(acceleration_lookback=11),
               dict(acceleration_lookback=5, acceleration_relaxation=1.5)):
        assert B.batch_plan([mpc], verbose=False, **kw)[0] == [0], kw
    # a large member keeps its place by giving up PSD workspaces: order-32 embedding next to a 2.3 k-entry matrix
    K = dict(z=0, l=3, cs=[16])
    big = prep(tp.gen_feasible(K, 30, 0.3, 44, with_P=False)[0], K)
    f, pl = B.batch_plan([big], verbose=False)
    assert f == [1] and 1 <= pl["psd_workspaces"] < 8 and pl["smem_bytes"] <= 200 * 1024, pl
    # mixed batch: maxima over the members; the oversized PSD member alone is turned away
    f, pl = B.batch_plan([mpc, mk(dict(z=1, l=2, s=[33])), mk(dict(z=1, l=2, s=[6])), box], verbose=False)
    assert f == [1, 0, 1, 1] and pl["psd_order"] == 6 and pl["box_bounds"] == 120
    # a member whose cone sizes do not add up to m fails validation (reported as -1), the rest is planned
    bad = list(mk(dict(z=1, l=2, q=[3])))
    bad[-1] = dict(z=1, l=2, q=[4])
    f, pl = B.batch_plan([mpc, tuple(bad)], verbose=False)
    assert f == [1, -1]
    assert B.batch_plan([], verbose=False)[0] == []
