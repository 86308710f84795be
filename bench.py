#!/usr/bin/env python
"""bench.py -- ADMM iterations/sec of the SCS hot path on the BASELINE.json headline workload.

    python bench.py --gpus N --steps K --warmup W            # B200 backend (this repo)
    python bench.py --impl reference --steps K --warmup W    # reference CPU_INDIRECT on host cores

Workload (BASELINE.json configs[1]): sparse LASSO in the reference's
documentation formulation with data matrix Ad in R^{2M x 1M}, 100 nnz per column =>
SCS n = 4M, m = 4M, nnz(A) = 106M, nnz(P) = 2M, cone {z: 2M, l: 2M}.  Synthetic, seeded.

A "step" is `--iters-per-step` (default 25 = CONVERGED_INTERVAL) consecutive ADMM iterations of ONE
cold-start solve with the stopping tolerances at 0.  Both arms time a WINDOW of iterations of such a
solve, so the per-call work that is not an ADMM iteration (upload of b and c, the tol-1e-12 solve for
g = (R+M)^-1 h of scs.c:1066-1076, download of x, y, s) is excluded from both:

  value : iterations [W*ips, (W+K)*ips) of one scs_solve call, timed on the device with CUDA events on the
          solve stream (inputs resident in HBM); barrier + device synchronize on both sides, MAX over ranks.
  e2e   : the same window through the public API with HOST buffers and the wall clock:
          T(update(b,c) + solve(max_iters=(W+K)*ips)) - T(update(b,c) + solve(max_iters=W*ips)); both calls
          upload b, c and download x, y, s inside the timed region.  `e2e.whole_call` is the un-differenced
          rate of the long call (g solve, warm-up iterations and copies included).
  --impl reference : the reference's CPU_INDIRECT path (oracle/_ref, OpenMP build, all host cores), timed
          the same way from info.solve_time of two solves.  One repo step (25 iterations) costs the reference
          about a minute, so a reference step is a bounded SAMPLE of it: n_s consecutive iterations
          (n_s = the largest value <= 25 that keeps the run inside --ref-budget-s), window =
          [1 + W*n_s, 1 + (W+K)*n_s).  The rate is per iteration, so the step length cancels; the repo arm
          prints its own rate on the same early window as `value_on_reference_window`.
  N > 1 : ONE problem, row-partitioned over the ranks (csrc/dist.cu, scs_solver.cu partition_rows): "strong"
          scaling.  Every rank times the same iterations; value = iterations / max over ranks.  The line
          carries the collective counts / bytes and the objective agreement with a single-GPU solve.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

FULL = dict(n0=1_000_000, m0=2_000_000, nnz_per_col=100)
GATHER_CEILING_GELEMS = 272.0  # measured on B200: tools/probes/gather_probe.cu, profiles/r1b_gather_probe_x8MB.txt


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--iters-per-step", type=int, default=25)
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the full workload (dev runs only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-time-to-eps", action="store_true")
    ap.add_argument("--no-batch", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--ref-budget-s", type=float, default=240.0)
    return ap.parse_args()


def workload(scale, seed):
    from scs_python_b200 import problems
    n0 = max(1000, int(FULL["n0"] * scale))
    m0 = 2 * n0
    data, cone, aux = problems.lasso(n0, m0, FULL["nnz_per_col"], seed)
    return data, cone


def config_of(data, cone, scale, ips, n_gpus):
    """The config object of the JSON line: identical in both arms."""
    return dict(workload="sparse LASSO QP (reference doc formulation), Ad %dx%d, %d nnz/col"
                         % (cone["z"], (data["A"].shape[1] - cone["z"]) // 2, FULL["nnz_per_col"]),
                n=int(data["A"].shape[1]), m=int(data["A"].shape[0]), nnz_A=int(data["A"].nnz),
                nnz_P=int(data["P"].nnz), cone="z=%d,l=%d" % (cone["z"], cone["l"]), scale=scale,
                l2_policy="inputs_exceed_l2 (A+A' = %.2f GB >> 126 MB L2)" % (2 * 12e-9 * data["A"].nnz),
                iters_per_step=ips,
                parallelism=("single GPU" if n_gpus == 1 else
                             "one problem, A row-partitioned over %d GPUs (all-reduce of the shared block of "
                             "A_g'z_g per CG iteration over NVLink peer memory / NCCL)" % n_gpus))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.samples.append(line.strip())
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["unavailable"])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(np.max(mx)), reasons=sorted(reasons),
                    samples=len(sm))


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("spmv_g_dram_bytes_per_launch")
        except Exception:
            return None
    return None


def dist_setup():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    td = None
    if world > 1:
        import torch
        import torch.distributed as td_
        torch.cuda.set_device(local)
        td_.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
        td = td_
    return rank, world, local, td


def barrier_sync(td, local):
    import torch
    if td is not None:
        td.barrier()
    torch.cuda.synchronize(local)


def _reduce(td, local, x, op):
    if td is None:
        return float(x)
    import torch
    dev = local if isinstance(local, torch.device) else torch.device("cuda", local)
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    td.all_reduce(t, op=getattr(td.ReduceOp, op))
    return float(t.item())


def max_over_ranks(td, local, x):
    return _reduce(td, local, x, "MAX")


def sum_over_ranks(td, local, x):
    return _reduce(td, local, x, "SUM")


# ------------------------------------------------------------------------------ reference --
def import_reference():
    """Reference python package from oracle/_ref (OpenMP build first).  Returns (module, kind)."""
    for sub in ("scs_omp", ""):
        p = os.path.join(ROOT, "oracle", "_ref", sub) if sub else os.path.join(ROOT, "oracle", "_ref")
        if os.path.exists(os.path.join(p, "scs", "__init__.py")):
            sys.path.insert(0, p)
            try:
                import scs  # noqa
                return scs, ("reference (oracle/_ref/%s, CPU_INDIRECT%s)" % (sub or "scs", ", OpenMP" if sub else ""))
            except Exception:
                sys.path.pop(0)
                for k in [k for k in sys.modules if k == "scs" or k.startswith("scs.")]:
                    del sys.modules[k]
    return None, None


def reference_factory():
    """(make(data, cone, max_iters) -> object with .solve() -> (iters, solve_seconds), kind, cores, is_port).
    The host thread count is forced: torch.distributed.run pre-sets OMP_NUM_THREADS=1."""
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)
    os.environ["OPENBLAS_NUM_THREADS"] = str(cores)
    scs, kind = import_reference()
    if scs is not None:
        class _R:
            def __init__(self, data, cone, k, time_limit=0.0):
                t = time.perf_counter()
                self.s = scs.SCS(data, cone, linear_solver=scs.LinearSolver.CPU_INDIRECT, verbose=False, max_iters=int(k),
                                 eps_abs=0.0, eps_rel=0.0, eps_infeas=0.0, time_limit_secs=float(time_limit))
                self.setup_s = time.perf_counter() - t

            def solve(self):
                info = self.s.solve(warm_start=False)["info"]
                return int(info["iter"]), float(info["solve_time"]) * 1e-3
        return _R, kind, cores, False
    from oracle import scs_oracle as O  # no compiled reference on this box: the numpy restatement ("port")

    class _W:
        def __init__(self, data, cone, k, time_limit=0.0):
            t = time.perf_counter()
            self.s = O.ScsOracle(data, cone, max_iters=int(k), eps_abs=0.0, eps_rel=0.0, eps_infeas=0.0)
            self.setup_s = time.perf_counter() - t

        def solve(self):
            t = time.perf_counter()
            info = self.s.solve(False)["info"]
            return int(info["iter"]), time.perf_counter() - t
    return _W, "port (oracle/scs_oracle.py, numpy)", 1, True


def reference_window(data, cone, ips, steps, warmup, budget_s, scale=None):
    """Rate of the reference over a window of iterations of one cold-start solve, by difference of the
    solve times of two solves that stop at the window's two ends (setup and the per-call g solve cancel).
    The window is sized to the budget from a pilot on a small instance of the same family (the cost of
    setup, of the g solve and of an iteration are all linear in nnz(A))."""
    t_begin = time.perf_counter()
    make, kind, cores, port = reference_factory()
    nnz = data["A"].nnz
    if scale is not None and nnz > 4_000_000:
        pd, pc = workload(scale * 3_000_000.0 / nnz, seed=0)
    else:
        pd, pc = data, cone
    ratio = nnz / max(1, pd["A"].nnz)
    p1 = make(pd, pc, 1)
    _, t1 = p1.solve()
    est_setup = p1.setup_s * ratio
    del p1
    p2 = make(pd, pc, 26)
    _, t26 = p2.solve()
    del p2
    # iterations get dearer as the CG tolerance tightens (scs.c:703-720): 1.5x head-room on the early rate
    est_it = 1.5 * max(1e-6, (t26 - t1) / 25.0) * ratio
    est_g = t1 * ratio
    pilot_s = time.perf_counter() - t_begin
    W, K = warmup, steps
    # The pilot is pessimistic at full size (measured on the bench box: it predicted 6.9 s per early iteration and 54 s
    # for the g solve where the full-size run took 1.6 s and 36 s): a sample longer than one iteration per step is
    # only taken when the pessimistic figures allow it, and steps are only cut when even the optimistic ones
    # (x 0.3 / x 0.6) say one iteration per step does not fit.  The second solve carries time_limit_secs as a net
    # (checked by the reference every 25 iterations, scs.c:1355-1360).
    room = budget_s - pilot_s - 2 * est_setup - 2 * est_g - 2 * est_it
    n_s = int(room / max(1e-9, (2 * W + K) * est_it))
    n_s = max(1, min(ips, n_s))
    room_opt = budget_s - pilot_s - 2 * est_setup - 2 * 0.6 * est_g - 2 * 0.3 * est_it
    if n_s == 1 and room_opt < (2 * W + K) * 0.3 * est_it:  # fewer steps, never a shorter solve
        W = min(W, 1)
        K = max(2, min(K, int(room_opt / (0.3 * est_it)) - 2 * W))
    it_lo, it_hi = 1 + W * n_s, 1 + (W + K) * n_s
    lo = make(data, cone, it_lo)
    setup_s = lo.setup_s
    i_lo, t_lo = lo.solve()
    del lo
    left = budget_s - (time.perf_counter() - t_begin) - setup_s
    hi = make(data, cone, it_hi, time_limit=max(10.0, left))
    i_hi, t_hi = hi.solve()
    del hi
    if i_hi <= i_lo:  # the net cut the second solve before the window opened: nothing to report but the fact
        i_hi, t_hi = i_lo + 1, t_lo + 1e9
    K = max(1, (i_hi - i_lo) // n_s) if i_hi < it_hi else K
    dt = max(1e-9, t_hi - t_lo)
    return dict(value=(i_hi - i_lo) / dt, kind=kind, cores=cores, port=port, setup_s=setup_s, window=[i_lo, i_hi],
                t_lo_s=t_lo, t_hi_s=t_hi, steps_timed=K, warmup_timed=W, iters_per_step_timed=n_s,
                ms_per_step=1e3 * dt / max(1, K), wall_s=time.perf_counter() - t_begin,
                pilot=dict(seconds=pilot_s, est_iter_s=est_it, est_g_solve_s=est_g, est_setup_s=est_setup))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0  # N > 1: rank 0 alone runs the CPU reference
    data, cone = workload(args.scale, seed=0)
    desc = config_of(data, cone, args.scale, args.iters_per_step, args.gpus)
    r = reference_window(data, cone, args.iters_per_step, args.steps, args.warmup, args.ref_budget_s, scale=args.scale)
    sample = ("full workload, %s, %d host threads; window = ADMM iterations [%d, %d) of one cold-start solve timed as "
              "info.solve_time(max_iters=%d) - info.solve_time(max_iters=%d) = %.1f s - %.1f s, so setup (%.1f s) and the "
              "per-call g solve cancel; each of the %d timed steps is a bounded sample of %d of the %d iterations of a "
              "repo step (the rate is per iteration)"
              % (r["kind"], r["cores"], r["window"][0], r["window"][1], r["window"][1], r["window"][0], r["t_hi_s"],
                 r["t_lo_s"], r["setup_s"], r["steps_timed"], r["iters_per_step_timed"], args.iters_per_step))
    line = dict(impl="reference", metric="admm_iters_per_sec", value=r["value"], unit="iters/s", n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=r["ms_per_step"], higher_is_better=True,
                scaling="weak" if args.gpus == 1 else "strong", vs_baseline=None, dtype="f64", data="synthetic", config=desc,
                window=r["window"], steps_timed=r["steps_timed"], warmup_timed=r["warmup_timed"],
                iters_per_step_timed=r["iters_per_step_timed"], wall_s=r["wall_s"], pilot=r["pilot"],
                cpu_baseline=dict(value=r["value"], unit="iters/s", cores=r["cores"],
                                  kind="port" if r["port"] else "reference", sample=sample),
                e2e=dict(value=r["value"], unit="iters/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))
    return 0


def cpu_baseline_sample(ips):
    """Reference CPU path on a bounded 1/16-scale sample of the same workload (rank 0, N=1), same window method."""
    scale = 1.0 / 16.0
    data, cone = workload(scale, seed=0)
    r = reference_window(data, cone, ips, steps=1, warmup=0, budget_s=60.0)
    return dict(value=r["value"], unit="iters/s", cores=r["cores"], kind="port" if r["port"] else "reference",
                sample="%s on a 1/16-scale instance of the workload (n=%d, m=%d, nnz(A)=%d): iterations [%d, %d) of one "
                       "cold-start solve by difference of two solve times (setup and g solve cancel).  The per-iteration "
                       "work is linear in nnz, so 1/16 of this is an UPPER bound on the full-size rate: measured at full "
                       "size (`bench.py --impl reference`, profiles/r2n_bench_reference.json) the reference does 0.73 "
                       "it/s, about half the scaled figure, because its working set then leaves the host's last-level "
                       "cache.  The full-size figure is what `bench.py --impl reference` prints"
                       % (r["kind"], data["A"].shape[1], data["A"].shape[0], data["A"].nnz, r["window"][0], r["window"][1]),
                value_scaled_to_full_workload=r["value"] * scale)


# ------------------------------------------------------------------------------------ ours --
def batch_cfg5(scsb, rank, per_gpu=1024):
    """BASELINE.json configs[4]: 8192 independent MPC QPs (n=120, m=360) over 8 GPUs = 1024 per GPU, no
    communication.  This rank's share through the public batch call (host buffers in, solutions out)."""
    from scs_python_b200 import problems as bp, _scs_b200 as B
    probs = [bp.mpc_qp(10_000 * rank + i)[:2] for i in range(per_gpu)]
    scsb.solve_batch(probs[:64], verbose=False)  # warm-up (arena, module load)
    t = time.perf_counter()
    sols = scsb.solve_batch(probs, verbose=False)
    wall = time.perf_counter() - t
    st = B.batch_stats()
    iters = int(sum(s["info"]["iter"] for s in sols))
    solved = int(sum(1 for s in sols if s["info"]["status_val"] == 1))
    return dict(problems=per_gpu, solved=solved, wall_s=wall, kernel_ms=st.get("kernel_ms"), admm_iters=iters)


def other_configs(scsb):
    """BASELINE.json configs[0], [2] and [3] on this GPU (N = 1 only): setup + solve to the default eps 1e-4 through the
    public API with host buffers, the faster of two runs; next to each, the compiled reference's run of the same
    seeded instance from tests/golden/full_ref.json (iterations and objective: the parity bar; its times were taken
    in the build container, tools/bench_configs.py has the reference arms on the bench box)."""
    from scs_python_b200 import problems as P
    gold = {}
    try:
        gold = json.load(open(os.path.join(ROOT, "tests", "golden", "full_ref.json")))
    except Exception:
        pass
    res = {}
    for name, gname, make in (("cfg1_cone_qp_n2000_m6000", "cfg1_qp", lambda: P.random_cone_qp(seed=1234, with_P=True)),
                              ("cfg3_socp_portfolio_n50k_10k_cones", "cfg3_socp", lambda: P.socp_portfolio(seed=0)),
                              ("cfg4_maxcut_sdp_64_psd200", "cfg4_sdp", lambda: P.maxcut_sdp(seed=0))):
        try:
            d, K, _ = make()
            best = None
            for _rep in range(2):
                t = time.perf_counter()
                sv = scsb.SCS(d, K, verbose=False)
                i = sv.solve(warm_start=False)["info"]
                wall = time.perf_counter() - t
                sv._solver.finish()
                del sv
                rec = dict(status=i["status"], iters=i["iter"], wall_s=wall, setup_ms=i["setup_time"], solve_ms=i["solve_time"],
                           iters_per_s=i["iter"] / max(i["solve_time"], 1e-9) * 1e3, pobj=i["pobj"], dobj=i["dobj"],
                           res_pri=i["res_pri"], res_dual=i["res_dual"], gap=i["gap"])
                if best is None or wall < best["wall_s"]:
                    best = rec
            g = gold.get(gname, {}).get("runs", {}).get("0.0001")
            if g:
                best["reference_fixture"] = dict(iters=g["iter"], pobj=g["pobj"], status=g["status"])
            res[name] = best
        except Exception as e:  # secondary keys never take the headline line down
            res[name] = dict(error=repr(e))
    return res


def run_b200(args):
    rank, world, local, td = dist_setup()
    import scs_python_b200 as scsb
    from scs_python_b200 import _scs_b200 as B
    if B.lib.scs_b200_device_count() <= 0:
        raise RuntimeError("bench.py: no CUDA device visible; the B200 backend has no CPU fallback")
    if B.lib.scs_b200_set_device(local) != 0:
        raise RuntimeError("bench.py: cannot select CUDA device %d" % local)
    import torch
    torch.cuda.set_device(local)
    ips, K, W = args.iters_per_step, args.steps, args.warmup
    data, cone = workload(args.scale, seed=0)  # the same problem on every rank
    desc = config_of(data, cone, args.scale, ips, world)
    zero_tol = dict(verbose=False, eps_abs=0.0, eps_rel=0.0, eps_infeas=0.0)

    # ---- N > 1: rank 0 solves the problem alone first (objective the partitioned solve must reproduce)
    single_obj = None
    tte = None
    if rank == 0 and not args.no_time_to_eps:
        t = time.perf_counter()
        # eps_infeas: with the default 1e-7 the REFERENCE itself stops at iteration 0 with a false
        # "unbounded" certificate on the full-size instance; 1e-12 in both arms.
        s2 = scsb.SCS(data, cone, verbose=False, max_iters=5000, eps_infeas=1e-12)
        r2 = s2.solve(warm_start=False)
        wall = time.perf_counter() - t
        i2 = r2["info"]
        tte = dict(eps=1e-4, eps_infeas=1e-12, n_gpus=1, status=i2["status"], iters=i2["iter"], setup_ms=i2["setup_time"],
                   solve_ms=i2["solve_time"], wall_s_incl_upload=wall, pobj=i2["pobj"], dobj=i2["dobj"],
                   res_pri=i2["res_pri"], res_dual=i2["res_dual"], gap=i2["gap"])
        single_obj = (i2["pobj"], i2["dobj"])
        s2._solver.finish()
        del s2, r2
    if world > 1:
        scsb.dist_init(rank, world)

    # ---- device-resident timed region == the long call of the e2e pair
    t_setup = time.perf_counter()
    solver = scsb.SCS(data, cone, max_iters=(W + K) * ips, **zero_tol)
    setup_s = time.perf_counter() - t_setup
    inner = solver._solver
    inner.set_marks(W * ips, (W + K) * ips)
    sampler = ClockSampler(local)
    barrier_sync(td, local)
    sampler.start()
    st0 = inner.stats()
    t = time.perf_counter()
    solver.update(data["b"], data["c"])
    sol = solver.solve(warm_start=False)
    t_full = time.perf_counter() - t
    barrier_sync(td, local)
    clocks = sampler.stop()
    st1 = inner.stats()
    mk = inner.get_marks()
    assert mk is not None and mk["iters"] == K * ips, "timed region did not cover exactly K steps: %s" % (mk,)
    assert int(sol["info"]["iter"]) == (W + K) * ips
    # the same two kernels, back-to-back launches bracketed by CUDA events on the solve stream
    iso_a_ms, _ = inner.bench_spmv(0, 20)
    iso_g_ms, _ = inner.bench_spmv(1, 20)
    ms_max = max_over_ranks(td, local, mk["ms"])
    value = K * ips / (ms_max * 1e-3) if world > 1 else mk["iters"] / (ms_max * 1e-3)
    launches = int(sum_over_ranks(td, local, mk["kernel_launches"]))
    eng = inner.stats()
    t_full_max = max_over_ranks(td, local, t_full)

    # ---- the short call of the e2e pair: same problem, stops where the timed window starts
    short = scsb.SCS(data, cone, max_iters=max(1, W * ips), **zero_tol)
    barrier_sync(td, local)
    t = time.perf_counter()
    short.update(data["b"], data["c"])
    so = short.solve(warm_start=False)
    t_short = time.perf_counter() - t
    barrier_sync(td, local)
    it_short = int(so["info"]["iter"]) if W > 0 else 0
    if W == 0:
        t_short = 0.0
    short._solver.finish()
    del short
    t_short_max = max_over_ranks(td, local, t_short)
    e2e_iters = (W + K) * ips - it_short
    e2e_value = e2e_iters / max(1e-9, t_full_max - t_short_max)
    h2d_step = (st1["h2d_bytes"] - st0["h2d_bytes"]) / max(1, K)
    d2h_step = (st1["d2h_bytes"] - st0["d2h_bytes"]) / max(1, K)

    # ---- this arm on the window the reference arm can afford (driver defaults: iterations [1+W, 1+W+K))
    ref_window = None
    if world == 1:
        lo, hi = 1 + W, 1 + W + K
        s3 = scsb.SCS(data, cone, max_iters=hi, **zero_tol)
        s3._solver.set_marks(lo, hi)
        s3.solve(warm_start=False)
        m3 = s3._solver.get_marks()
        ref_window = dict(window=[lo, hi], value=m3["iters"] / (m3["ms"] * 1e-3),
                          cg_iters_per_admm_iter=m3["cg_iters"] / max(1, m3["iters"]))
        s3._solver.finish()
        del s3

    out = None
    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        g_ms = mk["spmv_g_ms"] / max(1, mk["spmv_g_launches"])
        a_ms = mk["spmv_a_ms"] / max(1, mk["spmv_a_launches"])
        ach = mk["bytes_g"] / (g_ms * 1e-3) / 1e9 if g_ms > 0 else 0.0
        tiled = bool(eng.get("tiled_g"))
        if world > 1:
            k_g = (("tiled_kernel<0> + tiled_epilogue_kernel<EpiStoreT>" if tiled else "row_kernel<ElemMul,ElemMul,EpiStoreT>")
                   + " (A_g' z_g of this rank's rows; P p + R_x p and p'Gp are a separate pass, the shared block is then "
                     "all-reduced)")
        else:
            k_g = (("tiled_kernel<0> + tiled_epilogue_chunk_kernel<EpiG> (2-D tiled: x-slices by TMA into shared memory, "
                    "accumulators in shared memory; the launches are timed as one product)") if tiled
                   else "row_kernel<ElemMul,ElemMul,EpiG,DUAL>") + " (Gp = A' z + P p + R_x p, p'Gp fused)"
        k_a = ("tiled_kernel<0> + tiled_epilogue_kernel<EpiScaleRy>" if eng.get("tiled_a")
               else "row_kernel<ElemMul,ElemMul,EpiScaleRy>")
        nnz_g = desc["nnz_A"] + desc["nnz_P"]
        roofline = dict(bound="hbm", achieved=ach, peak=peak, unit="GB/s", frac=ach / peak,
                        traffic=ncu_traffic() if world == 1 else None, kernel=k_g,
                        engine=dict(tiled_a=int(eng.get("tiled_a", 0)), tiled_g=int(eng.get("tiled_g", 0)),
                                    stored_slots_per_nnz=(eng["tiled_slots"] / eng["tiled_nnz"] if eng.get("tiled_nnz") else None)),
                        avg_launch_ms=g_ms, launches_timed=int(mk["spmv_g_launches"]),
                        timing="device %globaltimer, first CTA start -> last CTA end of the product's last kernel, "
                               "every product inside the timed region (graph WHILE-body launches cannot carry CUDA "
                               "events)" + ("; rank 0's local block" if world > 1 else ""),
                        isolated_event_ms=iso_g_ms,
                        isolated_event_achieved=(mk["bytes_g"] / (iso_g_ms * 1e-3) / 1e9 if iso_g_ms > 0 else 0.0),
                        algorithmic_bytes_per_launch=mk["bytes_g"], peak_source=peak_src,
                        second_kernel=dict(kernel=k_a + " (z = R_y^-1 A p)",
                                           avg_launch_ms=a_ms, launches_timed=int(mk["spmv_a_launches"]),
                                           isolated_event_ms=iso_a_ms,
                                           achieved=(mk["bytes_a"] / (a_ms * 1e-3) / 1e9 if a_ms > 0 else 0.0)),
                        iteration_model=dict(algorithmic_bytes=mk["algorithmic_bytes"],
                                             achieved=mk["algorithmic_bytes"] / (mk["ms"] * 1e-3) / 1e9,
                                             frac=mk["algorithmic_bytes"] / (mk["ms"] * 1e-3) / 1e9 / peak,
                                             spmv_share_of_step=(mk["spmv_g_ms"] + mk["spmv_a_ms"]) / mk["ms"],
                                             note=("rank 0's local byte model" if world > 1 else "SURVEY.md 8d byte model")))
        if world == 1 and not tiled:
            roofline["gather_ceiling_gelem_s"] = GATHER_CEILING_GELEMS
            roofline["gather_ceiling_frac"] = (nnz_g / (g_ms * 1e-3) / 1e9 / GATHER_CEILING_GELEMS if g_ms > 0 else 0.0)
        out = dict(metric="admm_iters_per_sec", value=value, unit="iters/s", n_gpus=world, steps=K, warmup=W,
                   ms_per_step=ms_max / K, higher_is_better=True, scaling="weak" if world == 1 else "strong",
                   vs_baseline=None, dtype="f64", data="synthetic", config=desc, window=[W * ips, (W + K) * ips],
                   clocks=clocks,
                   e2e=dict(value=e2e_value, unit="iters/s", h2d_bytes_per_step=h2d_step, d2h_bytes_per_step=d2h_step,
                            window=[it_short, (W + K) * ips],
                            note="wall clock, host numpy buffers: T(update(b,c) + solve(max_iters=%d)) - T(update(b,c) + "
                                 "solve(max_iters=%d)) = %.3f s - %.3f s; both calls upload b, c and download x, y, s"
                                 % ((W + K) * ips, W * ips, t_full_max, t_short_max),
                            whole_call=dict(value=(W + K) * ips / t_full_max, wall_s=t_full_max,
                                            note="the long call alone: upload, g solve, all %d iterations, download"
                                                 % ((W + K) * ips))),
                   gpu_launches=launches, roofline=roofline,
                   cg_iters_per_admm_iter=mk["cg_iters"] / max(1, mk["iters"]), setup_s=setup_s)
        if ref_window is not None:
            out["value_on_reference_window"] = ref_window
        if world > 1:
            dst = st1
            out["collectives"] = dict(per_rank_calls_in_long_call=int(dst["collectives"] - st0["collectives"]),
                                      per_rank_bytes_in_long_call=int(dst["collective_bytes"] - st0["collective_bytes"]),
                                      note="collectives issued by rank 0 during update + solve of the long call: one sum of "
                                           "(shared block + p'Gp) and one scalar gather per CG iteration, scalar gathers per "
                                           "ADMM iteration, Anderson-acceleration trapezoids every 10th.  The two of the CG "
                                           "loop are kernels over CUDA-IPC-mapped peer memory (NVLink) when the ranks can map "
                                           "each other (SCS_B200_DIST_P2P=0: NCCL), the rest are ncclAllReduce")
    solver._solver.finish()
    del solver

    # ---- time to eps = 1e-4 (the second half of BASELINE.json's metric)
    if not args.no_time_to_eps:
        if world == 1:
            if rank == 0:
                out["time_to_eps"] = tte
        else:
            barrier_sync(td, local)
            t = time.perf_counter()
            s2 = scsb.SCS(data, cone, verbose=False, max_iters=5000, eps_infeas=1e-12)
            r2 = s2.solve(warm_start=False)
            wall = max_over_ranks(td, local, time.perf_counter() - t)
            i2 = r2["info"]
            s2._solver.finish()
            del s2
            if rank == 0:
                rel = lambda a, b: abs(a - b) / max(1.0, abs(b))
                agree = dict(pobj_rel=rel(i2["pobj"], single_obj[0]), dobj_rel=rel(i2["dobj"], single_obj[1]))
                out["time_to_eps"] = dict(eps=1e-4, eps_infeas=1e-12, n_gpus=world, status=i2["status"], iters=i2["iter"],
                                          setup_ms=i2["setup_time"], solve_ms=i2["solve_time"], wall_s_incl_upload=wall,
                                          pobj=i2["pobj"], dobj=i2["dobj"], res_pri=i2["res_pri"], res_dual=i2["res_dual"],
                                          gap=i2["gap"], single_gpu=tte, agreement_with_single_gpu=agree)
                # two eps = 1e-4 solves on different summation orders agree to the stopping tolerance
                assert i2["status_val"] == 1 and agree["pobj_rel"] < 1e-3 and agree["dobj_rel"] < 1e-3, agree
    if world > 1:
        scsb.dist_finalize()

    # ---- BASELINE.json configs[4]: this GPU's share of the 8192 independent MPC QPs, no communication
    if not args.no_batch:
        try:
            b5 = batch_cfg5(scsb, rank)
            wall5 = max_over_ranks(td, local, b5["wall_s"])
            it5 = sum_over_ranks(td, local, b5["admm_iters"])
            ok5 = sum_over_ranks(td, local, b5["solved"])
            if rank == 0:
                out["cfg5_batch_mpc"] = dict(problems=1024 * world, solved=int(ok5), wall_s=wall5,
                                             problems_per_s=1024 * world / wall5, admm_iters_per_s=it5 / wall5,
                                             kernel_ms_rank0=b5["kernel_ms"], scaling="weak (1024 problems per GPU, no collective)")
        except Exception as e:  # a secondary key must never take the headline line down
            if rank == 0:
                out["cfg5_batch_mpc"] = dict(error=repr(e))
    if rank == 0 and world == 1 and not args.no_other_configs:
        out["other_configs"] = other_configs(scsb)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            out["cpu_baseline"] = cpu_baseline_sample(ips)
        except Exception as e:  # the checker must never take the product line down
            out["cpu_baseline"] = dict(value=None, unit="iters/s", cores=0, kind="reference", sample="failed: %r" % (e,))
    if td is not None:
        td.barrier()
        td.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
