#!/usr/bin/env python
"""Time-to-eps = 1e-4 and ADMM iterations/s on every BASELINE.json config, this backend next to the
reference's CPU backends on the same box (QDLDL direct and CPU_INDIRECT from oracle/_ref, all host
cores for the OpenMP build).  One JSON line per (config, arm); not the driver's bench line
(bench.py is) -- these are the "every config" numbers of the north star.

    python tools/bench_configs.py [--configs 1,3,4,5,2s] [--ref-time-limit 120] > gpurun_out/configs.jsonl

Config 2 at full size is bench.py's workload; "2s" here is the 1/16-scale instance so that the CPU
arms finish.  Reference runs are capped with time_limit_secs (reported as "did not finish").
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build(cfg):
    from scs_python_b200 import problems as P
    if cfg == "1":
        d, K, aux = P.random_cone_qp(seed=1234, with_P=True)
        return "cfg1 random cone QP n=2000 m=6000 (l+q+ep), P=0.1I", [(d, K)], {}
    if cfg == "1lp":
        d, K, aux = P.random_cone_qp(seed=1234, with_P=False)
        return "cfg1 random cone LP n=2000 m=6000 (l+q+ep), no P", [(d, K)], {}
    if cfg == "2s":
        d, K, aux = P.lasso(1_000_000 // 16, 2_000_000 // 16, 100, seed=0)
        return "cfg2 sparse LASSO at 1/16 scale (n=m=250k, nnz 6.6M)", [(d, K)], dict(eps_infeas=1e-12)
    if cfg == "3":
        d, K, aux = P.socp_portfolio(seed=0)
        return "cfg3 SOCP portfolio n=50k, 10k second-order cones", [(d, K)], {}
    if cfg == "4":
        d, K, aux = P.maxcut_sdp(seed=0)
        return "cfg4 MaxCut SDP 64 x PSD(200)", [(d, K)], {}
    if cfg == "4s":
        d, K, aux = P.maxcut_sdp(seed=0, nodes=200, blocks=8)
        return "cfg4 MaxCut SDP 8 x PSD(200) (1/8 scale)", [(d, K)], {}
    if cfg == "5":
        return "cfg5 batch of 8192 MPC QPs n=120 m=360", [P.mpc_qp(s)[:2] for s in range(8192)], {}
    if cfg == "5s":
        return "cfg5 batch of 1024 MPC QPs n=120 m=360 (one GPU's share of 8)", [P.mpc_qp(s)[:2] for s in range(1024)], {}
    raise SystemExit("unknown config " + cfg)


def run_b200(name, probs, kw):
    import scs_python_b200 as scsb
    from scs_python_b200 import _scs_b200 as B
    out = dict(arm="b200", config=name, problems=len(probs))
    if len(probs) == 1:
        d, K = probs[0]
        best = None
        for rep in range(2):   # second run: clocks / allocator warm
            t = time.perf_counter()
            s = scsb.SCS(d, K, verbose=False, **kw)
            r = s.solve(warm_start=False)
            wall = time.perf_counter() - t
            i = r["info"]
            rec = dict(status=i["status"], iters=i["iter"], setup_ms=i["setup_time"], solve_ms=i["solve_time"], wall_s=wall,
                       iters_per_s=i["iter"] / max(i["solve_time"], 1e-9) * 1e3, pobj=i["pobj"], dobj=i["dobj"],
                       res_pri=i["res_pri"], res_dual=i["res_dual"], gap=i["gap"],
                       lin_sys_ms=i["lin_sys_time"], cone_ms=i["cone_time"], accel_ms=i["accel_time"],
                       scale_updates=i["scale_updates"])
            try:
                st = s._solver.stats()
                rec["cg_iters_per_admm_iter"] = st.get("cg_iters", 0) / max(1, i["iter"])
            except Exception:
                pass
            del s
            if best is None or rec["solve_ms"] + rec["setup_ms"] < best["solve_ms"] + best["setup_ms"]:
                best = rec
        out.update(best)
    else:
        prepared = [scsb._prepare(d, k) for d, k in probs]
        best = None
        for rep in range(2):
            t = time.perf_counter()
            sols = B.solve_batch(prepared, verbose=False, **kw)
            wall = time.perf_counter() - t
            st = B.batch_stats()
            its = np.array([s["info"]["iter"] for s in sols])
            rec = dict(status="solved %d / %d" % (sum(s["info"]["status_val"] == 1 for s in sols), len(sols)),
                       iters=int(its.sum()), iters_mean=float(its.mean()), iters_max=int(its.max()), wall_s=wall,
                       kernel_ms=st["kernel_ms"], pack_ms=st["pack_ms"], solve_ms=wall * 1e3, setup_ms=0.0,
                       iters_per_s=float(its.sum()) / wall, iters_per_s_kernel=float(its.sum()) / st["kernel_ms"] * 1e3,
                       problems_per_s=len(sols) / wall, fused=st["fused"], streamed=st["streamed"], direct=st["direct"],
                       h2d_bytes=st["h2d_bytes"], d2h_bytes=st["d2h_bytes"], ctas=st["ctas"], smem_per_cta=st["smem_per_cta"])
            if best is None or rec["wall_s"] < best["wall_s"]:
                best = rec
        out.update(best)
    return out


def run_ref(name, probs, kw, solver_name, limit_s, max_problems, build_dir):
    """one reference backend from one build of oracle/_ref ("" = single-threaded core, "scs_omp" =
    OpenMP core); runs in its own process (both builds are a package called `scs`)"""
    p = os.path.join(ROOT, "oracle", "_ref", build_dir) if build_dir else os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(p, "scs", "__init__.py")):
        return dict(arm="reference:" + solver_name, config=name, unavailable="oracle/_ref/%s not on this box" % build_dir)
    sys.path.insert(0, p)
    cores = (os.cpu_count() or 1) if build_dir == "scs_omp" else 1
    os.environ["OMP_NUM_THREADS"] = str(cores)
    os.environ["OPENBLAS_NUM_THREADS"] = str(cores)
    import scs
    ls = dict(qdldl=scs.LinearSolver.QDLDL, cpu_indirect=scs.LinearSolver.CPU_INDIRECT,
              gpu_indirect=scs.LinearSolver.GPU_INDIRECT)[solver_name]  # gpu_indirect: the reference's own cuSPARSE/cuBLAS backend
    out = dict(arm="reference:" + solver_name, config=name, problems=len(probs), cores=cores,
               build="oracle/_ref/" + (build_dir or "scs"))
    sample = probs[:max_problems]
    tot_it, tot_setup, tot_solve, statuses = 0, 0.0, 0.0, {}
    t0 = time.perf_counter()
    last = None
    for d, K in sample:
        s = scs.SCS(d, K, linear_solver=ls, verbose=False, time_limit_secs=float(limit_s), **kw)
        r = s.solve(warm_start=False)
        i = r["info"]
        tot_it += i["iter"]; tot_setup += i["setup_time"]; tot_solve += i["solve_time"]
        statuses[i["status"]] = statuses.get(i["status"], 0) + 1
        last = i
    wall = time.perf_counter() - t0
    scale = len(probs) / len(sample)
    out.update(status=(last["status"] if len(sample) == 1 else json.dumps(statuses)), iters=tot_it, setup_ms=tot_setup * scale,
               solve_ms=tot_solve * scale, wall_s=wall * scale, iters_per_s=tot_it / max(tot_solve, 1e-9) * 1e3,
               pobj=last["pobj"], dobj=last["dobj"], res_pri=last["res_pri"], res_dual=last["res_dual"], gap=last["gap"])
    if len(sample) != len(probs):
        out["sample"] = "%d of %d problems timed one after another on one core; solve_ms/setup_ms/wall_s extrapolated x%.0f" % (
            len(sample), len(probs), scale)
    if "time_limit" in last["status"]:
        out["note"] = "did not finish within time_limit_secs=%g" % limit_s
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,1lp,3,4,5,2s")
    ap.add_argument("--ref-time-limit", type=float, default=120.0)
    ap.add_argument("--ref-batch-sample", type=int, default=64)
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--ref-arms", default="qdldl:-,cpu_indirect:scs_omp,gpu_indirect:scs_refgpu",
                    help="solver:build pairs; builds are the package directories under oracle/_ref (- = scs, single-threaded)")
    ap.add_argument("--ref-worker", nargs=3, metavar=("CFG", "SOLVER", "BUILD"), help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.ref_worker:
        cfg, solver_name, build_dir = args.ref_worker
        name, probs, kw = build(cfg)
        kw = dict(kw, eps_abs=1e-4, eps_rel=1e-4)
        print("REF " + json.dumps(run_ref(name, probs, kw, solver_name, args.ref_time_limit, args.ref_batch_sample,
                                          "" if build_dir == "-" else build_dir)), flush=True)
        return
    for cfg in args.configs.split(","):
        name, probs, kw = build(cfg)
        kw = dict(kw, eps_abs=1e-4, eps_rel=1e-4)
        try:
            print(json.dumps(run_b200(name, probs, kw)), flush=True)
        except Exception as e:  # keep going: one config failing must not hide the others
            print(json.dumps(dict(arm="b200", config=name, error=repr(e))), flush=True)
        if args.no_ref:
            continue
        for arm in args.ref_arms.split(","):
            solver_name, build_dir = arm.split(":")
            if cfg == "2s" and solver_name == "qdldl":
                # the factorisation (scs_init, not covered by time_limit_secs) fills in to a dense 62500^2 Schur complement
                print(json.dumps(dict(arm="reference:qdldl", config=name, build="oracle/_ref/scs", cores=1,
                                      unavailable="not attempted: QDLDL fill-in on the random sparse LASSO matrix")), flush=True)
                continue
            if True:
                try:
                    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--ref-worker", cfg, solver_name, build_dir,
                                        "--ref-time-limit", str(args.ref_time_limit), "--ref-batch-sample",
                                        str(args.ref_batch_sample)], capture_output=True, text=True,
                                       timeout=args.ref_time_limit * 2 + 180)
                    lines = [ln[4:] for ln in r.stdout.splitlines() if ln.startswith("REF ")]
                    print(lines[-1] if lines else json.dumps(dict(arm="reference:" + solver_name, config=name,
                                                                  error=(r.stderr or r.stdout)[-400:])), flush=True)
                except Exception as e:
                    print(json.dumps(dict(arm="reference:" + solver_name, config=name, error=repr(e))), flush=True)


if __name__ == "__main__":
    main()
