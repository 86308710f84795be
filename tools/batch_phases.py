#!/usr/bin/env python
"""Phase clocks of the batch engine on one GPU's share of BASELINE.json configs[4] (1024 MPC QPs)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scs_python_b200 as scsb
from scs_python_b200 import problems as bp, _scs_b200 as B

probs = [bp.mpc_qp(10_000 + i)[:2] for i in range(1024)]
for kw in (dict(), dict(acceleration_lookback=0)):
    scsb.solve_batch(probs[:64], verbose=False, **kw)
    t = time.perf_counter()
    sols = scsb.solve_batch(probs, verbose=False, **kw)
    wall = time.perf_counter() - t
    st = B.batch_stats()
    its = sum(s["info"]["iter"] for s in sols)
    tot = st["clk_total"]
    print(json.dumps(dict(settings=kw, wall_s=wall, kernel_ms=st["kernel_ms"], iters=its, us_per_iter_per_cta=st["kernel_ms"] * 1e3 * st["ctas"] / its,
                          cycles_per_iter=tot / its,
                          share=dict(equil=st["clk_equil"] / tot, factor=st["clk_factor"] / tot, linsys=st["clk_linsys"] / tot,
                                     aa=st["clk_aa"] / tot, resid=st["clk_resid"] / tot),
                          ctas=st["ctas"], smem=st["smem_per_cta"], direct=st["direct"])), flush=True)
