#!/usr/bin/env python
"""Markdown table from the JSON lines of tools/bench_configs.py (time to eps = 1e-4 per config and arm)."""
import json
import sys

rows = [json.loads(l) for f in sys.argv[1:] for l in open(f) if l.strip().startswith("{")]
cfgs, arms = [], []
for r in rows:
    if r["config"] not in cfgs:
        cfgs.append(r["config"])
    a = r["arm"] + ("" if "build" not in r else " (%s, %d core%s)" % (r["build"].split("/")[-1], r.get("cores", 1), "s" if r.get("cores", 1) > 1 else ""))
    r["_arm"] = a
    if a not in arms:
        arms.append(a)
print("| config | " + " | ".join(arms) + " |")
print("|---|" + "---|" * len(arms))
for c in cfgs:
    cells = []
    for a in arms:
        m = [r for r in rows if r["config"] == c and r["_arm"] == a]
        if not m:
            cells.append("—")
            continue
        r = m[-1]
        if "error" in r or "unavailable" in r:
            cells.append("n/a: " + str(r.get("error", r.get("unavailable")))[:40])
            continue
        t = (r.get("setup_ms", 0.0) + r.get("solve_ms", 0.0)) / 1e3
        cell = "%.3g s (%s it, %s)" % (t, r.get("iters"), str(r.get("status"))[:28])
        if r.get("note"):
            cell += " **" + r["note"] + "**"
        if r.get("sample"):
            cell += " [extrapolated]"
        cells.append(cell)
    print("| " + c + " | " + " | ".join(cells) + " |")
