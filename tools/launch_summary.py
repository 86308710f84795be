#!/usr/bin/env python
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel.

    python tools/launch_summary.py gpurun_out/launches.csv > profiles/<name>.txt
"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    agg = OrderedDict()
    tot = 0.0
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        name = re.sub(r"\(.*", "", name)
        t = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"):
            t *= 1e3
        elif r["Metric Unit"] in ("ms", "msecond"):
            t *= 1e6
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
        tot += t
    print("# %s : %d launches, %.3f ms total (ncu per-launch times are cold-cache and serialised: compare shares)" %
          (path, sum(a[0] for a in agg.values()), tot / 1e6))
    print("%-110s %8s %12s %10s %7s" % ("kernel", "launches", "total_us", "avg_us", "share"))
    for name, (cnt, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-110s %8d %12.1f %10.2f %6.1f%%" % (name[:110], cnt, t / 1e3, t / 1e3 / cnt, 100.0 * t / tot))


if __name__ == "__main__":
    main()
