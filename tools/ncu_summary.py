#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full capture) into a small text table for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt
"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    print("# %s" % rep)
    for n, r in enumerate(rows[2:]):
        print("\n[%d] %s" % (n, r[kn][:160]))
        rd = wr = None
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print("  %-80s %12s %s" % (w, r[i], units[i]))
                if w == "dram__bytes_read.sum":
                    rd = (float(r[i].replace(",", "")), units[i])
                if w == "dram__bytes_write.sum":
                    wr = (float(r[i].replace(",", "")), units[i])
        if rd and wr:
            sc = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            print("  %-80s %12.0f byte" % ("traffic = dram read + write", rd[0] * sc[rd[1]] + wr[0] * sc[wr[1]]))


if __name__ == "__main__":
    main()
