#!/usr/bin/env python
"""Condense an `ncu -i X.ncu-rep --page raw --csv` dump into the handful of metrics DESIGN.md quotes, one row per kernel
launch.  usage: python tools/ncu_summary.py raw.csv [raw2.csv ...] > profiles/<name>_summary.csv"""
import csv
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
    "launch__block_size", "launch__grid_size", "smsp__inst_executed.sum", "sm__inst_executed.sum.per_cycle_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def main():
    w = csv.writer(sys.stdout)
    first = True
    for path in sys.argv[1:]:
        rows = list(csv.reader(open(path, newline="")))
        hdr, units, body = rows[0], rows[1], rows[2:]
        cols = [(k, hdr.index(k)) for k in KEEP if k in hdr]
        if first:
            w.writerow(["source", "kernel"] + [f"{k} [{units[i]}]" for k, i in cols])
            first = False
        kn = hdr.index("Kernel Name")
        for r in body:
            name = r[kn].split("(")[0]
            w.writerow([path.split("/")[-1], name] + [r[i] for _, i in cols])


if __name__ == "__main__":
    main()
