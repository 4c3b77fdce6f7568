#!/usr/bin/env python
"""One CSV row per profiled launch from an `ncu --set full` report: the counters DESIGN.md and bench.py quote.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv && python tools/ncu_summary.py raw.csv > profiles/<name>.csv
"""
import csv
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "smsp__inst_executed.sum",
           "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
           "lts__t_sector_hit_rate.pct", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [m for m in METRICS if m in idx]
    w = csv.writer(sys.stdout)
    w.writerow(["Kernel Name"] + cols)
    w.writerow([""] + [units[idx[m]] for m in cols])
    for r in rows[2:]:
        if len(r) == len(hdr):
            w.writerow([r[idx["Kernel Name"]]] + [r[idx[m]] for m in cols])


if __name__ == "__main__":
    main()
