#!/usr/bin/env python
"""Headline counters of one kernel from an .ncu-rep (ncu -i X --page raw --csv | ncu_kernel_summary.py)."""
import csv
import sys

rows = list(csv.reader(sys.stdin))
h = rows[0]
WANT = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for v in rows[2:]:
    name = v[h.index("Kernel Name")] if "Kernel Name" in h else "?"
    print("==", name[:60])
    for i, n in enumerate(h):
        if n in WANT:
            print("  %-80s %s %s" % (n, v[i], rows[1][i]))
    stalls = sorted(((float(v[i]), n) for i, n in enumerate(h) if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio") and v[i]), reverse=True)
    print("  stalls per issue:", ", ".join("%s %.2f" % (n.split("stalled_")[1].split("_per_")[0], x) for x, n in stalls[:7]))
