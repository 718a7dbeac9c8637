#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page CSV) into the handful of metrics DESIGN.md / profiles/ cite.
usage: ncu -i X.ncu-rep --page raw --csv | python tools/ncu_summary.py"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
WANT = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_blocks",
    "launch__occupancy_limit_warps", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum",
    "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
]
idx = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if "warp_issue_stalled" in h and h.endswith("_per_warp_active.pct")]
for r in rows[2:]:
    print("=" * 100)
    for w in WANT:
        if w in idx:
            print(f"  {w:78s} {r[idx[w]]:>18s} {units[idx[w]]}")
    st = sorted(((float(r[idx[h]].replace(',', '') or 0), h) for h in stall), reverse=True)[:8]
    for v, h in st:
        print(f"  stall {h.split('stalled_')[1].replace('_per_warp_active.pct', ''):40s} {v:8.1f} %")
