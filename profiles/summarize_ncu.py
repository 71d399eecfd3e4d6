#!/usr/bin/env python
"""Turns an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the short text summary kept in profiles/."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]
STALL = "smsp__average_warps_issue_stalled_"

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("kernel:", d.get("Kernel Name"), "| grid", d.get("Grid Size"), "block", d.get("Block Size"))
    for k in KEYS:
        if k in d:
            print(f"  {k:85s} {d[k]:>18s} {units[hdr.index(k)]}")
    st = sorted(((float(d[h]), h[len(STALL):-len('_per_issue_active.ratio')]) for h in hdr
                 if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and d[h]), reverse=True)
    print("  warp stall reasons (warps per issue-active cycle):", ", ".join(f"{n} {v:.2f}" for v, n in st[:8]))
