#!/usr/bin/env python
"""Print the metrics that matter from an .ncu-rep (raw page) -- used to write the summaries under profiles/."""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warps_issue_stalled", "sm__cycles_elapsed.avg ",
        "smsp__cycles_active.avg", "launch__waves_per_multiprocessor", "sm__cycles_active.avg", "lts__t_bytes.sum ", "launch__shared_mem",
        "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_lsu", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct"]
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
for r in rows[2:]:
    print("==", r[h.index("Kernel Name")][:100], r[h.index("Grid Size")], r[h.index("Block Size")])
    for i, c in enumerate(h):
        if any(k in c for k in KEYS) and r[i] not in ("", "0"):
            print(f"   {c:90s} {r[i]:>16s} {rows[1][i]}")
