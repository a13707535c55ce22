#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into the few numbers the roofline uses.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]
for w in want:
    for i, h in enumerate(hdr):
        if h == w:
            print(f"{w} [{units[i]}]: " + " | ".join(r[i][:44] for r in data))
print("-- warp stall reasons (stalled warps per issue-active cycle, > 0.2) --")
for i, h in enumerate(hdr):
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
        try:
            if max(float(r[i]) for r in data) > 0.2:
                print(h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "") + ": " + " | ".join(r[i][:8] for r in data))
        except ValueError:
            pass
