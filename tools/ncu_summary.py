#!/usr/bin/env python3
"""Print the key raw metrics and the stall mix of the kernels in an .ncu-rep (development aid).
usage: ncu_summary.py report.ncu-rep [all | index]   (default: the first kernel)"""
import csv
import subprocess
import sys

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
which = sys.argv[2] if len(sys.argv) > 2 else "0"
picked = range(2, len(rows)) if which == "all" else [2 + int(which)]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.per_cycle_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.max",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers"]
for idx in picked:
    val = rows[idx]
    if len(val) < len(hdr):
        continue
    print(val[hdr.index("Kernel Name")][:100])
    for h, u, v in zip(hdr, units, val):
        if h in want:
            print(f"{h:70s} {u:16s} {v}")
    st = []
    for h, u, v in zip(hdr, units, val):
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            try:
                st.append((float(v.replace(",", "")), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    tot = sum(x for x, _ in st)
    print("stall cycles per issued instruction: " + ", ".join(f"{n} {x:.2f}" for x, n in sorted(st, reverse=True) if x > 0.04) + f"  (total {tot:.2f})")
    print()
