#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full) as a small text table for profiles/.
usage: ncu_summary.py report.ncu-rep > profiles/xxx.txt"""
import csv, io, subprocess, sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, units = rows[0], rows[1]
want = [
    ("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"),
    ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex_%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"),
    ("launch__registers_per_thread", "regs"), ("thread_inst_executed", "thread_inst"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_notsel"),
]
idx = [(h.index(n), lab) for n, lab in want if n in h]
print(f"# {rep}: ncu --set full --clock-control none (per-launch values; cold-cache, serialised)")
for r in rows[2:]:
    parts = []
    for i, lab in idx:
        v = r[i]
        if lab == "kernel":
            v = v.split("(")[0][-40:]
        else:
            try:
                v = f"{float(v.replace(',', '')):.4g}{units[i] if lab in ('time','dram_rd','dram_wr') else ''}"
            except ValueError:
                pass
        parts.append(f"{lab}={v}")
    print("  ".join(parts))
