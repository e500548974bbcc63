#!/usr/bin/env python
"""Markdown table of the headline counters of every launch in an .ncu-rep (`ncu --set full` capture).
usage: ncu_rep_table.py report.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
h = r[0]
M = [("gpu__time_duration.sum", "ms"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
     ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/instr"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"), ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
     ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
     ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
     ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instr"),
     ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
     ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait")]
units = r[1]
names = {"0": "FETCH", "1": "HEAD", "2": "FRONT", "3": "SEARCH", "4": "TAIL", "5": "CORR"}
print("| # | kernel | " + " | ".join(m[1] for m in M) + " |")
print("|---|---|" + "---:|" * len(M))
for row in r[2:]:
    n = row[h.index("Kernel Name")].split("(")[0]
    if "reset" in n or "clear" in n:
        continue
    if "lgar_wave_kernel<" in n:
        n = names[n.split("<")[1][0]] + " `lgar_wave_kernel<%s>`" % n.split("<")[1][0]
    else:
        n = "`%s`" % n.replace("void ", "")
    cells = []
    for m, _ in M:
        if m in h:
            v, u = row[h.index(m)], units[h.index(m)]
            try:
                f = float(v.replace(",", ""))
                cells.append(("%.2f %s" % (f, u.replace("byte", "B"))) if "byte" in u else ("%.3g" % f))
            except ValueError:
                cells.append(v)
        else:
            cells.append("-")
    print("| %s | %s | " % (row[h.index("ID")], n) + " | ".join(cells) + " |")
