#!/usr/bin/env python3
"""Headline metrics of every kernel instance in an .ncu-rep as a markdown table: tools/ncu_table.py <rep> <out.md> [title]"""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], open(sys.argv[2], "w")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u = rows[0], rows[1]
K = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
     "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
     "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
     "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
     "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
print("# %s — %s (ncu --set full --clock-control none)\n" % (rep.split("/")[-1], sys.argv[3] if len(sys.argv) > 3 else ""), file=out)
for v in rows[2:]:
    print("## `%s`\n\n| metric | value | unit |\n|---|---|---|" % v[h.index("Kernel Name")], file=out)
    for k in K:
        if k in h:
            print("| %s | %s | %s |" % (k, v[h.index(k)], u[h.index(k)]), file=out)
    print("", file=out)
