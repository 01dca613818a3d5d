#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here with `ncu -i`, no GPU needed): headline metrics + per-opcode executed counts.
usage: tools/ncu_summary.py <rep> <units-per-launch> [out.md]"""
import collections
import csv
import io
import subprocess
import sys

rep, units = sys.argv[1], float(sys.argv[2])
out = open(sys.argv[3], "w") if len(sys.argv) > 3 else sys.stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
K = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
     "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
     "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
     "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
     "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
     "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]
K += [x for x in h if x.startswith("smsp__average_warps_issue_stalled") and x.endswith("per_issue_active.ratio")]
print("# %s\n\nkernel: `%s`\n" % (rep.split("/")[-1], v[h.index("Kernel Name")] if "Kernel Name" in h else "?"), file=out)
print("| metric | value | unit |\n|---|---|---|", file=out)
for k in K:
    if k in h:
        print("| %s | %s | %s |" % (k, v[h.index(k)], u[h.index(k)]), file=out)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
isrc, iex, ist = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
hist, stall = collections.Counter(), collections.Counter()
for r in data:
    if not r[iex].isdigit():
        continue
    t = r[isrc].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    hist[op] += int(r[iex]); stall[op] += int(r[ist] or 0)
tot = sum(hist.values())
print("\nwarp-instructions executed per unit (%g units/launch): **%.1f**\n" % (units, tot / units), file=out)
print("| opcode | per unit | stall samples |\n|---|---|---|", file=out)
for op, c in hist.most_common(32):
    print("| %s | %.1f | %d |" % (op, c / units, stall[op]), file=out)

# optional 4th argument: write {dram_bytes_read, dram_bytes_write} (bytes per launch) as JSON for bench.py's roofline.traffic
if len(sys.argv) > 4:
    import json
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    def val(k):
        return float(v[h.index(k)].replace(",", "")) * scale[u[h.index(k)]]
    json.dump({"dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
               "gpu_time_us_under_ncu": float(v[h.index("gpu__time_duration.sum")].replace(",", "")),
               "kernel": v[h.index("Kernel Name")] if "Kernel Name" in h else "?", "source": rep.split("/")[-1] + " (ncu --set full --clock-control none)"},
              open(sys.argv[4], "w"), indent=1)
