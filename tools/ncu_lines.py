#!/usr/bin/env python3
"""Per-CUDA-source-line executed warp-instructions per unit from an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv, io, subprocess, sys
rep, units = sys.argv[1], float(sys.argv[2])
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 3.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur = "?"
tot = 0.0
for r in csv.reader(io.StringIO(txt)):
    if len(r) < 8:
        if r and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        continue
    if r[0] in ("Line No", ""):
        continue
    try:
        c = int(r[7]) / units
    except ValueError:
        continue
    tot += c
    if c >= thr:
        print("%7.1f %6s %s:%s  %s" % (c, r[4], cur, r[0], r[1].strip()[:105]))
print("total %.1f" % tot)
