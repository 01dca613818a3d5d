#!/usr/bin/env python3
"""Small driver for ncu captures: fills the bench-sized ring on the device (tests/synth_device.py shapes) and runs the fused
reduce a few times; `range` as third argument also runs the W = CAP range call.

  python tools/prof_ring.py [reps] [shape] [range]      shape in synth_device.SHAPES (default uniform)
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import gpud_b200 as g
import synth_device as sd

F, CAP, W = 512, 1 << 20, 1000
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
shape = sys.argv[2] if len(sys.argv) > 2 else "uniform"
fields = os.environ.get("PROF_FIELDS", "")          # e.g. "counter" / "util" / "temp" / "power": only that class of the survey mix
ctx = g.Context([0])
seed = 0x67707564
ring = g.Ring(ctx, F, CAP, W, thresholds=sd.thresholds(shape, F, CAP, seed))
dev = torch.device("cuda", 0)
if shape == "survey" and fields:
    p = sd.survey_params(F, CAP, seed)
    sel = {"counter": p["counter"], "temp": (p["kind"] == 0) & ~p["counter"], "power": (p["kind"] == 1) & ~p["counter"],
           "util": (p["kind"] == 2) & ~p["counter"]}[fields]
    idx = torch.tensor(np.flatnonzero(sel), device=dev)
    idx = idx[torch.arange(F, device=dev) % idx.numel()]                 # every column becomes a field of the chosen class
    for x, n in sd.chunks(shape, F, CAP, 1 << 16, seed, dev):
        x = x[:, idx].contiguous()
        torch.cuda.synchronize()
        ring.push_device(x.data_ptr(), n)
        ring.sync()
else:
    sd.fill_ring(ring, shape, F, CAP, seed, dev)
for _ in range(reps):
    ring.reduce()
ring.sync()
print(shape, fields, "kernel ms (reduce, carry):", ring.kernel_ms())
if len(sys.argv) > 3 and sys.argv[3] == "range":
    ring.reduce_range(0)
    ring.sync()
    ts, st = [], []
    for _ in range(5):
        t0 = time.perf_counter()
        out = ring.reduce_range(0)
        ts.append(time.perf_counter() - t0)
        st.append(ring.range_stats())
    print("reduce_range(whole ring) host ms:", [round(t * 1e3, 3) for t in ts], "device (pass ms, total ms, fields redone):", st, "p99[0..3]", out["p99"][:3],
          "open reasons:", ring.range_open_reasons)
