#!/usr/bin/env python3
"""Small driver for ncu captures: fills the bench-sized ring on the device and runs the fused reduce a few times."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import gpud_b200 as g

F, CAP, W = 512, 1 << 20, 1000
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ctx = g.Context([0])
ring = g.Ring(ctx, F, CAP, W, thresholds=np.full(F, 88.0))
gen = torch.Generator(device="cuda").manual_seed(0x67707564)
for i in range(CAP // (1 << 16)):
    t = torch.rand((1 << 16, F), dtype=torch.float64, device="cuda", generator=gen) * 60.0 + 30.0
    ring.push_device(t.data_ptr(), 1 << 16)
    ring.sync()
for _ in range(reps):
    ring.reduce()
ring.sync()
print("kernel ms (reduce, carry):", ring.kernel_ms())
if len(sys.argv) > 2 and sys.argv[2] == "range":
    import time
    ring.reduce_range(0)
    ring.sync()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        out = ring.reduce_range(0)
        ts.append(time.perf_counter() - t0)
    print("reduce_range(whole ring) ms:", [round(t * 1e3, 3) for t in ts], "p99[0..3]", out["p99"][:3])
