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
# data shape: argv[2] in {uniform (default), mw, temp, const, walk, mixed}: how tie-heavy the gauges are
shape = sys.argv[2] if len(sys.argv) > 2 else "uniform"
walk = torch.zeros((1, F), dtype=torch.float64, device="cuda")
for i in range(CAP // (1 << 16)):
    n = 1 << 16
    if shape == "mw":          # integer readings, few ties (power in mW)
        t = torch.randint(30000, 90000, (n, F), dtype=torch.int32, device="cuda", generator=gen).to(torch.float64)
    elif shape == "temp":      # integer degrees: ~16 copies of every value per window
        t = torch.randint(30, 90, (n, F), dtype=torch.int32, device="cuda", generator=gen).to(torch.float64)
    elif shape == "const":     # a flat gauge (utilisation pinned at 100 %)
        t = torch.full((n, F), 100.0, dtype=torch.float64, device="cuda")
    elif shape == "walk":      # slow random walk in integer steps: long runs of equal values, trends
        steps = torch.randint(-1, 2, (n, F), dtype=torch.int32, device="cuda", generator=gen).to(torch.float64)
        t = walk + torch.cumsum(steps, 0) * (torch.rand((n, F), device="cuda", generator=gen) < 0.05)
        t = walk + torch.cumsum(steps * (torch.rand((n, F), device="cuda", generator=gen) < 0.05), 0)
        walk = t[-1:].clone()
        t = t + 60.0
    elif shape == "mixed":     # a quarter of the fields each: mw / temp / const / uniform
        t = torch.rand((n, F), dtype=torch.float64, device="cuda", generator=gen) * 60.0 + 30.0
        q = F // 4
        t[:, :q] = torch.randint(30000, 90000, (n, q), dtype=torch.int32, device="cuda", generator=gen).to(torch.float64)
        t[:, q:2 * q] = torch.randint(30, 90, (n, q), dtype=torch.int32, device="cuda", generator=gen).to(torch.float64)
        t[:, 2 * q:3 * q] = 100.0
    else:
        t = torch.rand((n, F), dtype=torch.float64, device="cuda", generator=gen) * 60.0 + 30.0
    ring.push_device(t.data_ptr(), 1 << 16)
    ring.sync()
for _ in range(reps):
    ring.reduce()
ring.sync()
print(shape, "kernel ms (reduce, carry):", ring.kernel_ms())
if len(sys.argv) > 3 and sys.argv[3] == "range":
    import time
    ring.reduce_range(0)
    ring.sync()
    ts, st = [], []
    for _ in range(5):
        t0 = time.perf_counter()
        out = ring.reduce_range(0)
        ts.append(time.perf_counter() - t0)
        st.append(ring.range_stats())
    print("reduce_range(whole ring) host ms:", [round(t * 1e3, 3) for t in ts], "device (pass ms, total ms, fields redone):", st, "p99[0..3]", out["p99"][:3], "open reasons:", ring.range_open_reasons)
