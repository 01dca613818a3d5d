#!/usr/bin/env python3
"""Scan end-to-end timing from host memory: pageable bytes vs pinned (gpud_host_alloc) input, and a plain memcpy for scale."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import gpud_b200 as g
import synth

ctx = g.Context([0])
L = ctx._L
buf = synth.dmesg_buffer(4 << 20, hit_every=1000) * 25
n = len(buf)
harr = (g.XidHit * (1 << 17))()
nh, nu = C.c_int64(), C.c_int64()


def scan(ptr):
    rc = L.gpud_kmsg_scan(ctx._h, 0, C.c_void_p(ptr), n, 0, harr, 1 << 17, C.byref(nh), C.byref(nu))
    assert rc == 0, rc


src = np.frombuffer(buf, dtype=np.uint8)
for name, ptr in (("pageable", src.ctypes.data),):
    scan(ptr)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); scan(ptr); ts.append((time.perf_counter() - t0) * 1e3)
    print(name, "ms:", [round(t, 2) for t in ts], "hits", nh.value)
p = C.c_void_p()
assert L.gpud_host_alloc(n, C.byref(p)) == 0
C.memmove(p, src.ctypes.data, n)
scan(p.value)
ts = []
for _ in range(5):
    t0 = time.perf_counter(); scan(p.value); ts.append((time.perf_counter() - t0) * 1e3)
print("pinned ms:", [round(t, 2) for t in ts], "hits", nh.value)
dst = np.empty_like(src)
ts = []
for _ in range(5):
    t0 = time.perf_counter(); C.memmove(dst.ctypes.data, src.ctypes.data, n); ts.append((time.perf_counter() - t0) * 1e3)
print("plain memcpy ms:", [round(t, 2) for t in ts])
print("kernel ms:", ctx.scan_kernel_ms(dev=0))
