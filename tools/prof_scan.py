#!/usr/bin/env python3
"""Driver for the kmsg-scan measurements (BASELINE configs[2]: 100 MB synthetic dmesg buffer) and its ncu captures."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import gpud_b200 as g
import synth

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
cpu = len(sys.argv) > 2 and sys.argv[2] == "cpu"
unit = synth.dmesg_buffer(4 << 20, hit_every=1000)
buf = unit * 25                                             # 100 MiB of lines; every tile starts on a line boundary
ctx = g.Context([0])
d = torch.frombuffer(bytearray(buf), dtype=torch.uint8).cuda()
torch.cuda.synchronize()
tot = []
for _ in range(reps):                                      # the product's default: one overlapped chain of launches, whole device time only
    hits, n_units = ctx.kmsg_scan_device(d.data_ptr(), len(buf), cap=1 << 20)
    tot.append(ctx.scan_kernel_ms()[0])
total_ms = float(np.mean(tot[1:] if reps > 1 else tot))
ctx.scan_phase_timing(True)                                # plain launches split by events: the per-phase numbers (their sum is larger)
ms = []
for _ in range(reps):
    hits, n_units = ctx.kmsg_scan_device(d.data_ptr(), len(buf), cap=1 << 20)
    ms.append(ctx.scan_kernel_ms())
ms = np.array(ms[1:] if reps > 1 else ms)
filt, pre, mat = ms.mean(axis=0)
ctx.scan_phase_timing(False)
harr = (g.XidHit * (1 << 17))()
ctx.kmsg_scan_c(buf, harr, 1 << 17)                          # warm-up (allocations, pinned staging)
t0 = time.perf_counter()
nh2, _ = ctx.kmsg_scan_c(buf, harr, 1 << 17)                 # host bytes -> hits in host memory, through the C ABI
e2e = time.perf_counter() - t0
assert nh2 == len(hits)
out = {"bytes": len(buf), "lines": n_units, "hits": len(hits), "filter_ms": float(filt), "prefix_ms": float(pre), "match_ms": float(mat),
       "device_total_ms": total_ms, "phases_sum_ms": float(filt + pre + mat), "filter_GBps": len(buf) / filt / 1e6, "total_GBps": len(buf) / total_ms / 1e6,
       "e2e_host_ms": e2e * 1e3, "e2e_GBps": len(buf) / e2e / 1e9, "stats": ctx.scan_stats()}
if cpu:
    from oracle import coracle
    t0 = time.perf_counter()
    ch, nl = coracle.scan_lines(buf)
    dt = time.perf_counter() - t0
    assert [(h.line, h.kind, h.code) for h in ch] == [(h.unit_index, h.kind, h.code) for h in hits]
    out.update(cpu_threads=coracle.max_threads(), cpu_ms=dt * 1e3, cpu_GBps=len(buf) / dt / 1e9, cpu_hits_equal=True)
print(json.dumps(out))
if len(sys.argv) > 2 and sys.argv[2] == "sharded":
    # the same 100 MiB buffer over every GPU of the box through gpud_kmsg_scan_sharded (host bytes in, merged hits out)
    n_gpu = torch.cuda.device_count()
    res = {}
    for n in sorted({1, 2, 4, n_gpu} & set(range(1, n_gpu + 1))):
        c = g.Context(list(range(n)))
        harr2 = (g.XidHit * (1 << 17))()
        c.kmsg_scan_sharded_c(buf, harr2, 1 << 17)             # warm-up (allocations, pinned staging)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            nh, nu = c.kmsg_scan_sharded_c(buf, harr2, 1 << 17)
            ts.append(time.perf_counter() - t0)
        assert nh == len(hits) and nu == n_units
        res[n] = {"host_ms": min(ts) * 1e3, "GBps": len(buf) / min(ts) / 1e9}
        # the same from pinned caller memory (gpud_host_alloc): no staging copy, every GPU pulls its piece over its own PCIe link
        import ctypes as C
        hp = C.c_void_p()
        assert g.lib().gpud_host_alloc(C.c_int64(len(buf)), C.byref(hp)) == 0
        C.memmove(hp, buf, len(buf))
        L = g.lib()
        nh, nu = C.c_int64(), C.c_int64()
        tp = []
        for _ in range(4):
            t0 = time.perf_counter()
            rc = L.gpud_kmsg_scan_sharded(c._h, hp, C.c_int64(len(buf)), 0, harr2, C.c_int64(1 << 17), C.byref(nh), C.byref(nu))
            tp.append(time.perf_counter() - t0)
        assert rc == 0 and nh.value == len(hits)
        res[n]["pinned_host_ms"] = min(tp[1:]) * 1e3
        res[n]["pinned_GBps"] = len(buf) / min(tp[1:]) / 1e9
        L.gpud_host_free(hp)
        c.close()
    print(json.dumps({"sharded_scan_100MiB_host_bytes": res}))
