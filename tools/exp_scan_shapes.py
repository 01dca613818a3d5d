import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import gpud_b200 as g, synth
ctx = g.Context([0])
ctx.scan_phase_timing(True)          # per-phase event timing (plain launches)
def run(name, buf):
    d = torch.frombuffer(bytearray(buf), dtype=torch.uint8).cuda(); torch.cuda.synchronize()
    ms = []
    for _ in range(5):
        hits, nu = ctx.kmsg_scan_device(d.data_ptr(), len(buf), cap=1 << 20)
        ms.append(ctx.scan_kernel_ms())
    print(name, len(buf), "hits", len(hits), "filter/prefix/match ms", np.array(ms[1:]).mean(axis=0).round(4), "stats", ctx.scan_stats())
edge = list(synth.EDGE_LINES)
run("default", synth.dmesg_buffer(4 << 20, hit_every=1000) * 25)
synth.EDGE_LINES[:] = []
run("no-edge", synth.dmesg_buffer(4 << 20, hit_every=1000) * 25)
hl = synth.hit_lines()
lens = sorted(((len(l), l[:80]) for l in hl + edge), reverse=True)[:5]
print("longest hit lines:", [(a, b) for a, b in lens])
# single-kind buffers
import types
for name, lines in (("plain-xid", [l for l in hl if "Xid (PCI" in l and "Link" not in l][:50]), ("extended", [l for l in hl if " Link " in l][:50]), ("sxid", [l for l in hl if "SXid" in l][:50])):
    orig = synth.hit_lines
    synth.hit_lines = lambda lines=lines: lines
    run(name + "(%d)" % len(lines), synth.dmesg_buffer(4 << 20, hit_every=1000) * 25)
    synth.hit_lines = orig
