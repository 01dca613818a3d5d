"""match-step time against hit density and hit kind (what is k_scan_match's duration made of?)"""
import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import gpud_b200 as g, synth
ctx = g.Context([0])
ctx.scan_phase_timing(True)
def run(name, buf):
    d = torch.frombuffer(bytearray(buf), dtype=torch.uint8).cuda(); torch.cuda.synchronize()
    ms = []
    for _ in range(5):
        hits, nu = ctx.kmsg_scan_device(d.data_ptr(), len(buf), cap=1 << 20)
        ms.append(ctx.scan_kernel_ms())
    print(name, "hits", len(hits), "cands", ctx.scan_stats()["candidates"], "filter/prefix/match ms", np.array(ms[1:]).mean(axis=0).round(4))
    del d
all_hits = synth.hit_lines()
edge = list(synth.EDGE_LINES)
kinds = {"none": None, "plain79": [all_hits[3]], "xid14": [all_hits[7]], "ext149": [all_hits[29]], "ext145pid": [all_hits[30]], "xid154": [all_hits[20]],
         "fallenR4": [all_hits[13]], "fallenR3multi": [all_hits[12]], "first42": all_hits[:42], "edge": edge}
for name, lines in kinds.items():
    if lines is None:
        run(name, synth.dmesg_buffer(4 << 20, hit_every=10 ** 9) * 25)
        continue
    synth.hit_lines = lambda lines=lines: list(lines)
    synth.EDGE_LINES = []
    run(name, synth.dmesg_buffer(4 << 20, hit_every=1000) * 25)
