"""does k_scan_match speed up when the lanes of a warp see identical lines?  (hypothesis behind sorting candidates by line shape)"""
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
    print(name, "bytes", len(buf), "hits", len(hits), "cands", ctx.scan_stats()["candidates"], "filter/prefix/match ms", np.array(ms[1:]).mean(axis=0).round(4))
    del d
fix = [l for l in synth.golden("xid_kmsg.json")["dmesg_xid_119"]["lines"] if l and "Xid" in l]
print(len(fix), "fixture Xid lines; distinct:", len(set(fix)))
rng = np.random.default_rng(1)
noise = lambda: ("[%12.6f] " % rng.random() + "x" * int(rng.integers(40, 200))).encode()
def build(lines, n_hits=25000, pad_every=40):
    out = []
    for i in range(n_hits):
        out.append(lines[i % len(lines)].encode())
        for _ in range(pad_every): out.append(noise())
    return b"\n".join(out) + b"\n"
run("one line repeated      ", build(fix[:1]))
run("4 distinct lines       ", build(list(dict.fromkeys(fix))[:4]))
run("all distinct fixture   ", build(list(dict.fromkeys(fix))))
ext = [l for l in synth.hit_lines() if " 149, " in l or " 145, " in l]
run("one extended line      ", build(ext[:1]))
run("16 distinct extended   ", build(list(dict.fromkeys(ext))[:16]))
