import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import gpud_b200 as g, synth
ctx = g.Context([0])
ctx.scan_phase_timing(True)          # per-phase event timing (plain launches)
for name, buf, mode in (("default", synth.dmesg_buffer(4 << 20, hit_every=1000) * 25, g.SCAN_LINES), ("ext", synth.ext_buffer(4 << 20, hit_every=1000) * 25, g.SCAN_LINES | g.SCAN_EXT_MATCHERS)):
    d = torch.frombuffer(bytearray(buf), dtype=torch.uint8).cuda(); torch.cuda.synchronize()
    ms = []
    for _ in range(6):
        hits, nu = ctx.kmsg_scan_device(d.data_ptr(), len(buf), mode=mode, cap=1 << 20)
        ms.append(ctx.scan_kernel_ms())
    print(os.environ.get("GPUD_SCAN_SORT", "1"), name, "hits", len(hits), "filter/prefix/match ms", np.array(ms[1:]).mean(axis=0).round(4))
