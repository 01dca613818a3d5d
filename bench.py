#!/usr/bin/env python3
"""bench.py — counter-samples/sec of the telemetry hot path (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (config.workload, BASELINE configs[3] per GPU): 512 fields x 1 Mi-sample f64 ring (4 GiB, >> the 126 MB L2, so
every step streams from HBM), tumbling W = 1000 windows, all six aggregates fused (min/max/mean/EMA/p99/n_over).
A "step" = one pass of the hot path over the whole ring: the fused window-reduce kernel + the EMA carry kernel
(+ at N > 1 the 128-byte NVLink/fabric summary all-gather and verdict).  Weak scaling: every GPU owns its own ring.

`value`  : whole-job samples/s with the ring already resident in HBM (CUDA events on the launch stream, max over ranks).
`e2e`    : same metric through the C ABI from pinned HOST rows: gpud_ring_push (H2D + append kernel) + reduce + D2H
           of every aggregate, all inside the timed region.
`roofline`: algorithmic bytes (8 B/sample) / the fused kernel's CUDA-event time, against MEASURED_PEAKS.json hbm_gbs.
`cpu_baseline`: the C oracle (oracle/oracle.c, "port": the Go reference cannot be built here and has no windowed
           aggregation at all) on the host cores, bounded sample, rank 0 only.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

F, CAP, W = 512, 1 << 20, 1000
LO, HI, THR = 30000, 90000, 88000.0   # synthetic raw counter samples: integer readings (a power gauge in mW), threshold in the same unit
METRIC = "counter-samples/sec"
FALLBACK_HBM_GBS = 6650.0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_window_reduce launch on this workload, from the committed
    `ncu --set full` capture (profiles/window_reduce_traffic.json, written by tools/ncu_summary.py); None if absent."""
    p = os.path.join(ROOT, "profiles", "window_reduce_traffic.json")
    try:
        d = json.load(open(p))
        return float(d["dram_bytes_read"]) + float(d["dram_bytes_write"])
    except Exception:
        return None


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML every 5 ms during the timed region (the recipe's clocks line;
    NVML instead of the nvidia-smi CLI because the timed region lasts tens of milliseconds)."""

    def __init__(self, index):
        self.index, self.sm, self.reasons, self.max_mhz, self._stop, self._t = index, [], set(), None, False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4, "hw_power_brake": 0x80}
        while not self._stop:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        self._stop = True
        if self._t:
            self._t.join(timeout=1.0)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "source": "nvml"}


def cpu_leg(steps, warmup, sample_fields, threads=0):
    """Times the C oracle on a bounded sample of the same workload: `sample_fields` fields x CAP samples per step."""
    from oracle import coracle
    rng = np.random.default_rng(0x67707564)
    ring = rng.integers(LO, HI, (sample_fields, CAP)).astype(np.float64)     # the same raw-counter stream shape as the GPU arm
    thr = np.full(sample_fields, THR)
    cores = coracle.max_threads() if threads <= 0 else threads
    for _ in range(warmup):
        coracle.windows_fields(ring[: max(1, sample_fields // 8)], W, thr, threads=cores)
    t0 = time.perf_counter()
    for _ in range(steps):
        coracle.windows_fields(ring, W, thr, threads=cores)
    dt = (time.perf_counter() - t0) / steps
    return sample_fields * CAP / dt, dt, cores


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-fields", type=int, default=0, help="fields in the CPU sample (0 = auto, about 10-30 s of CPU work)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-scan", action="store_true")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(a.warmup, 3) if a.impl == "b200" else a.warmup
    config = {"workload": "ring 512 fields x 1Mi f64 samples per GPU (4 GiB), W=1000 tumbling, fused min/max/mean/ema/p99/n_over",
              "samples": "synthetic raw counters: integer readings uniform in [30000, 90000) held as float64 in the ring",
              "n_fields": F, "capacity": CAP, "window": W, "parallelism": "one ring per GPU, no data-path collective (weak)",
              "l2": "input 4 GiB per GPU >> 126 MB L2; every step re-streams HBM"}

    if a.impl == "reference":
        # The reference's own CPU path for this metric does not exist (no windowed aggregation in gpud) and Go cannot be
        # built here; the timed arm is the C oracle port with every host thread.  Rank 0 only.
        if rank != 0:
            return 0
        from oracle import coracle
        cores = coracle.max_threads()
        fields = a.cpu_fields or max(8, min(F, cores * 8))
        v, dt, cores = cpu_leg(max(1, a.steps), a.warmup, fields)
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "samples/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                                 "sample": "%d of 512 fields x 1Mi samples per step (C oracle, pthreads over fields)" % fields},
                "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import gpud_b200 as g
    import synth

    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ctx = g.Context([local])
    thr = np.full(F, THR)
    ring = g.Ring(ctx, F, CAP, W, thresholds=thr, dev=local)
    stream = torch.cuda.Stream(device=dev)            # a real (non-NULL) stream: the library launches on it and torch events time it
    torch.cuda.set_stream(stream)
    ring.set_stream(stream.cuda_stream)
    assert stream.cuda_stream != 0

    # ---- synthetic resident data: gauge-like values generated on the device, appended through the real append kernel ----
    gen = torch.Generator(device=dev).manual_seed(0x67707564 + rank)
    chunk = 1 << 16
    for i in range(CAP // chunk):
        t = torch.randint(LO, HI, (chunk, F), dtype=torch.int32, device=dev, generator=gen).to(torch.float64)
        ring.push_device(t.data_ptr(), chunk)
    torch.cuda.synchronize()
    del t

    # fabric leg (N > 1): 128-byte summary per GPU, all-gather over NCCL, verdict kernel
    fab_send = torch.zeros(128, dtype=torch.uint8, device=dev)
    fab_all = torch.zeros(128 * world, dtype=torch.uint8, device=dev)
    raw = g.FabricRaw()
    raw.gpu_index, raw.nvlink_supported, raw.system_expected_nvlink, raw.n_links = rank, 1, 1, 18
    for i in range(18):
        raw.link_feature_enabled[i] = 1
    for j in range(16):
        raw.p2p_status[j] = 0 if (j < world and j != rank) else 0xFF
    raw.fabric_valid, raw.fabric_state, raw.fabric_summary, raw.fabric_health_mask = 1, 3, 1, 0xAA

    def step():
        ring.reduce()
        if world > 1:
            ctx.fabric_pack(raw, fab_send.data_ptr(), dev=local, stream=stream.cuda_stream)
            dist.all_gather_into_tensor(fab_all, fab_send)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms = []
    barrier()
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    # per-kernel device time of the dominant kernel: re-run the same step K times reading the library's own events
    for _ in range(a.steps):
        ring.reduce()
        kms.append(ring.kernel_ms())
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        v = ctx.fabric_verdict(fab_all.data_ptr(), world, world, dev=local)
        assert v.nvlink_health == 0 and v.active == world, v.as_dict()
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / a.steps
    samples_step = F * CAP * world
    value = samples_step / (ms_step * 1e-3)
    k_reduce = float(np.mean([k[0] for k in kms]))
    k_carry = float(np.mean([k[1] for k in kms]))
    peak, how = peaks()
    achieved = F * CAP * 8 / (k_reduce * 1e-3) / 1e9

    # ---- e2e: pinned host rows -> push (H2D + append) -> reduce -> D2H of all aggregates, every step ----
    # Headline e2e: the poller hands over RAW NVML samples in the getter's own type (uint32) through gpud_ring_push_raw; the
    # widening to float64 happens in the append kernel.  e2e_f64: the same stream already widened on the host (gpud_ring_push).
    e2e = e2e_f64 = None
    if not a.no_e2e:
        nw = (CAP + W - 1) // W
        outs = {k: np.empty((F, nw), dtype=np.uint32 if k == "n_over" else np.float64) for k in g.OPS}
        d2h = sum(o.nbytes for o in outs.values())

        def run_e2e(host, rows_chunk, dtype_code, esz, steps):
            def e2e_step():
                for _ in range(CAP // rows_chunk):
                    ring.push_raw_ptr(host.data_ptr(), rows_chunk, dtype_code)
                ring.reduce()
                for k in g.OPS:
                    ctx._check(ring._L.gpud_ring_read(ring._h, g.OPS[k], C.c_void_p(outs[k].ctypes.data), outs[k].nbytes))
            e2e_step()
            barrier()
            t0 = time.perf_counter()
            e0.record()
            for _ in range(steps):
                e2e_step()
            e1.record()
            barrier()
            ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / steps
            if world > 1:
                t = torch.tensor([ms], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            return {"value": samples_step / (ms * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": F * CAP * esz, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms, "steps": steps}

        e2e_steps = max(2, min(a.steps, 5))
        host = torch.randint(LO, HI, (1 << 19, F), dtype=torch.int32).pin_memory()        # 1 GiB pinned, pushed 2x per step = the whole ring
        e2e = run_e2e(host, 1 << 19, g.DTYPES["uint32"], 4, e2e_steps)
        e2e["sample_dtype"] = "uint32 raw NVML counters (gpud_ring_push_raw), widened to f64 on the device"
        del host
        host = torch.randint(LO, HI, (1 << 18, F), dtype=torch.int32).to(torch.float64).pin_memory()   # 1 GiB pinned, pushed 4x per step
        e2e_f64 = run_e2e(host, 1 << 18, g.DTYPES["float64"], 8, 2)
        e2e_f64["sample_dtype"] = "float64 widened on the host (gpud_ring_push)"
        del host

    cpu = None
    if rank == 0 and world == 1:
        from oracle import coracle
        cores = coracle.max_threads()
        fields = a.cpu_fields or max(8, min(F, cores * 8))
        v, dt, cores = cpu_leg(1, 1, fields)
        cpu = {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": "%d of 512 fields x 1Mi samples, one pass (C oracle oracle/oracle.c, pthreads over fields)" % fields}

    # ---- secondary workload (BASELINE configs[2]): Xid/SXid scan of a 100 MiB synthetic dmesg buffer, rank 0 at N=1 only ----
    scan = None
    if rank == 0 and world == 1 and not a.no_scan:
        try:
            unit = synth.dmesg_buffer(4 << 20, hit_every=1000)
            buf = unit * 25
            d = torch.frombuffer(bytearray(buf), dtype=torch.uint8).to(dev)
            torch.cuda.synchronize()
            ms = []
            for _ in range(4):
                hits, n_units = ctx.kmsg_scan_device(d.data_ptr(), len(buf), cap=1 << 17, dev=local)
                ms.append(ctx.scan_kernel_ms(dev=local))
            filt, pre, mat = np.array(ms[1:]).mean(axis=0)
            harr = (g.XidHit * (1 << 17))()
            ctx.kmsg_scan_c(buf, harr, 1 << 17, dev=local)
            t0 = time.perf_counter()
            nh2, _ = ctx.kmsg_scan_c(buf, harr, 1 << 17, dev=local)
            e2e_s = time.perf_counter() - t0
            from oracle import coracle
            t0 = time.perf_counter()
            ch, _nl = coracle.scan_lines(buf)
            cpu_s = time.perf_counter() - t0
            same = [(h.line, h.kind, h.code) for h in ch] == [(h.unit_index, h.kind, h.code) for h in hits]
            scan = {"workload": "100 MiB synthetic dmesg (reference fixtures + noise + injected Xid/SXid lines)", "bytes": len(buf), "lines": n_units,
                    "hits": len(hits), "device_ms": float(filt + pre + mat), "device_GBps": len(buf) / float(filt + pre + mat) / 1e6,
                    "filter_kernel_ms": float(filt), "filter_frac_of_hbm_peak": len(buf) / float(filt) / 1e6 / peaks()[0],
                    "e2e_host_ms": e2e_s * 1e3, "e2e_GBps": len(buf) / e2e_s / 1e9, "cpu_oracle_ms": cpu_s * 1e3, "cpu_oracle_GBps": len(buf) / cpu_s / 1e9,
                    "cpu_threads": coracle.max_threads(), "hits_identical_to_oracle": bool(same and nh2 == len(hits))}
            # same buffer with the nccl / peermem matchers switched on: four anchor words instead of two in the filter
            ebuf = synth.ext_buffer(4 << 20, hit_every=1000) * 25
            d2 = torch.frombuffer(bytearray(ebuf), dtype=torch.uint8).to(dev)
            ems = []
            for _ in range(4):
                ehits, _eu = ctx.kmsg_scan_device(d2.data_ptr(), len(ebuf), mode=g.SCAN_LINES | g.SCAN_EXT_MATCHERS, cap=1 << 18, dev=local)
                ems.append(ctx.scan_kernel_ms(dev=local))
            ech, _ = coracle.scan_lines(ebuf, ext=True)
            ef, ep, em = np.array(ems[1:]).mean(axis=0)
            scan["ext_matchers"] = {"bytes": len(ebuf), "hits": len(ehits), "nccl_hits": sum(h.kind == 3 for h in ehits),
                                    "peermem_hits": sum(h.kind == 4 for h in ehits), "device_ms": float(ef + ep + em), "filter_kernel_ms": float(ef),
                                    "hits_identical_to_oracle": [(h.line, h.kind, h.code) for h in ech] == [(h.unit_index, h.kind, h.code) for h in ehits]}
            del d2
            del d
        except Exception as ex:   # the scan leg must never cost the headline line
            scan = {"error": repr(ex)}

    # ---- real ingest (BASELINE configs[1] shape): NVML getters -> pinned uint32 rows -> K1, rank 0 at N=1 only ----
    ingest = None
    if rank == 0 and world == 1 and not a.no_scan:
        try:
            iring = g.Ring(ctx, len(g.POLL_FIELDS), 1 << 16, 1000, thresholds=np.zeros(len(g.POLL_FIELDS)), dev=local)
            poller = g.Poller(ctx, iring, dev=local)
            poller.poll(200)
            poller.poll(5000)
            rows, sec = poller.last_rows()
            iring.reduce()
            iring.sync()
            ingest = {"source": "NVML getters (temperature, power, 3 clocks, 2 utilisations, memory used) via the library's poller",
                      "fields": len(g.POLL_FIELDS), "polls": int(rows.shape[0]), "polls_per_s": rows.shape[0] / sec,
                      "samples_per_s": rows.shape[0] * len(g.POLL_FIELDS) / sec,
                      "last_row": {k: int(v) for k, v in zip(g.POLL_FIELDS, rows[-1])}}
            poller.close()
            iring.close()
        except Exception as ex:
            ingest = {"error": repr(ex)}

    ring.close()
    ctx.close()
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": a.steps, "warmup": warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "gpu_launches": (3 + (1 if world > 1 else 0)) * a.steps,   # window_reduce (specialised + generic tail) + ema_carry (+ fabric pack)
                "clocks": clocks, "e2e": e2e,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(),
                             "kernel": "k_window_reduce", "kernel_ms": k_reduce, "carry_kernel_ms": k_carry, "peak_source": how + " (MEASURED_PEAKS.json hbm_gbs, burst copy)"
                             if how == "measured" else "fallback 6650 GB/s (B200_PROFILING.md)", "algorithmic_bytes_per_launch": F * CAP * 8},
                "cpu_baseline": cpu, "e2e_f64": e2e_f64, "scan": scan, "ingest": ingest}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
