"""N > 1 host logic on CPU: world_size 2 and 3, gloo backend, rendezvous on 127.0.0.1.
The ranks split one kmsg buffer at unit boundaries, scan their part (the oracle stands in for the device scan here — this
test is about the sharding/merge plumbing, the CUDA scan itself is covered by the gpu tests), all-gather the hits and every
rank checks the merged result against a single-process scan.  The fabric verdict is checked the same way: every rank
contributes its own record and evaluates the gathered table."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import synth
from gpud_b200 import dist as gd
from oracle import fabric as OF
from oracle import pyoracle as O
from test_fabric_host import scenario


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, raw_mode, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    reported = False
    try:
        buf = synth.raw_kmsg_buffer(1500) if raw_mode else synth.dmesg_buffer(300_000, hit_every=50)
        b, e = gd.split_at_units(buf, world, raw_kmsg=raw_mode)[rank]
        part = buf[b:e]
        if raw_mode:
            hits, n_units = O.scan_raw_kmsg(part)
        else:
            hits, n_units = O.scan_lines(part), part.count(b"\n") + 1
        slim = [{"line": h["line"], "offset": h["offset"], "kind": h["kind"], "code": h["code"], "device": h["device"]} for h in hits]
        gathered = [None] * world
        dist.all_gather_object(gathered, (b, n_units, slim))
        merged, total_units = gd.merge_hits(gathered)
        if raw_mode:
            want, want_units = O.scan_raw_kmsg(buf)
        else:
            want, want_units = O.scan_lines(buf), buf.count(b"\n") + 1
        ok = total_units == want_units and [(h["line"], h["offset"], h["kind"], h["code"], h["device"]) for h in merged] == \
            [(h["line"], h["offset"], h["kind"], h["code"], h["device"]) for h in want]
        # fabric: each rank contributes one record; every rank evaluates the same table
        gpus, at_least = scenario("all_pairs_ns_with_threshold", world)
        mine = gpus[rank]
        table = [None] * world
        dist.all_gather_object(table, mine)
        v = OF.verdict(table, at_least)
        vt = torch.tensor([v["nvlink_health"], v["active"], v["p2p_ok_pairs"]])
        dist.all_reduce(vt, op=dist.ReduceOp.MAX)          # identical on every rank, so MAX == own
        ok = ok and vt.tolist() == [v["nvlink_health"], v["active"], v["p2p_ok_pairs"]] and len(want) > 10
        q.put((rank, ok, len(merged)))
        reported = True
    finally:
        if not reported:
            q.put((rank, False, -1))
        dist.destroy_process_group()


@pytest.mark.parametrize("world,raw_mode", [(2, False), (3, False), (2, True)])
def test_sharded_scan_and_fabric_gloo(world, raw_mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, raw_mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert len({n for _, _, n in res}) == 1


def test_split_at_units_properties():
    buf = synth.dmesg_buffer(100_000, hit_every=40)
    for world in (1, 2, 4, 8):
        parts = gd.split_at_units(buf, world)
        assert parts[0][0] == 0 and parts[-1][1] == len(buf)
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        for b, e in parts[:-1]:
            assert e == b or buf[e - 1:e] == b"\n"
    raw = synth.raw_kmsg_buffer(500)
    for b, e in gd.split_at_units(raw, 4, raw_kmsg=True)[:-1]:
        assert raw[e - 1:e] == b"\n" and raw[e:e + 1] != b" "
