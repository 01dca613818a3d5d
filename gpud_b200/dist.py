"""Host-side multi-GPU plumbing (one process per GPU, torch.distributed for rendezvous / collectives).

Only two things on this path cross GPUs:
  * the whole-box NVLink/fabric view: one 128-byte record per GPU, all-gathered (NCCL over NVLink) and evaluated
    redundantly on every rank  — replaces the single-process loops of nvlink/component.go:164-311 and
    fabric-manager/fabric_state.go:67-113;
  * an optional split of one large kmsg buffer across ranks at unit boundaries, hits concatenated in unit order.
Ring data is never exchanged: every GPU reduces its own counter stream.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def split_at_units(buf: bytes, world: int, raw_kmsg: bool = False) -> List[Tuple[int, int]]:
    """Byte ranges [b, e) for `world` ranks, cut just after a unit separator ('\\n'; in RAW_KMSG mode a '\\n' that is not
    followed by ' ', so a record and its continuation lines stay together).  Every range except the last ends with its
    separator, so the ranks' unit counts add up exactly: units(whole) = sum(units(part_i) - 1) + 1."""
    n = len(buf)
    cuts = [0]
    for r in range(1, world):
        pos = max(cuts[-1], (n * r) // world)
        while True:
            nl = buf.find(b"\n", pos)
            if nl < 0:
                pos = n
                break
            pos = nl + 1
            if not raw_kmsg or pos >= n or buf[pos:pos + 1] != b" ":
                break
        cuts.append(pos)
    cuts.append(n)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def merge_hits(parts: Sequence[Tuple[int, int, list]]):
    """parts[i] = (byte_offset_of_part, n_units_in_part, hits) with hits as dicts carrying 'line' and 'offset' relative to
    the part.  Returns (hits with global line/offset, n_units_total)."""
    out, line0 = [], 0
    for i, (off, n_units, hits) in enumerate(parts):
        for h in hits:
            g = dict(h)
            g["line"] = h["line"] + line0
            g["offset"] = h["offset"] + off
            out.append(g)
        # a part that ends with its separator counts one trailing empty unit that really belongs to the next part
        line0 += n_units - 1 if i + 1 < len(parts) else n_units
    return out, line0


def gather_fabric(ctx, raw, at_least: int = 0, group=None, device=None):
    """pack kernel -> all_gather_into_tensor (NCCL) -> verdict kernel, one process per GPU.  Returns FabricVerdict."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    send = torch.zeros(128, dtype=torch.uint8, device=dev)
    table = torch.zeros(128 * world, dtype=torch.uint8, device=dev)
    # the library treats a NULL stream as "use my own": give it a real stream so kernels and the collective are ordered
    cur = torch.cuda.current_stream(dev)
    side = cur if cur.cuda_stream != 0 else torch.cuda.Stream(device=dev)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        ctx.fabric_pack(raw, send.data_ptr(), dev=dev.index, stream=side.cuda_stream)   # K7 writes straight into the send buffer
        dist.all_gather_into_tensor(table, send, group=group)
        v = ctx.fabric_verdict(table.data_ptr(), world, at_least, dev=dev.index, stream=side.cuda_stream)
    cur.wait_stream(side)
    return v
