"""Multi-GPU legs (need >= 2 CUDA devices; skipped otherwise).  Run with `gpurun --gpus 2 -- pytest -m gpu tests/test_gpu_multi.py`."""
import os
import socket

import pytest

from oracle import fabric as OF
from test_fabric_host import SCENARIOS, scenario

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _raw(g, d):
    r = g.FabricRaw()
    r.gpu_index, r.nvlink_supported, r.system_expected_nvlink, r.n_links = d["gpu_index"], d["nvlink_supported"], d["system_expected_nvlink"], d["n_links"]
    for i in range(18):
        r.link_feature_enabled[i] = d["link_feature_enabled"][i]
        r.link_replay_errors[i] = d["link_replay_errors"][i]
        r.link_recovery_errors[i] = d["link_recovery_errors"][i]
        r.link_crc_errors[i] = d["link_crc_errors"][i]
    for j in range(16):
        r.p2p_status[j] = d["p2p_status"][j]
    r.fabric_valid, r.fabric_state, r.fabric_summary = d["fabric_valid"], d["fabric_state"], d["fabric_summary"]
    r.fabric_status, r.fabric_health_mask, r.clique_id = d["fabric_status"], d["fabric_health_mask"], d.get("clique_id", 1)
    return r


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_fabric_gather_p2p_single_process():
    """single process, all devices: pack kernels store into every peer's table over NVLink, verdict kernels spin on arrival"""
    import gpud_b200 as g
    n = min(_n_gpus(), 8)
    ctx = g.Context(list(range(n)))
    for rep in range(3):                       # epochs advance; tables are reused
        for name in ("all_pairs_ns_with_threshold", "all_healthy", "zero_active_no_threshold", "fabric_summary_unhealthy" if n > 5 else "all_healthy"):
            gpus, at_least = scenario(name, n)
            at_least = min(at_least, n)
            recs, verdicts = ctx.fabric_gather_p2p([_raw(g, d) for d in gpus], at_least)
            want = OF.verdict(gpus, at_least)
            for v in verdicts:                 # every GPU reaches the same verdict
                vd = v.as_dict()
                for k, val in want.items():
                    assert vd[k] == val, (name, k, vd[k], val)
            assert [r.gpu_index for r in recs] == list(range(n))
    ctx.close()


def _nccl_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    import gpud_b200 as g
    from gpud_b200 import capi, dist as gd
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ok = True
    try:
        ctx = g.Context([rank])
        # (a) plumbing through torch.distributed: pack kernel -> all_gather_into_tensor -> verdict kernel
        for name in ("all_pairs_ns_with_threshold", "all_healthy"):
            gpus, at_least = scenario(name, world)
            at_least = min(at_least, world)
            v = gd.gather_fabric(ctx, _raw(g, gpus[rank]), at_least).as_dict()
            want = OF.verdict(gpus, at_least)
            ok = ok and all(v[k] == val for k, val in want.items())
        # (b) the library's own communicator (dlopen'ed NCCL): unique id from rank 0, broadcast as an object
        uid = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(world, rank, uid[0], dev=rank)
        gpus, at_least = scenario("zero_active_but_p2p_ok", world)
        recs, v = ctx.fabric_gather(_raw(g, gpus[rank]), world, at_least, dev=rank)
        want = OF.verdict(gpus, at_least)
        vd = v.as_dict()
        ok = ok and all(vd[k] == val for k, val in want.items()) and [r.gpu_index for r in recs] == list(range(world))
        ctx.close()
    except Exception as e:  # noqa
        import traceback
        traceback.print_exc()
        ok = False
    finally:
        q.put((rank, ok))
        dist.destroy_process_group()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_fabric_gather_nccl_one_process_per_gpu():
    import torch.multiprocessing as mp
    world = min(_n_gpus(), 8)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
