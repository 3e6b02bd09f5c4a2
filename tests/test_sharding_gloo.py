"""CPU, world_size 2 over gloo: the N>1 host path - contiguous problem sharding and the scalar gather that
bench.py --gpus N uses (NCCL there).  The solve itself has no collective, so a shard's results are by construction
those of the single-GPU run of the same problems (tested on the GPU as batch-composition invariance)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B_total, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "ilqr-admm_b200"))
    from isls_b200.sharding import gather_scalars, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(B_total, rank, world)
    # stand-in for the per-problem results of this rank's solve: a deterministic function of the global index
    idx = torch.arange(lo, hi, dtype=torch.float64)
    cost = idx * 0.5 + 1.0
    status = (torch.arange(lo, hi) % 7).to(torch.int32)
    g_cost = gather_scalars(cost, B_total)
    g_status = gather_scalars(status, B_total)
    q.put((rank, lo, hi, g_cost.numpy(), g_status.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_batch():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "ilqr-admm_b200"))
    from isls_b200.sharding import shard_range
    for B in (1, 7, 64, 65536, 65537):
        for world in (1, 2, 4, 8):
            r = [shard_range(B, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


def test_gather_scalars_world2_gloo():
    world, B_total = 2, 101          # ragged on purpose
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    exp_cost = np.arange(B_total) * 0.5 + 1.0
    exp_status = (np.arange(B_total) % 7).astype(np.int32)
    for rank, lo, hi, g_cost, g_status in res:
        assert np.array_equal(g_cost, exp_cost)
        assert np.array_equal(g_status, exp_status)
    assert sorted((lo, hi) for _, lo, hi, _, _ in res) == [(0, 51), (51, 101)]
