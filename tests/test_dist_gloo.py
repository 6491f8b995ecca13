"""N>1 host logic on CPU: two gloo ranks each own a contiguous shard of the frame sets, compute their
results independently (here with the oracle standing in for the GPU replica) and all-gather [N,K,4]
once at the end — the same shard_range/gather_results code bench.py runs over NCCL."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_items, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    from jarvis_hybridnet_b200 import gather_results, shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s, e = shard_range(n_items, rank, world)
    local = torch.stack([_fake_result(i) for i in range(s, e)]) if e > s else torch.zeros((0, 5, 4))
    full = gather_results(local, n_items)
    if rank == 0:
        q.put(full.numpy())
    dist.barrier()
    dist.destroy_process_group()


def _fake_result(i):
    g = torch.Generator().manual_seed(i)
    return torch.rand((5, 4), generator=g)


def test_two_rank_shard_and_gather():
    n_items, world = 7, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = torch.stack([_fake_result(i) for i in range(n_items)]).numpy()
    assert np.array_equal(got, want)
