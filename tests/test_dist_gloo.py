"""CPU: the N>1 host path (sharding + the single all-gather of per-shape records) with world_size 2 over gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sednet_b200 import shard


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.shard_range(total, rank, world)
    B, N, S = hi - lo, 40, 3
    g = torch.Generator().manual_seed(100 + rank)
    labels = torch.randint(0, S, (B, N), generator=g)
    status = torch.zeros((B, S), dtype=torch.int32)
    rec = shard.make_records(torch.arange(lo, hi), torch.full((B,), S), status, torch.full((B, S), 0.5),
                             torch.full((B,), 0.1 * (rank + 1)), labels)
    table = shard.gather_records(rec, (total + world - 1) // world)
    q.put((rank, table.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_over_gloo():
    world, total = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    a, b = res[0], res[1]
    assert a.shape == (total, len(shard.RECORD_FIELDS))
    assert (a == b).all()                       # every rank ends with the same table
    assert a[:, 0].tolist() == list(range(total))
    assert abs(a[:3, 4] - 0.1).max() < 1e-6 and abs(a[3:, 4] - 0.2).max() < 1e-6   # rank 0 owns 3 shapes, rank 1 owns 2
