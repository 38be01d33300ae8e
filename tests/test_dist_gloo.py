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


def test_record_checksum_is_numbering_invariant():
    """The per-shape label checksum of the gathered records compares PARTITIONS: relabelling a cloud's segments (which of
    several identical converged points becomes a centre may change with the batch composition) leaves it unchanged, moving a
    point to another segment changes it."""
    import numpy as np
    import oracle as O
    g = torch.Generator().manual_seed(1)
    labels = torch.randint(0, 7, (3, 400), generator=g)
    c = shard.canonical_labels(labels)
    for b in range(3):
        assert (c[b].numpy() == O.canonical_labels(labels[b].numpy())).all()
    perm = torch.tensor([4, 2, 6, 0, 1, 5, 3])
    args = lambda L: (torch.arange(3), torch.full((3,), 7), torch.zeros((3, 7), dtype=torch.int32), torch.zeros((3, 7)), torch.ones(3), L)
    r0, r1 = shard.make_records(*args(labels)), shard.make_records(*args(perm[labels]))
    assert torch.equal(r0, r1)
    moved = labels.clone(); moved[1, 17] = (moved[1, 17] + 1) % 7
    assert not torch.equal(shard.make_records(*args(moved))[:, 5], r0[:, 5])
