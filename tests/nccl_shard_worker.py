"""Worker of test_two_rank_nccl_sharded_equals_single_rank (tests/test_gpu_round2.py): one process per GPU under torchrun.
Each rank runs the C-ABI pipeline on its contiguous shard of 6 small clouds and the per-shape records meet in ONE
all-gather at the end (sednet_b200.shard); rank 0 saves the gathered table."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

TOTAL, NPTS, K, ITERS = 6, 2048, 32, 10


def run_shard(rank, world, dev):
    from sednet_b200 import shard, synth
    from sednet_b200.pipeline import Pipeline
    lo, hi = shard.shard_range(TOTAL, rank, world)
    B = hi - lo
    pts = np.empty((B, NPTS, 3), np.float32); nrm = np.empty_like(pts)
    X = torch.empty((B, NPTS, 128)); pred = np.empty((B, NPTS), np.int32)
    for i, sid in enumerate(range(lo, hi)):                 # every shape is a function of its global id only
        pts[i], nrm[i], lab, typ, _ = synth.make_cloud(500 + sid, NPTS, n_patches=4 + sid % 3, min_pts=200)
        X[i] = torch.from_numpy(synth.make_embedding(lab, 128, 0.02, 900 + sid))
        pred[i] = typ
    pipe = Pipeline(B, NPTS, K, max_segments=64)
    pipe.set_weights(synth.make_state_dict(0), synth.make_state_dict(1, randomize_gn=True))
    P, Nn = torch.from_numpy(pts).to(dev), torch.from_numpy(nrm).to(dev)
    pipe.run_forward(P, Nn)
    pipe.device_tensor_view("X")[:B].copy_(X.to(dev))
    pipe.device_tensor_view("pred_type")[:B].copy_(torch.from_numpy(pred).to(dev))
    pipe.run_cluster(P, Nn, 0.015, ITERS, 1)
    v = lambda name: pipe.device_tensor(name)[:B]
    rec = shard.make_records(torch.arange(lo, hi, device=dev), v("n_labels"), v("status"), v("residual"), v("bw"), v("labels"))
    table = shard.gather_records(rec, (TOTAL + world - 1) // world)
    pipe.close()
    return table


if __name__ == "__main__":
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    table = run_shard(rank, world, dev)
    if rank == 0:
        torch.save(table.cpu(), sys.argv[1])
    dist.barrier()
    dist.destroy_process_group()
