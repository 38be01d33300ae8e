"""Runs the kNN kernels alone at the bench shape (for ncu captures).  python tools/prof_knn.py [B] [k]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from sednet_b200.src import PointNet
from util import cloud_input
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
k = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda")
x = torch.from_numpy(np.random.default_rng(0).normal(size=(B, 64, 10000)).astype(np.float32)).to(dev)
x6 = torch.cat([torch.from_numpy(cloud_input(21 + b, 10000)[4]) for b in range(B)]).to(dev)
for _ in range(3):
    PointNet.knn(x, k, k); PointNet.knn_points_normals(x6, k, k, 1.0)
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record()
for _ in range(5): PointNet.knn(x, k, k)
e1.record()
for _ in range(5): PointNet.knn_points_normals(x6, k, k, 1.0)
e2.record(); torch.cuda.synchronize()
print(f"knn_l2 B{B} k{k}: {e0.elapsed_time(e1)/5:.3f} ms   knn_pn: {e1.elapsed_time(e2)/5:.3f} ms")
