"""Times compute_entropy (10 000 points, the driver's CHUNK = 1000) and the mean-shift stack on a 148-column embedding."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import oracle_hpnet as OH
from sednet_b200.src import _lib, smooth_normal_matrix as snm
from sednet_b200.src.mean_shift import MeanShift
dev = torch.device("cuda")
feat, v, types, edges = OH.hpnet_case(3, 10000)
def timed(fn, reps=10):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
for name, x in (("features K=128", feat), ("spectral K=12", v), ("types+edges K=8", torch.cat((types.exp(), torch.softmax(edges, -1)), -1))):
    xd = x.to(dev)
    print(f"compute_entropy {name}: {timed(lambda: snm.compute_entropy(xd, CHUNK=1000)):.3f} ms")
t0 = time.perf_counter()
with torch.no_grad():
    OH.compute_entropy(v[:, :5000], CHUNK=1000)
print(f"CPU restatement (K=12, 5000 x 5000 pairs): {(time.perf_counter() - t0) * 1e3:.0f} ms")
emb = OH.hpnet_combine(feat[:, :, :], v, torch.tensor(0.3), types, edges, 0.5, 1000)
X = torch.nn.functional.normalize(emb[0], p=2, dim=1).contiguous().to(dev)
ms = MeanShift()
t = timed(lambda: ms.mean_shift(X, 10000, 0.015, 50), reps=2)
print(f"mean_shift on (10000,148), 50 iterations, FFMA path: {t:.1f} ms")
X128 = torch.nn.functional.normalize(feat[0], p=2, dim=1).contiguous().to(dev)
print(f"mean_shift on (10000,128), tensor-core path: {timed(lambda: ms.mean_shift(X128, 10000, 0.015, 50), reps=2):.1f} ms")
# ---- the branch of hpnet_process that builds the spectral vectors (farthest-50 table, factored affinity operator, LOBPCG)
from sednet_b200 import synth
p, nrm, lab, typ, _ = synth.make_cloud(1234, 10000)
P, Nn = torch.from_numpy(p)[None].to(dev), torch.from_numpy(nrm)[None].to(dev)
X0 = torch.randn((1, 10000, 12), generator=torch.Generator().manual_seed(0))
fd, td, ed = feat.to(dev), types.to(dev), edges.to(dev)
def build():
    return snm.hpnet_process(fd, P, Nn, id=None, types=td, edges=ed, normal_smooth_w=0.5, CHUNK=1000, X=X0)
build(); torch.cuda.synchronize()
t0 = time.perf_counter(); build(); torch.cuda.synchronize()
print(f"hpnet_process, no cache (10000 points: farthest-50 + affinity + 10 LOBPCG steps + 3 entropies): {(time.perf_counter() - t0) * 1e3:.1f} ms wall")
print(f"  knn_idx (farthest 50): {timed(lambda: snm.knn_idx(P, 50), reps=3):.3f} ms")
op = snm.construction_affinity_matrix_normal(P, Nn)
Y = torch.randn((10000, 36), device=dev)
op.matmul(0, Y)
print(f"  affinity block product (N = 10000, 36 columns): {timed(lambda: op.matmul(0, Y)):.3f} ms   "
      f"(in-degree max {int(torch.bincount(op.idx[0][op.w[0] != 0].flatten()).max())})")
t0 = time.perf_counter()
with torch.no_grad():
    A = OH.construction_affinity_matrix_normal(torch.from_numpy(p)[None, :3000], torch.from_numpy(nrm)[None, :3000])
    torch.lobpcg(A, k=12, niter=10, X=X0[:, :3000])
print(f"CPU restatement at N = 3000 (dense matrices + torch.lobpcg): {(time.perf_counter() - t0) * 1e3:.0f} ms")
