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
