"""Times the training-path mean-shift (forward with saved states + backward) at the size the triplet loss uses
(src/segment_loss.py:50-56: one cloud, 5 iterations) against the reference formulation's autograd on the same GPU.
python tools/bench_ms_backward.py [N] [iterations]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
from sednet_b200 import synth
from sednet_b200.src.mean_shift import MeanShift
import oracle as O
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda")
torch.backends.cuda.matmul.allow_tf32 = False
lab = np.random.default_rng(3).integers(0, 14, N)
x = torch.from_numpy(synth.make_embedding(lab, 128, 0.02, 11)).to(dev)
W = torch.randn(N, 128, device=dev)
b = torch.tensor(0.2, device=dev)
ms = MeanShift(prec_mode=1)

def ours():
    X = x.clone().requires_grad_(True)
    out, _ = ms.mean_shift_(X, b=b, iterations=iters)
    (out * W).sum().backward()
    return X.grad

def eager():
    X = x.clone().requires_grad_(True)
    (O.ms_shift(X, b, iters) * W).sum().backward()
    return X.grad

for name, fn in (("kernels", ours), ("eager autograd (reference formulation)", eager)):
    g = fn(); torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): g = fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 3:.2f} ms per forward+backward (N {N}, {iters} iterations), peak memory "
          f"{torch.cuda.max_memory_allocated() / 2**20:.0f} MiB")
    if name == "kernels": g0 = g
print("max |grad diff| / max |grad|:", float((g0 - g).abs().max() / g.abs().max()))
