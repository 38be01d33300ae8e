"""Times sed_ms_bandwidth (K = 150) on planted and on unstructured embeddings, batch of 8 x 10 000.  Env overrides apply."""
import os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from sednet_b200 import synth
from sednet_b200.src import _lib
dev = torch.device("cuda"); B, N = 8, 10000
Xp = torch.stack([torch.from_numpy(synth.make_embedding(synth.make_cloud(400 + b, N, n_patches=12)[2], 128, 0.02, b)) for b in range(B)]).to(dev)
g = torch.Generator().manual_seed(0)
Xu = torch.nn.functional.normalize(torch.randn((B, N, 128), generator=g) * 0.05 + torch.randn((B, 1, 128), generator=g), dim=2).to(dev).contiguous()
kth = torch.empty((B, N), device=dev); bw = torch.empty(B, device=dev)
for name, X in (("planted", Xp), ("blob", Xu)):
    fn = lambda: _lib.call("sed_ms_bandwidth", _lib.ptr(X), B, N, 128, 150, 0.003, _lib.ptr(kth), _lib.ptr(bw), _lib.stream())
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: bandwidth {e0.elapsed_time(e1) / 10:.3f} ms  bw0 {float(bw[0]):.7f} sum {float(bw.double().sum()):.7f}")
