"""Times sed_ms_shift alone for several batch sizes (wave-quantisation experiments).  python tools/prof_ms.py [prec] [B ...]"""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from sednet_b200 import synth
from sednet_b200.src import _lib
prec = int(sys.argv[1]) if len(sys.argv) > 1 else 3
Bs = [int(a) for a in sys.argv[2:]] or [1, 4, 7, 8, 9]
dev = torch.device("cuda")
N, d, iters = 10000, 128, 50
_, _, lab, _, _ = synth.make_cloud(407, N, n_patches=14, min_pts=100)
X1 = torch.from_numpy(synth.make_embedding(lab, d, 0.02, 5)).to(dev)
for B in Bs:
    X = X1.unsqueeze(0).repeat(B, 1, 1).contiguous()
    bw = torch.full((B,), 0.3, device=dev)
    out, tmp = torch.empty_like(X), torch.empty_like(X)
    def run():
        _lib.call("sed_ms_shift", _lib.ptr(X), _lib.ptr(bw), B, N, d, iters, 0, prec, _lib.ptr(out), _lib.ptr(tmp), _lib.stream())
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); run(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    tiles = B * ((N + 127) // 128)
    print(f"prec {prec} B {B}: {ms:.2f} ms / {iters} it = {ms / iters * 1e3:.1f} us per iteration; {tiles} q-tiles = {tiles / 148:.2f} waves; "
          f"{ms / iters * 1e3 / (tiles / 148):.1f} us per wave-equivalent")
