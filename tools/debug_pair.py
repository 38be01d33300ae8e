"""Debug: sed_ms_shift outputs for a few shapes, saved per SEDNET_B200_MS_PAIR setting; `cmp` compares two dumps.
python tools/debug_pair.py run TAG | cmp TAG_A TAG_B"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
OUT = os.path.join(ROOT, "gpurun_out")
if sys.argv[1] == "cmp":
    a, b = (torch.load(os.path.join(OUT, f"pair_{t}.pt")) for t in sys.argv[2:4])
    for k in a:
        d = (a[k] - b[k]).abs().max().item()
        print(k, "max abs diff", d, "nan" if torch.isnan(b[k]).any() else "")
    sys.exit(0)
from sednet_b200 import synth
from sednet_b200.src import _lib
tag = sys.argv[2]
dev = torch.device("cuda")
res = {}
for (B, N, iters, mode) in ((1, 256, 1, 1), (2, 1000, 3, 1), (2, 1000, 3, 3), (3, 4100, 5, 1), (8, 10000, 4, 1), (8, 10000, 4, 3)):
    _, _, lab, _, _ = synth.make_cloud(400 + N, N, n_patches=8, min_pts=20)
    X1 = torch.from_numpy(synth.make_embedding(lab, 128, 0.02, 5)).to(dev)
    X = torch.stack([torch.roll(X1, b * 17, 0) for b in range(B)]).contiguous()
    bw = torch.full((B,), 0.3, device=dev) + 0.01 * torch.arange(B, device=dev)
    out, tmp = torch.empty_like(X), torch.empty_like(X)
    _lib.call("sed_ms_shift", _lib.ptr(X), _lib.ptr(bw), B, N, 128, iters, 0, mode, _lib.ptr(out), _lib.ptr(tmp), _lib.stream())
    torch.cuda.synchronize()
    res[f"B{B}_N{N}_it{iters}_m{mode}"] = out.cpu()
    print("done", B, N, iters, mode, float(out.abs().max()), flush=True)
torch.save(res, os.path.join(OUT, f"pair_{tag}.pt"))
