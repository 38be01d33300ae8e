"""How many query rows of the mean-shift iteration are at a bitwise fixed point after t iterations (prec mode 3)?"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from sednet_b200 import synth
from sednet_b200.src import _lib
from sednet_b200.pipeline import Pipeline
dev = torch.device("cuda")
N, d = 10000, 128

def shift(X, bw, it, prec=3):
    out, tmp = torch.empty_like(X), torch.empty_like(X)
    _lib.call("sed_ms_shift", _lib.ptr(X), _lib.ptr(bw), X.shape[0], N, d, it, 0, prec, _lib.ptr(out), _lib.ptr(tmp), _lib.stream())
    return out

def report(name, X, bw):
    prev = shift(X, bw, 5)
    for t in (6, 10, 15, 20, 25, 30, 35, 40, 45, 50):
        a, b = shift(X, bw, t - 1), shift(X, bw, t)
        same = (a == b).all(dim=2).float().mean().item()
        tiles = (a == b).all(dim=2)[:, :9984].reshape(X.shape[0], -1, 128).all(dim=2).float().mean().item()
        print(f"{name}: t={t:2d} rows fixed {same*100:6.2f}%  128-row tiles fixed {tiles*100:6.2f}%  max|d| {float((a-b).abs().max()):.2e}")

_, _, lab, _, _ = synth.make_cloud(407, N, n_patches=14, min_pts=100)
X = torch.from_numpy(synth.make_embedding(lab, d, 0.02, 5)).to(dev)[None].contiguous()
bwv = torch.empty(1, device=dev); kth = torch.empty((1, N), device=dev)
_lib.call("sed_ms_bandwidth", _lib.ptr(X), 1, N, d, 150, 0.003, _lib.ptr(kth), _lib.ptr(bwv), _lib.stream())
print("planted bw", float(bwv))
report("planted", X, bwv)
pts, nrm, _, _ = synth.make_batch(2, N, seed0=1234)
pipe = Pipeline(2, N, 64)
pipe.set_weights(synth.make_state_dict(0), synth.make_state_dict(1, randomize_gn=True))
pipe.run_device(torch.from_numpy(pts).to(dev), torch.from_numpy(nrm).to(dev), 0.015, 50, 3)
Xn = pipe.device_tensor("X")[:1].contiguous(); bwn = pipe.device_tensor("bw")[:1].contiguous()
print("network bw", float(bwn), "labels", int(pipe.device_tensor("n_labels")[0]))
report("network", Xn, bwn)
