"""Sweeps the prune window / soft mark of the streaming selection kernels (env-tunable) -- one subprocess per setting."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys, torch, numpy as np
sys.path[:0] = [%r, %r]
from sednet_b200 import synth
from sednet_b200.src import _lib, PointNet
dev = torch.device("cuda"); B, N = 8, 10000
x = torch.from_numpy(np.random.default_rng(0).normal(size=(B, 64, N)).astype(np.float32)).to(dev)
X = torch.stack([torch.from_numpy(synth.make_embedding(synth.make_cloud(400 + b, N, n_patches=12)[2], 128, 0.02, b)) for b in range(B)]).to(dev)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
idx = torch.empty((B, N, 64), dtype=torch.int32, device=dev)
kth = torch.empty((B, N), device=dev); bw = torch.empty(B, device=dev)
a = t(lambda: _lib.call("sed_knn_l2", _lib.ptr(x), B, 64, N, 64, _lib.ptr(idx), 0, _lib.stream()))
b = t(lambda: _lib.call("sed_ms_bandwidth", _lib.ptr(X), B, N, 128, 150, 0.003, _lib.ptr(kth), _lib.ptr(bw), _lib.stream()))
print(f"knn {a:.3f} ms  bandwidth {b:.3f} ms  bw0 {float(bw[0]):.6f} idxsum {int(idx.long().sum())}")
''' % (ROOT, os.path.join(ROOT, "tests"))
for env in ({}, {"SEDNET_B200_SS_WIN": "12"}, {"SEDNET_B200_SS_WIN": "40"}, {"SEDNET_B200_SS_WIN": "56"},
            {"SEDNET_B200_SS_SOFT": "100"}, {"SEDNET_B200_SS_SOFT": "140"}, {"SEDNET_B200_SS_WIN": "40", "SEDNET_B200_SS_SOFT": "140"},
            {"SEDNET_B200_KS_WIN": "4"}, {"SEDNET_B200_KS_WIN": "24"}, {"SEDNET_B200_KS_WIN": "40"}, {"SEDNET_B200_KS_WIN": "56"}):
    r = subprocess.run([sys.executable, "-c", code], env={**os.environ, **env}, capture_output=True, text=True)
    print(env, r.stdout.strip() or r.stderr[-300:], flush=True)
