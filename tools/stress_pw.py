"""Stress: one pointwise layer (sed_pointwise_forward) launched `reps` times, synchronised and compared with
torch.nn.functional.conv1d every time.   python tools/stress_pw.py B N Cin Cout raw|aff [stats] [reps]"""
import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from sednet_b200.src import _lib

B, N, Cin, Cout = (int(a) for a in sys.argv[1:5])
aff = sys.argv[5] == "aff"
stats_on = len(sys.argv) > 6 and sys.argv[6] == "stats"
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 20
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(B, Cin, N, device=dev, generator=g)
W = torch.randn(Cout, Cin, device=dev, generator=g) / Cin ** 0.5
bias = torch.randn(Cout, device=dev, generator=g)
a = torch.rand(B, Cin, device=dev, generator=g) + 0.5 if aff else None
s = torch.randn(B, Cin, device=dev, generator=g) if aff else None
xin = F.relu(a[:, :, None] * x + s[:, :, None]) if aff else x
ref = F.conv1d(xin, W[:, :, None], bias)
P = (N + 127) // 128
bad = 0
for r in range(reps):
    y = torch.full((B, Cout, N), float("nan"), device=dev)
    stats = torch.zeros(B, P, (Cout + 31) // 32, 2, device=dev, dtype=torch.float64) if stats_on else None
    mm = torch.zeros(B, P, Cout, 2, device=dev) if stats_on else None
    try:
        _lib.call("sed_pointwise_forward", _lib.ptr(x), Cin * N, _lib.ptr(W), Cin, _lib.ptr(bias), _lib.ptr(a), _lib.ptr(s),
                  1 if aff else 0, _lib.ptr(y), Cout * N, _lib.ptr(stats), _lib.ptr(mm), B, Cin, Cout, N, _lib.stream())
        torch.cuda.synchronize()
    except RuntimeError as e:
        print(f"rep {r}: CUDA failure: {str(e)[:120]}")
        sys.exit(1)
    err = float((y - ref).abs().max())
    if not err < 1e-3:
        nb = int(((y - ref).abs() > 1e-3).sum() + torch.isnan(y).sum())
        print(f"rep {r}: max err {err} ({nb} bad values)")
        bad += 1
print(f"B={B} N={N} {Cin}->{Cout} {'aff' if aff else 'raw'}{' stats' if stats_on else ''}: {reps} reps, {bad} wrong")
