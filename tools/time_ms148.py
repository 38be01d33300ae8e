"""Stage timings of the mean-shift stack on a 148-column embedding (one 10 000-point cloud and a batch of 8)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import oracle_hpnet as OH
from sednet_b200.src import _lib
dev = torch.device("cuda"); N, d = 10000, 148
def timed(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
for B in (1, 8):
    Xs = []
    for b in range(B):
        feat, v, types, edges = OH.hpnet_case(10 + b, N)
        Xs.append(torch.nn.functional.normalize(OH.hpnet_combine(feat, v, torch.tensor(0.3), types, edges, 0.5, 1000)[0], p=2, dim=1))
    X = torch.stack(Xs).contiguous().to(dev)
    kth = torch.empty((B, N), device=dev); bw = torch.empty(B, device=dev)
    out, tmp = torch.empty_like(X), torch.empty_like(X)
    labels = torch.empty((B, N), dtype=torch.int64, device=dev); ids = torch.empty((B, 64), dtype=torch.int32, device=dev)
    nc = torch.empty(B, dtype=torch.int32, device=dev); nl = torch.empty(B, dtype=torch.int32, device=dev)
    cen = torch.empty((B, 64, d), device=dev)
    ws = torch.empty(_lib.load().sed_ms_nms_workspace_bytes(B, N), dtype=torch.uint8, device=dev)
    t_bw = timed(lambda: _lib.call("sed_ms_bandwidth", _lib.ptr(X), B, N, d, 150, 0.003, _lib.ptr(kth), _lib.ptr(bw), _lib.stream()))
    t_sh = timed(lambda: _lib.call("sed_ms_shift", _lib.ptr(X), _lib.ptr(bw), B, N, d, 50, 0, 3, _lib.ptr(out), _lib.ptr(tmp), _lib.stream()))
    t_nms = timed(lambda: _lib.call("sed_ms_nms", _lib.ptr(out), _lib.ptr(X), _lib.ptr(bw), B, N, d, 64, _lib.ptr(labels), _lib.ptr(ids), _lib.ptr(nc), _lib.ptr(nl), _lib.ptr(cen), _lib.ptr(ws), _lib.stream()))
    print(f"B={B}: bandwidth {t_bw:.2f} ms  shift(50 it, tc192) {t_sh:.2f} ms  nms {t_nms:.2f} ms   labels {nl.tolist()}")
