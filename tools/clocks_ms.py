"""SM clock and board power while the mean-shift iteration runs alone (B = 8 x 10 000 points, mode 1 by default) -- the
evidence for "power-capped" in DESIGN.md.   python tools/clocks_ms.py [prec] [seconds]"""
import os, subprocess, sys, threading, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from sednet_b200 import synth
from sednet_b200.src import _lib
prec = int(sys.argv[1]) if len(sys.argv) > 1 else 1
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
dev = torch.device("cuda")
B, N, d, iters = 8, 10000, 128, 50
_, _, lab, _, _ = synth.make_cloud(407, N, n_patches=14, min_pts=100)
X = torch.from_numpy(synth.make_embedding(lab, d, 0.02, 5)).to(dev).unsqueeze(0).repeat(B, 1, 1).contiguous()
bw = torch.full((B,), 0.3, device=dev)
out, tmp = torch.empty_like(X), torch.empty_like(X)
def run():
    _lib.call("sed_ms_shift", _lib.ptr(X), _lib.ptr(bw), B, N, d, iters, 0, prec, _lib.ptr(out), _lib.ptr(tmp), _lib.stream())
run(); torch.cuda.synchronize()
samples, stop = [], False
def sampler():
    while not stop:
        r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,clocks_throttle_reasons.active",
                            "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True)
        samples.append(r.stdout.strip())
        time.sleep(0.2)
th = threading.Thread(target=sampler); th.start()
t0 = time.time(); n = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < secs:
    run(); n += 1
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
stop = True; th.join()
ms = e0.elapsed_time(e1) / n
print(f"mode {prec}: {ms / iters * 1e3:.1f} us per iteration over {n} x {iters} iterations")
for s in samples[2:]: print("  sm MHz, max MHz, W, W limit, reasons:", s)
