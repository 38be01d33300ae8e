"""Deviation of every mean-shift precision mode from an FP64 evaluation of the same iterations, as a function of the iteration
count, on one cloud of the randomised sweep (default: seed 24, the worst case of tools/sweep_parity.py) -- the numbers of the
table in DESIGN.md section 5.   python tools/debug_seed24.py [seed]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import oracle as O
from sednet_b200 import synth
from sednet_b200.src.mean_shift import MeanShift
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rng = np.random.default_rng(seed)
N = int(rng.integers(600, 3200)); npatch = int(rng.integers(3, 17)); sigma = float(rng.choice([0.005, 0.01, 0.02, 0.04])); iters = int(rng.choice([10, 25, 50]))
pts, nrm, lab, typ, _ = synth.make_cloud(9000 + seed, N, n_patches=npatch, min_pts=30)
X = torch.from_numpy(synth.make_embedding(lab, 128, sigma, 100 + seed))
dev = torch.device("cuda")
with torch.no_grad():
    bw = torch.clamp(O.ms_bandwidth(X, 10000, 0.015), min=0.003)
    for it in (5, 10, 15, 20, 25, 35, 50):
        o32 = O.ms_shift(X, bw, it)
        o64 = O.ms_shift(X.double(), bw.double(), it)
        g64 = O.ms_shift(X.double().to(dev), bw.double().to(dev), it).cpu()
        res = {}
        for prec in (0, 1, 3, 4):
            out, _ = MeanShift(prec_mode=prec).mean_shift_(X.to(dev), b=bw, iterations=it)
            res[prec] = out.cpu()
        print(f"it {it}: |cpu32-fp64| {float((o32.double()-o64).abs().max()):.2e}  |gpu64-cpu64| {float((g64-o64).abs().max()):.1e}  "
              + "  ".join(f"|m{p}-fp64| {float((res[p].double()-o64).abs().max()):.2e} |m{p}-cpu32| {float((res[p]-o32).abs().max()):.2e}" for p in (0,1,3,4)))
