"""Runs the fused SEDNet forward alone at the bench shape (for ncu captures).  python tools/prof_fwd.py [B] [reps]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from sednet_b200 import synth
from sednet_b200.src import SEDNet
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda")
sd = synth.make_state_dict(1, randomize_gn=True)
m = SEDNet.SEDNet(embedding=True, emb_size=128, primitives=True, num_primitives=6, mode=5, num_channels=6,
                  combine_label_prim=True, edge_module=True, late_fusion=True, nn_nb=64)
m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
m = m.to(dev).eval()
pts, nrm, _, _ = synth.make_batch(B, 10000, seed0=1234)
x = torch.from_numpy(np.concatenate([pts, nrm], 2).transpose(0, 2, 1).copy()).to(dev)
for _ in range(reps):
    out = m(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = m(x); e1.record(); torch.cuda.synchronize()
print(f"forward B{B}: {e0.elapsed_time(e1):.3f} ms")
