"""Debug: encoder forward at a given batch size (python tools/debug_enc.py B [N])."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from sednet_b200 import synth
from sednet_b200.src import SEDNet
B = int(sys.argv[1]); N = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
dev = torch.device("cuda")
sd = synth.make_state_dict(1, randomize_gn=True)
m = SEDNet.SEDNet(embedding=True, emb_size=128, primitives=True, num_primitives=6, mode=5, num_channels=6,
                  combine_label_prim=True, edge_module=True, late_fusion=True, nn_nb=64)
m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
m = m.to(dev).eval()
pts, nrm, _, _ = synth.make_batch(B, N, seed0=900)
x = torch.from_numpy(np.concatenate([pts, nrm], 2).transpose(0, 2, 1).copy()).to(dev)
x4, feats = m.encode(x)
torch.cuda.synchronize()
print("encode ok", float(x4.abs().max()), float(feats.abs().max()))
out = m(x)
torch.cuda.synchronize()
print("forward ok", float(out[0].abs().max()))
