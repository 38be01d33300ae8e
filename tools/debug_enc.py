"""Debug: encoder + full forward at a given batch size, optionally repeated with a bit-equality check
(python tools/debug_enc.py B [N] [reps])."""
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
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 0      # repeat and require identical bits (synchronisation bugs show as differences)
first = [o.clone() for o in (out[0], out[1], out[3])]
for r in range(reps):
    o = m(x)
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(first, (o[0], o[1], o[3]))), f"rep {r}: outputs differ"
if reps:
    print(f"{reps} repetitions bit-identical")
