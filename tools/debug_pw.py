"""Debug: forward outputs for several N, saved for comparison between SEDNET_B200_PW settings."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from sednet_b200 import synth
from sednet_b200.src import SEDNet
tag = sys.argv[1]
dev = torch.device("cuda")
sd = synth.make_state_dict(1, randomize_gn=True)
out = {}
for N, k, B in ((600, 16, 1), (640, 16, 1), (1024, 16, 2), (1500, 32, 2)):
    m = SEDNet.SEDNet(embedding=True, emb_size=128, primitives=True, num_primitives=6, mode=5, num_channels=6,
                      combine_label_prim=True, edge_module=True, late_fusion=True, nn_nb=k)
    m.load_state_dict({kk: torch.from_numpy(v) for kk, v in sd.items()})
    m = m.to(dev).eval()
    pts, nrm, _, _ = synth.make_batch(B, N, seed0=77)
    x = torch.from_numpy(np.concatenate([pts, nrm], 2).transpose(0, 2, 1).copy()).to(dev)
    x4, feats = m.encode(x)
    o = m(x)
    out[N] = dict(x4=x4.cpu(), feats=feats.cpu(), emb=o[0].cpu(), logp=o[1].cpu(), edges=o[3].cpu())
torch.save(out, os.path.join(ROOT, "gpurun_out", f"dbg_{tag}.pt"))
