"""compute_entropy and hpnet_process with the reference's signatures (src/smooth_normal_matrix.py:95-154, :157-233).

compute_entropy runs on the pairwise kernel of libsednet_b200.so (no CHUNK x CHUNK x K temporaries).  hpnet_process is
implemented for the branch the reference takes when the spectral vectors of the shape are cached
(src/normal_smooth_cache/Us_{id}_{sigma}_{knn}.pt and WUs_..., :181-189) or are handed in through the extra keyword
arguments `v`, `ent`; the branch that builds them (:190-196) calls torch.lobpcg for 10 iterations from a random start on
a nearly degenerate spectrum -- its result is not reproducible run to run in the reference itself, so it is not
re-implemented here and raises NotImplementedError."""
import os

import torch

from . import _lib


def compute_entropy(features, CHUNK=2000):
    """features (1,N,K) -> 0-d tensor E (src/smooth_normal_matrix.py:95-154)."""
    assert features.shape[0] == 1
    feat = _lib.require_cuda(features[0], name="features")
    N, K = feat.shape
    out = torch.empty(3, dtype=torch.float32, device=feat.device)
    _lib.call("sed_compute_entropy", _lib.ptr(feat), N, K, int(CHUNK), _lib.ptr(out), _lib.stream())
    return out[0]


def hpnet_process(affinity_feat, inputs_xyz, normals, id=None, types=None, edges=None, normal_smooth_w=0.5, CHUNK=2000,
                  gpu='cuda:0', drop_rest_idx=None, v=None, ent=None):
    """src/smooth_normal_matrix.py:157-233: affinity_feat (B,N,K) without L2, types (B,N,6) log-probabilities, edges
    (B,N,2) logits -> (B,N,K+12[+6[+2]]) weighted concatenation [features, spectral vectors, type/edge probabilities]."""
    weight_ent, parts = [], []
    weight_ent.append(1.7 - float(compute_entropy(affinity_feat, CHUNK=CHUNK)))                    # :172-174
    parts.append(affinity_feat)
    edge_topk, normal_sigma, edge_knn = 12, 0.1, 50                                                # :179-182
    fn = "src/normal_smooth_cache/Us_{}_{}_{}.pt".format(id, normal_sigma, edge_knn)
    fn_ent = "src/normal_smooth_cache/WUs_{}_{}_{}.pt".format(id, normal_sigma, edge_knn)
    if v is None:
        if id is not None and os.path.exists(fn) and os.path.exists(fn_ent):                       # :186-189
            v = torch.load(fn).to(affinity_feat.device)
            ent = torch.load(fn_ent)
        else:
            raise NotImplementedError(
                "hpnet_process: no cached spectral vectors for this shape.  The reference would build them with "
                "torch.lobpcg(niter=10) from a random start (src/smooth_normal_matrix.py:190-196), which is not "
                "reproducible; pass v= (B,N,12) and ent=, or provide the cache files " + fn)
    if ent is None:
        ent = compute_entropy(v, CHUNK=CHUNK)                                                      # :195
    if drop_rest_idx is not None:
        v = v[:, drop_rest_idx, :]
    weight_ent.append(normal_smooth_w - float(ent))                                                # :202-205
    parts.append(v)
    if types is not None:                                                                          # :208-216
        types = torch.exp(types)
        if edges is not None:
            types = torch.cat((types, torch.softmax(edges, dim=-1)), dim=-1)
        weight_ent.append(0.25 - float(compute_entropy(types, CHUNK=CHUNK)))
        parts.append(types)
    return torch.cat([p * w for p, w in zip(parts, weight_ent)], dim=-1)                           # :221-233
