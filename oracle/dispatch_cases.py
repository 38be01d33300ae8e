"""TEST INFRASTRUCTURE ONLY.  Seeded inputs of the a18 dispatcher fixtures (tests/golden/dispatch.npz), shared by
oracle/make_golden_dispatch.py (which records the reference's outputs on them) and the tests (which replay them):
``data`` / ``weights`` as Evaluation.residual_eval_mode builds them (Fitting_patches_and_edges/residual_utils.py:210-321).
No reference code is needed to rebuild them."""
import importlib.util
import os

import numpy as np
import torch

import oracle as O

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("sednet_synth", os.path.join(_ROOT, "sed-net_b200", "synth.py"))
synth = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(synth)


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


SPLINE_OPEN = 2


def make_case(seed, n=4000, noise=0.0):
    """Cloud + cluster ids + per-point predicted types: the synthetic patches, one 15-point cluster (dropped by the
    20-point minimum, :974) and one 60-point open-spline cluster (dropped by the 100-point minimum, :1027)."""
    p, nrm, lab, typ, _ = synth.make_cloud(seed, n, n_patches=8)
    if noise > 0:
        p = (p + noise * np.random.default_rng(seed + 3).normal(size=p.shape)).astype(np.float32)
    lab, typ = lab.copy(), typ.copy()
    S = int(lab.max()) + 1
    big = np.argsort(-np.bincount(lab))[:2]
    i0 = np.where(lab == big[0])[0][:15]
    i1 = np.where(lab == big[1])[0][:60]
    lab[i0] = S
    lab[i1] = S + 1
    typ[i1] = SPLINE_OPEN
    # a few mislabelled point types inside the segments: the dispatcher takes the per-segment mode
    flip = np.random.default_rng(seed).choice(n, n // 50, replace=False)
    typ[flip] = np.random.default_rng(seed + 1).choice([1, 3, 4, 5], flip.shape[0])
    typ[i1] = SPLINE_OPEN
    return p, nrm, lab, typ


def build_data(p, nrm, cluster_ids, pred_types):
    """residual_utils.py:245-285 with gt labels == cluster ids (identity matching): data entries
    [points, normals, mode type, gt points, pred mask, (index, id)]."""
    P, Nn = t(p), t(nrm)
    data = []
    for index, i in enumerate(np.unique(cluster_ids)):
        m = cluster_ids == i
        l = int(np.bincount(pred_types[m]).argmax())          # stats.mode(...)[0]: the smallest of the most frequent
        data.append([P[m], Nn[m], l, P[m], m, (index, int(i))])
    return data


def eval_weights(cluster_ids, bw):
    """residual_utils.py:237-239,300-308: one-hot -> weights_normalize -> argmax one-hot, (N, C)."""
    C = np.unique(cluster_ids).shape[0]
    w0 = O.to_one_hot(cluster_ids, C).numpy().T
    w = O.weights_normalize(t(w0.astype(np.float32)), float(bw))
    w = torch.transpose(w, 1, 0)
    return O.to_one_hot(torch.max(w, 1)[1].numpy(), w.shape[1])


def train_weights(seed, lab):
    """Soft membership weights of the training-mode case (N, C)."""
    C = np.unique(lab).shape[0]
    soft = np.random.default_rng(seed + 7).uniform(0.0, 0.1, (lab.shape[0], C)).astype(np.float32)
    soft[np.arange(lab.shape[0]), np.searchsorted(np.unique(lab), lab)] += 0.9
    return t(soft)


def build_case(seed, mode_eval, noise):
    """-> (p, nrm, lab, typ, data, W) of one fixture case."""
    p, nrm, lab, typ = make_case(seed, noise=noise)
    data = build_data(p, nrm, lab, typ)
    if mode_eval:
        W = eval_weights(lab, 0.1)
    else:
        # training mode (src/primitive_forward.py:945-963): every entry carries the whole cloud, soft weights
        W = train_weights(seed, lab)
        data = [[t(p), t(nrm), d[2], d[3], d[4], d[5]] for d in data if d[2] != SPLINE_OPEN]
    return p, nrm, lab, typ, data, W
