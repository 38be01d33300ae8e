"""Shared helpers of the parity tests."""
import numpy as np
import torch

import oracle as O
from sednet_b200 import synth


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def cloud_input(seed, n, **kw):
    p, nrm, lab, typ, patches = synth.make_cloud(seed, n, **kw)
    x = np.concatenate([p, nrm], 1).T[None].copy()
    return p, nrm, lab, typ, x


def knn_set_agreement(a, b):
    """Fraction of rows whose neighbour SETS agree, and fraction of individual entries shared."""
    a, b = np.asarray(a), np.asarray(b)
    a2, b2 = np.sort(a.reshape(-1, a.shape[-1]), 1), np.sort(b.reshape(-1, b.shape[-1]), 1)
    rows = (a2 == b2).all(1).mean()
    shared = np.mean([len(np.intersect1d(x, y)) / len(x) for x, y in zip(a2, b2)])
    return float(rows), float(shared)


def canon(labels):
    return O.canonical_labels(np.asarray(labels))


def sign_align(a, ref):
    """Resolve the arbitrary sign of a singular vector: returns a or -a, whichever is closer to ref."""
    a, ref = np.asarray(a, np.float64).ravel(), np.asarray(ref, np.float64).ravel()
    return a if float(a @ ref) >= 0 else -a


def rel_err(a, ref):
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    return float(np.max(np.abs(a - ref)) / max(np.max(np.abs(ref)), 1e-12))
