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


def cylinder_fp64(p, n, w, eps32=float(np.finfo(np.float32).eps)):
    """FP64 evaluation of the reference's cylinder fit formulas (src/primitive_forward.py:788-810 + :750-773 +
    src/fitting_utils.py:36-85) with the FP32 rank tolerances of torch.matrix_rank: what the formulas give without
    the rounding noise of the FP32 explicit-inverse solve (cond ~ 1e6 in the regularised branch).  Returns
    axis (3,), centre (3,), radius."""
    p, n, w = (np.asarray(v, np.float64) for v in (p, n, w))
    w = w.reshape(-1, 1)
    _, V = np.linalg.eigh((w * n).T @ (w * n))
    ax = V[:, 0] / (np.linalg.norm(V[:, 0]) + eps32)
    v = p - (p @ ax)[:, None] * ax
    ws = w.sum() + eps32
    m = (w * v).sum(0) / ws
    q = (v * v).sum(1)
    tq = (w[:, 0] * q).sum() / ws
    A = 2 * w * (m - v)
    Y = w[:, 0] * (w[:, 0] * q - tq)
    AtA = A.T @ A
    mu = np.linalg.eigvalsh(AtA)
    smax, smin = np.sqrt(max(mu[2], 0)), np.sqrt(max(mu[0], 0))
    lam = 0.0
    if not smin > smax * max(A.shape[0], 3) * eps32:
        lam = 1e-6
        for _ in range(7):
            if (mu[0] + lam) > (mu[2] + lam) * 3 * eps32:
                break
            lam *= 10
    c = -np.linalg.solve(AtA + lam * np.eye(3), A.T @ Y)
    r2 = (w[:, 0] * ((v - c) ** 2).sum(1)).sum() / ws
    return ax, c, float(np.sqrt(max(r2, 1e-3)))


def flat_params(v):
    """["name", tensors...] of fitting.parameters -> flat float64 vector."""
    return np.concatenate([np.asarray(x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else x, np.float64).ravel()
                           for x in v[1:]])


def comparable_params(name, got, ref):
    """Resolve the conventions the reference leaves open before comparing a fit with a recorded one: the sign of a plane's
    (a, d) and of a cylinder's axis follow LAPACK's singular vector; a cylinder's centre is only determined orthogonally
    to its axis (SURVEY.md 7.3-4,5).  Returns (got', ref') flat vectors."""
    got, ref = np.asarray(got, np.float64).ravel().copy(), np.asarray(ref, np.float64).ravel().copy()
    if name == "plane":
        if got[:3] @ ref[:3] < 0:
            got = -got
    elif name == "cylinder":
        if got[:3] @ ref[:3] < 0:
            got[:3] = -got[:3]
        ax = ref[:3]
        got[3:6] -= (got[3:6] @ ax) * ax
        ref[3:6] -= (ref[3:6] @ ax) * ax
    return got, ref
