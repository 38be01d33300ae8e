"""chamfer_distance with the reference's signature (src/utils.py:273-323); the row / column minima of the pairwise
distance matrix come from the nearest-point kernel of libsednet_b200.so (the n x m matrix is never formed)."""
import numpy as np
import torch

from . import _lib
from .guard import guard_sqrt


def _mins(pred, gt):
    if isinstance(pred, np.ndarray):
        pred = torch.from_numpy(pred.astype(np.float32)).cuda()
    if isinstance(gt, np.ndarray):
        gt = torch.from_numpy(gt.astype(np.float32)).cuda()
    pred = _lib.require_cuda(pred, name="pred")
    gt = _lib.require_cuda(gt, name="gt")
    B, n, _ = pred.shape
    m = gt.shape[1]
    mp = torch.empty((B, n), dtype=torch.float32, device=pred.device)
    mg = torch.empty((B, m), dtype=torch.float32, device=pred.device)
    _lib.call("sed_chamfer_min", _lib.ptr(pred), _lib.ptr(gt), B, n, m, _lib.ptr(mp), _lib.ptr(mg), _lib.stream())
    return mp, mg


def chamfer_distance(pred, gt, sqrt=False):
    """src/utils.py:273-296: pred (B,N,3), gt (B,M,3) -> scalar mean over the batch of
    (mean_n min_m d + mean_m min_n d) / 2, d the squared (or guarded-root) distance."""
    mp, mg = _mins(pred, gt)
    if sqrt:
        mp, mg = guard_sqrt(mp), guard_sqrt(mg)           # min commutes with the monotone guarded root
    cd = torch.mean(mp, 1) + torch.mean(mg, 1)
    return torch.mean(cd) / 2.0


def chamfer_distance_one_side(pred, gt, side=1):
    """src/utils.py:299-323: side 0 = mean over pred of the distance to gt, side 1 = mean over gt of the distance to pred."""
    mp, mg = _mins(pred, gt)
    return torch.mean(torch.mean(mp if side == 0 else mg, 1))
