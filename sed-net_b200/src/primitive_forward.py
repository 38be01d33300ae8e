"""Fit with the reference's method names (reference src/primitive_forward.py:422-429, :712-847), each call running
the batched fit kernel of libsednet_b200.so on one segment; ``fit_segments_batched`` fits every segment of every
cloud in one launch (what fit_one_shape_torch :929-1051 does in a Python loop)."""
import numpy as np
import torch

from . import _lib
from .fitting_utils import LeastSquares

PLANE, CONE, CYLINDER, SPHERE = 1, 3, 4, 5
EPS = float(np.finfo(np.float32).eps)


def _fit_one(prim, points, normals, weights):
    points = _lib.require_cuda(points, name="points")
    n = points.shape[0]
    dev = points.device
    normals = _lib.require_cuda(normals, name="normals") if normals is not None else None
    w = _lib.require_cuda(weights, name="weights").reshape(-1) if weights is not None else None
    seg_type = torch.tensor([prim], dtype=torch.int32, device=dev)
    params = torch.empty(8, dtype=torch.float32, device=dev)
    status = torch.empty(1, dtype=torch.int32, device=dev)
    _lib.call("sed_fit_segments", _lib.ptr(points), _lib.ptr(normals), _lib.ptr(w), _lib.ptr(None), _lib.ptr(seg_type),
              1, n, 1, 0, _lib.ptr(params), _lib.ptr(status), _lib.stream())
    return params, status


class Fit:
    def __init__(self):
        LS = LeastSquares()
        self.lstsq = LS.lstsq
        self.parameters = {}

    def fit_plane_torch(self, points, normals, weights, ids=0, show_warning=False):
        """src/primitive_forward.py:712-733 -> a (1,3), d ().  The sign of (a, d) is canonical (largest |a_i| > 0);
        the reference's follows LAPACK and is arbitrary."""
        p, _ = _fit_one(PLANE, points, None, weights)
        return p[0:3].reshape(1, 3), p[3]

    def fit_sphere_torch(self, points, normals, weights, ids=0, show_warning=False):
        """src/primitive_forward.py:750-773 -> center (1,3), radius ()."""
        p, _ = _fit_one(SPHERE, points, None, weights)
        return p[0:3].reshape(1, 3), p[3]

    def fit_cylinder_torch(self, points, normals, weights, ids=0, show_warning=False):
        """src/primitive_forward.py:788-810 -> a (3,1), center (1,3), radius ()."""
        p, _ = _fit_one(CYLINDER, points, normals, weights)
        return p[0:3].reshape(3, 1), p[3:6].reshape(1, 3), p[6]

    def fit_cone_torch(self, points, normals, weights, ids=0, show_warning=False):
        """src/primitive_forward.py:812-847 -> apex (3,1), axis (1,3), theta ()."""
        p, _ = _fit_one(CONE, points, normals, weights)
        return p[0:3].reshape(3, 1), p[3:6].reshape(1, 3), p[6]


def fit_segments_batched(points, normals, labels, seg_type, weights=None, min_pts=20):
    """All segments of a batch in one launch.  points, normals (B,N,3); labels (B,N) int64; seg_type (B,S) int32;
    returns params (B,S,8) and status (B,S) (0 fitted, 1 skipped, 2 regularised branch, 3 degenerate cone)."""
    points = _lib.require_cuda(points, name="points")
    normals = _lib.require_cuda(normals, name="normals")
    labels = _lib.require_cuda(labels, torch.int64, "labels")
    seg_type = _lib.require_cuda(seg_type, torch.int32, "seg_type")
    w = _lib.require_cuda(weights, name="weights") if weights is not None else None
    B, N, _ = points.shape
    S = seg_type.shape[1]
    params = torch.empty((B, S, 8), dtype=torch.float32, device=points.device)
    status = torch.empty((B, S), dtype=torch.int32, device=points.device)
    _lib.call("sed_fit_segments", _lib.ptr(points), _lib.ptr(normals), _lib.ptr(w), _lib.ptr(labels),
              _lib.ptr(seg_type), B, N, S, int(min_pts), _lib.ptr(params), _lib.ptr(status), _lib.stream())
    return params, status
