"""Stage-2 ``Fit`` with the reference's method names and signatures (Fitting_patches_and_edges/primitive_forward_v2.py:
716-891); ``fit_segments_batched_v2`` fits every segment of every cloud in one launch (what fit_one_shape_torch
:929-1051 does in a Python loop).  Type ids of the stage-2 dispatcher (:953-976): 1 plane, 3 cone, 2 cylinder, 4 sphere."""
import numpy as np
import torch

from ..src import _lib
from ..src.fitting_utils import LeastSquares

PLANE, CONE, CYLINDER, SPHERE = 1, 3, 4, 5             # ids of the C ABI (stage-1 convention)
STAGE2_TO_ABI = {1: PLANE, 3: CONE, 2: CYLINDER, 4: SPHERE}
EPS = float(np.finfo(np.float32).eps)


def _fit_one(prim, points, normals, weights, ratio=0.0):
    points = _lib.require_cuda(points, name="points")
    n = points.shape[0]
    dev = points.device
    normals = _lib.require_cuda(normals, name="normals") if normals is not None else None
    w = _lib.require_cuda(weights, name="weights").reshape(-1) if weights is not None else None
    seg_type = torch.tensor([prim], dtype=torch.int32, device=dev)
    params = torch.empty(8, dtype=torch.float32, device=dev)
    status = torch.empty(1, dtype=torch.int32, device=dev)
    _lib.call("sed_fit_segments_v2", _lib.ptr(points), _lib.ptr(normals), _lib.ptr(w), _lib.ptr(None), _lib.ptr(seg_type),
              1, n, 1, 0, float(ratio), _lib.ptr(params), _lib.ptr(status), _lib.stream())
    return params


class Fit:
    def __init__(self):
        self.lstsq = LeastSquares().lstsq
        self.parameters = {}

    def fit_plane_torch(self, points, normals, weights, ids=0, show_warning=False, nofilter=False, filter_ratio=0.5):
        """:716-752 -> a (1,3), d (); fit on the int(n * filter_ratio) points nearest to the mean unless nofilter.
        The sign of (a, d) is canonical (largest |a_i| > 0); the reference's follows LAPACK and is arbitrary."""
        p = _fit_one(PLANE, points, None, weights, 0.0 if nofilter else filter_ratio)
        return p[0:3].reshape(1, 3), p[3]

    def fit_sphere_torch(self, points, normals, weights, ids=0, show_warning=False):
        """:772-796 -> center (1,3), radius ()."""
        p = _fit_one(SPHERE, points, None, weights)
        return p[0:3].reshape(1, 3), p[3]

    def fit_cylinder_torch(self, points, normals, weights, ids=0, show_warning=False):
        """:810-849 -> a (3,1), center (3,), radius ()."""
        p = _fit_one(CYLINDER, points, normals, weights)
        return p[0:3].reshape(3, 1), p[3:6], p[6]

    def fit_cone_torch(self, points, normals, weights, ids=0, show_warning=False):
        """:851-891 -> apex (3,1), axis (1,3), theta ()."""
        p = _fit_one(CONE, points, normals, weights)
        return p[0:3].reshape(3, 1), p[3:6].reshape(1, 3), p[6]


def fit_segments_batched_v2(points, normals, labels, seg_type, weights=None, min_pts=20, plane_filter_ratio=0.25,
                            stage2_type_ids=False):
    """All segments of a batch in one launch.  points, normals (B,N,3); labels (B,N) int64; seg_type (B,S) int32 in the
    C-ABI ids (or the stage-2 dispatcher's with stage2_type_ids); returns params (B,S,8), status (B,S) (0 fitted,
    1 skipped).  plane_filter_ratio: MyFittingModule.sample_ratio (0.25 in primitive_forward_v2.py:1086)."""
    points = _lib.require_cuda(points, name="points")
    normals = _lib.require_cuda(normals, name="normals")
    labels = _lib.require_cuda(labels, torch.int64, "labels")
    seg_type = _lib.require_cuda(seg_type, torch.int32, "seg_type")
    if stage2_type_ids:
        lut = torch.zeros(16, dtype=torch.int32, device=seg_type.device)
        for k, v in STAGE2_TO_ABI.items():
            lut[k] = v
        seg_type = lut[seg_type.clamp(0, 15).long()].contiguous()
    w = _lib.require_cuda(weights, name="weights") if weights is not None else None
    B, N, _ = points.shape
    S = seg_type.shape[1]
    params = torch.empty((B, S, 8), dtype=torch.float32, device=points.device)
    status = torch.empty((B, S), dtype=torch.int32, device=points.device)
    _lib.call("sed_fit_segments_v2", _lib.ptr(points), _lib.ptr(normals), _lib.ptr(w), _lib.ptr(labels),
              _lib.ptr(seg_type), B, N, S, int(min_pts), float(plane_filter_ratio), _lib.ptr(params), _lib.ptr(status),
              _lib.stream())
    return params, status
