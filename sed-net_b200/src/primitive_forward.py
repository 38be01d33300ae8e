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


_SPLINE_TYPES = (0, 2, 6, 7, 8, 9)


def parameters_from_row(prim, row):
    """One (8,) row of sed_fit_segments -> the list FittingModule.forward_pass_* stores in ``fitting.parameters[ids]``
    (reference src/fitting_optimization.py:167,195,215,237; shapes as the reference's)."""
    if prim == PLANE:
        return ["plane", row[0:3].reshape(3, 1), row[3]]
    if prim == CONE:
        return ["cone", row[0:3].reshape(1, 3), row[3:6].reshape(3, 1), row[6]]
    if prim == CYLINDER:
        return ["cylinder", row[0:3].reshape(3, 1), row[3:6].reshape(1, 3), row[6]]
    if prim == SPHERE:
        return ["sphere", row[0:3].reshape(1, 3), row[3]]
    raise ValueError(prim)


def fit_one_shape_torch(data, fitter, weights, bw, eval=False, sample_points=False, if_optimize=False,
                        if_visualize=False):
    """Reference src/primitive_forward.py:929-1051 for the analytic primitive types: same arguments, same
    ``fitter.fitting.parameters`` afterwards (None for patches under 20 points or spline patches under 100), same
    ``(gt_points, reconstructed_shape)`` return.  Where the reference fits segment after segment in Python (one SVD /
    QR / host condition number each), every segment of the shape goes through ONE launch of the batched fit kernel
    (sed_fit_segments): the per-segment rows are concatenated into one cloud whose labels are the segment numbers."""
    if sample_points:
        raise NotImplementedError("sample_points=True (mesh sampling of the fitted primitives) is outside the hot path")
    fitter.fitting.parameters = {}
    gt_points, reconstructed_shape = {}, []
    P, Nn, Wt, seg_types, seg_ids = [], [], [], [], []
    for d in data:
        points, normals, labels, gpoints, segment_indices, part_index = d
        part_index, label_index = part_index
        labels = int(labels)
        points = _lib.require_cuda(points, name="points")
        normals = _lib.require_cuda(normals, name="normals")
        if not eval:                                          # :945-963: every second point (twice for primitives)
            weight = weights[:, part_index:part_index + 1] + EPS
            keep = torch.arange(0, points.shape[0], 2, device=points.device)
            points, normals, weight = points[keep], normals[keep], weight[keep]
            if labels not in _SPLINE_TYPES:
                keep = torch.arange(0, points.shape[0], 2, device=points.device)
                points, normals, weight = points[keep], normals[keep], weight[keep]
        else:
            idx = torch.as_tensor(segment_indices, device=weights.device)
            weight = weights[idx, part_index:part_index + 1] + EPS                     # :954
        reconstructed_shape.append(None)
        if points.shape[0] < 20 or (labels in _SPLINE_TYPES and points.shape[0] < 100):    # :974, :985, :1027
            gt_points[label_index] = None
            fitter.fitting.parameters[label_index] = None
            continue
        if labels in _SPLINE_TYPES:
            raise NotImplementedError("SplineNet patches are outside the analytic-primitive hot path")
        if labels not in (PLANE, CONE, CYLINDER, SPHERE):
            raise ValueError(f"unknown primitive type id {labels}")
        P.append(points); Nn.append(normals); Wt.append(weight.to(points.device).reshape(-1))
        seg_types.append(labels); seg_ids.append(label_index)
        gt_points[label_index] = gpoints
    if P:
        dev = P[0].device
        sizes = torch.tensor([p.shape[0] for p in P], device=dev)
        lab = torch.repeat_interleave(torch.arange(len(P), device=dev), sizes).reshape(1, -1)
        st = torch.tensor([seg_types], dtype=torch.int32, device=dev)
        params, _ = fit_segments_batched(torch.cat(P)[None], torch.cat(Nn)[None], lab, st, weights=torch.cat(Wt)[None],
                                         min_pts=0)
        for s, (prim, label_index) in enumerate(zip(seg_types, seg_ids)):
            fitter.fitting.parameters[label_index] = parameters_from_row(prim, params[0, s])
    return gt_points, reconstructed_shape
