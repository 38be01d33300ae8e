"""Graph ops of the DGCNN encoder -- same names and signatures as reference src/PointNet.py:62-208, executed by
the fused distance + top-k kernels of libsednet_b200.so (the N x N distance matrix is never materialised)."""
import numpy as np
import torch

from . import _lib


def _subsample_columns(idx, k1, k2, normal):
    """Column selection applied after topk(k2): src/PointNet.py:64-71 (identity when k1 == k2)."""
    if not normal:
        cols = np.arange(0, k2, k2 // k1)
    else:
        y = np.linspace(0.0, 3.0, k2)
        p_n = np.exp(-y ** 2 / 2)
        cols = np.random.choice(np.arange(0, k2), k1, p=p_n / p_n.sum(), replace=False)
    if len(cols) == k2 and (cols == np.arange(k2)).all():
        return idx
    return idx[:, :, torch.as_tensor(cols, device=idx.device)]


def knn(x, k1, k2, normal=False):
    """src/PointNet.py:62-87: x (B,C,N) -> idx (B,N,k1) int64, nearest first under squared L2."""
    x = _lib.require_cuda(x, name="x")
    B, Cc, N = x.shape
    idx = torch.empty((B, N, k2), dtype=torch.int64, device=x.device)
    _lib.call("sed_knn_l2", _lib.ptr(x), B, Cc, N, k2, _lib.ptr(idx), 1, _lib.stream())
    return _subsample_columns(idx, k1, k2, normal)


def knn_points_normals(x, k1, k2, normal_metric_W=1., normal=False):
    """src/PointNet.py:90-137: x (B,6,N); metric p_dist * (1 + W * n_dist)."""
    x = _lib.require_cuda(x, name="x")
    B, Cc, N = x.shape
    if Cc != 6:
        raise RuntimeError("knn_points_normals expects (B, 6, N): xyz + normals")
    idx = torch.empty((B, N, k2), dtype=torch.int64, device=x.device)
    _lib.call("sed_knn_pn", _lib.ptr(x), B, N, k2, float(normal_metric_W), _lib.ptr(idx), 1, _lib.stream())
    return _subsample_columns(idx, k1, k2, normal)


def _gather(x, idx):
    B, Cc, N = x.shape
    idx = _lib.require_cuda(idx, torch.int64, "idx")
    k = idx.shape[2]
    out = torch.empty((B, 2 * Cc, N, k), dtype=torch.float32, device=x.device)
    _lib.call("sed_graph_feature", _lib.ptr(x), _lib.ptr(idx), B, Cc, N, k, _lib.ptr(out), _lib.stream())
    return out


def get_graph_feature(x, k1=20, k2=20, idx=None, Norm_sample=False):
    """src/PointNet.py:140-171: (B,C,N) -> cat([x_j - x_i, x_i]) as (B,2C,N,k1)."""
    x = _lib.require_cuda(x, name="x")
    x = x.view(x.shape[0], -1, x.shape[2])
    if idx is None:
        idx = knn(x, k1=k1, k2=k2, normal=Norm_sample)
    return _gather(x, idx)


def get_graph_feature_with_normals(x, k1=20, k2=20, idx=None, normal_metric_W=1., Norm_sample=False):
    """src/PointNet.py:174-208: as above with the point x normal metric for the neighbour search."""
    x = _lib.require_cuda(x, name="x")
    x = x.view(x.shape[0], -1, x.shape[2])
    if idx is None:
        idx = knn_points_normals(x, k1=k1, k2=k2, normal_metric_W=normal_metric_W, normal=Norm_sample)
    return _gather(x, idx)
