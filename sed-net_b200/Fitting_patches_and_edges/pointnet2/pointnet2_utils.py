"""three_nn with the signature of the reference's pointnet2 extension
(Fitting_patches_and_edges/pointnet2/pointnet2_utils.py:118-147, kernel _ext_src/src/interpolate_gpu.cu:14-66)."""
import torch

from ...src import _lib


def three_nn(unknown, known):
    """unknown (B,n,3), known (B,m,3) -> (dist (B,n,3) L2 distances ascending, idx (B,n,3) int32)."""
    unknown = _lib.require_cuda(unknown, name="unknown")
    known = _lib.require_cuda(known, name="known")
    B, n, _ = unknown.shape
    m = known.shape[1]
    dist2 = torch.empty((B, n, 3), dtype=torch.float32, device=unknown.device)
    idx = torch.empty((B, n, 3), dtype=torch.int32, device=unknown.device)
    _lib.call("sed_three_nn", _lib.ptr(unknown), _lib.ptr(known), B, n, m, _lib.ptr(dist2), _lib.ptr(idx), _lib.stream())
    return torch.sqrt(dist2), idx
