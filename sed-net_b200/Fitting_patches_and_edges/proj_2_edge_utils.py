"""Instance adjacency maps of stage 2 (reference Fitting_patches_and_edges/proj_2_edge_utils.py:45-110)."""
import numpy as np
import torch

from ..src import _lib
from .pointnet2.pointnet2_utils import three_nn


def torch_type(x):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    return x.cuda()


def _idx3(points):
    return three_nn(points.unsqueeze(0), points.unsqueeze(0))[1][0].contiguous()


def get_edges_between_insts(points, insts, strict=True):
    """proj_2_edge_utils.py:45-60 -> bool (N,): the first (strict: and second) non-self neighbour of the point belongs to
    another instance."""
    points = _lib.require_cuda(torch_type(points), name="points")[:, :3].contiguous()
    insts = _lib.require_cuda(torch_type(insts), torch.int64, "insts")
    n = points.shape[0]
    out = torch.empty(n, dtype=torch.uint8, device=points.device)
    idx3 = _idx3(points)
    _lib.call("sed_inst_edges", _lib.ptr(idx3), _lib.ptr(insts), n, 1 if strict else 0, _lib.ptr(out), _lib.stream())
    return out.bool()


def face_face_inter_map(points, insts, primitive_ids, nn_num_thresh=3):
    """proj_2_edge_utils.py:63-110 -> bool (30,30) adjacency of instance ids."""
    points = _lib.require_cuda(torch_type(points), name="points")
    insts = _lib.require_cuda(torch_type(insts), torch.int64, "insts")
    ids = _lib.require_cuda(torch_type(primitive_ids) if not isinstance(primitive_ids, torch.Tensor) else primitive_ids.cuda(),
                            torch.int64, "primitive_ids")
    n = points.shape[0]
    mat = torch.empty((30, 30), dtype=torch.uint8, device=points.device)
    idx3 = _idx3(points)
    _lib.call("sed_face_face_map", _lib.ptr(points), _lib.ptr(insts), _lib.ptr(idx3), _lib.ptr(ids), int(ids.numel()),
              n, int(nn_num_thresh), _lib.ptr(mat), _lib.stream())
    return mat.bool().cpu()
