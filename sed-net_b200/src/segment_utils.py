"""Hot-path-adjacent helpers of reference src/segment_utils.py (to_one_hot :536-545)."""
import torch

from . import _lib


def to_one_hot(target, maxx=50, device_id=0):
    """src/segment_utils.py:536-545: labels (N,) -> one-hot (N, maxx) f32 on the GPU."""
    if not isinstance(target, torch.Tensor):
        target = torch.as_tensor(target)
    if not target.is_cuda:
        target = target.cuda(device_id)
    target = target.to(torch.int64).contiguous()
    out = torch.empty((target.shape[0], maxx), dtype=torch.float32, device=target.device)
    _lib.call("sed_one_hot", _lib.ptr(target), target.shape[0], int(maxx), _lib.ptr(out), _lib.stream())
    return out
