"""Hot-path subset of reference src/fitting_utils.py: LeastSquares.lstsq / best_lambda (:32-85), customsvd forward
with its custom backward (:385-455) and weights_normalize (:306-325), on the sm_100a fit kernels (FP64 moment accumulation + 3x3 Jacobi)."""
import numpy as np
import torch

from . import _lib
from .guard import guard_exp

EPS = float(np.finfo(np.float32).eps)


class LeastSquares:
    def __init__(self):
        self.last_status = None

    def lstsq(self, A, Y, lamb=0.0):
        """src/fitting_utils.py:36-65: A (m,3), Y (m,1) -> x (3,1). Full column rank -> the QR solution; otherwise
        (AtA + lambda I)^-1 At Y with lambda from best_lambda's ladder.  `lamb` is ignored, as in the reference."""
        A = _lib.require_cuda(A, name="A")
        Y = _lib.require_cuda(Y, name="Y").reshape(-1)
        if A.dim() != 2 or A.shape[1] != 3:
            raise NotImplementedError("lstsq is implemented for (m, 3) systems, the only shape on the fitting path")
        x = torch.empty((3, 1), dtype=torch.float32, device=A.device)
        st = torch.empty(1, dtype=torch.int32, device=A.device)
        _lib.call("sed_lstsq3", _lib.ptr(A), _lib.ptr(Y), A.shape[0], _lib.ptr(x), _lib.ptr(st), _lib.stream())
        self.last_status = st
        return x


def best_lambda(A):
    """src/fitting_utils.py:68-85 for a symmetric 3x3 A: smallest 1e-6 * 10^i with A + lambda I of full rank."""
    _, S, _ = _svd3_forward(A)
    s_max, s_min = float(S[0]), float(S[-1])   # singular values of a symmetric PSD matrix = eigenvalues
    lamb = 1e-6
    for _ in range(7):
        if (s_min + lamb) > (s_max + lamb) * 3 * EPS:
            break
        lamb *= 10
    return lamb


def _svd3_forward(x):
    x = _lib.require_cuda(x, name="x")
    if x.dim() != 2 or x.shape[1] != 3:
        raise NotImplementedError("customsvd is implemented for (m, 3) matrices, the only shape on the fitting path")
    S = torch.empty(3, dtype=torch.float32, device=x.device)
    V = torch.empty((3, 3), dtype=torch.float32, device=x.device)
    _lib.call("sed_svd3", _lib.ptr(x), x.shape[0], _lib.ptr(S), _lib.ptr(V), _lib.stream())
    U = (x @ V) / torch.clamp(S, min=1e-30)
    return U, S, V


class CustomSVD(torch.autograd.Function):
    """src/fitting_utils.py:420-452: forward = torch.svd(input, some=True) of an (m,3) matrix; backward lets only grad_V flow
    back (compute_grad_V / svd_grad_K, :385-417), with |S_i - S_j| floored at 1e-6 so equal singular values do not blow up."""

    @staticmethod
    def forward(ctx, input):
        U, S, V = _svd3_forward(input.detach())
        ctx.save_for_backward(U, S, V)
        return U, S, V

    @staticmethod
    def backward(ctx, grad_U, grad_S, grad_V):
        U, S, V = ctx.saved_tensors
        if grad_V is None:
            return torch.zeros_like(U)
        gV = _lib.require_cuda(grad_V, name="grad_V")
        gin = torch.empty_like(U)
        _lib.call("sed_svd3_backward", _lib.ptr(U.contiguous()), _lib.ptr(S), _lib.ptr(V), _lib.ptr(gV), U.shape[0],
                  _lib.ptr(gin), _lib.stream())
        return gin


customsvd = CustomSVD.apply


def weights_normalize(weights, bw):
    """src/fitting_utils.py:306-325 (elementwise; not a hot loop)."""
    prob = guard_exp(weights / (bw ** 2) / 2)
    prob = prob / torch.sum(prob, 0, keepdim=True)
    if weights.shape[0] == 1:
        return prob
    prob = prob - torch.min(prob, 1, keepdim=True)[0]
    return prob / (torch.max(prob, 1, keepdim=True)[0] + EPS)
