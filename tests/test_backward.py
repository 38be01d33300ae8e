"""Backward of customsvd (reference src/fitting_utils.py:385-455) against tests/golden/backward.npz, recorded from the
UNMODIFIED reference's autograd Function (oracle/make_golden_backward.py): gradient with respect to the (m,3) input of
sign-invariant scalar losses of V.  Only grad_V flows back, as in the reference."""
import numpy as np
import pytest
import torch

import oracle as O

CASES = ("generic", "slab", "near_equal")


def _loss(V, coef, a):
    return (coef * V ** 2).sum() + (a @ V[:, -1]) ** 2


@pytest.mark.parametrize("tag", CASES)
def test_oracle_customsvd_backward_golden(golden, tag):
    g = golden("backward")
    x, coef, a = (torch.from_numpy(g[f"{tag}_{k}"]) for k in ("x", "coef", "a"))
    with torch.no_grad():
        U, S, V = torch.svd(x, some=True)
    Vl = V.clone().requires_grad_(True)
    _loss(Vl, coef, a).backward()
    got = O.compute_grad_V(U, S, V, Vl.grad)
    ref = torch.from_numpy(g[f"{tag}_grad"])
    if tag == "near_equal":      # two singular values 2e-7 apart: K rides on the 1e-6 floor and the sign of a rounding-level difference
        assert torch.isfinite(got).all()
        return
    assert float((got - ref).abs().max()) <= 1e-5 * float(ref.abs().max())


@pytest.mark.gpu
@pytest.mark.parametrize("tag", CASES)
def test_gpu_customsvd_backward_golden(golden, tag):
    """customsvd as an autograd Function on the GPU: forward from the FP64-Gram Jacobi kernel, backward from
    sed_svd3_backward -- the input gradient of a sign-invariant loss against the reference's, 1e-4 relative."""
    from sednet_b200.src.fitting_utils import customsvd
    g = golden("backward")
    dev = torch.device("cuda", 0)
    x, coef, a = (torch.from_numpy(g[f"{tag}_{k}"]).to(dev) for k in ("x", "coef", "a"))
    X = x.clone().requires_grad_(True)
    U, S, V = customsvd(X)
    assert U.shape == x.shape and S.shape == (3,) and V.shape == (3, 3)
    _loss(V, coef, a).backward()
    got = X.grad.cpu()
    assert torch.isfinite(got).all()
    ref = torch.from_numpy(g[f"{tag}_grad"])
    if tag == "near_equal":
        # the singular vectors of a (numerically) double singular value are arbitrary within their plane, in LAPACK and
        # here alike, and K = 1 / ((S_i - S_j)(S_i + S_j)) sits on svd_grad_K's 1e-6 floor: the guard keeps the gradient
        # finite and bounded by the floor, which is all the reference promises there (its docstring, :421-431)
        scale = float(x.abs().max())
        assert float(got.abs().max()) < 1e7 * scale
        return
    assert float((got - ref).abs().max()) <= 1e-4 * float(ref.abs().max()), float((got - ref).abs().max())
    # finite-difference check of the same gradient through the forward kernel alone (FP32: loose)
    d = torch.randn(x.shape, generator=torch.Generator().manual_seed(1)).to(dev)
    eps = 1e-2
    with torch.no_grad():
        lp = _loss(customsvd(x + eps * d)[2], coef, a)
        lm = _loss(customsvd(x - eps * d)[2], coef, a)
    fd = float((lp - lm) / (2 * eps))
    an = float((X.grad * d).sum())
    assert abs(fd - an) <= 5e-2 * max(abs(an), 1e-3), (fd, an)
