"""Gradient of the mean-shift iterations (training path: reference src/segment_loss.py:50-56 differentiates
src/mean_shift.py:45-79 with autograd).  tests/golden/ms_backward.npz holds the gradients the UNMODIFIED reference returns
(oracle/make_golden_ms_backward.py); the CPU test pins the oracle's autograd to them, the GPU tests compare
MeanShift.mean_shift_ + csrc/meanshift_bwd.cu with the fixture and, at a size the fixture does not cover, with the oracle's
autograd on the same device."""
import numpy as np
import pytest
import torch

import oracle as O
from sednet_b200 import synth
from util import t

CASES = ("a", "b", "c")


def test_oracle_gradient_matches_reference_fixture(golden):
    g = golden("ms_backward")
    for tag in CASES:
        n, d, _, iters = [int(v) for v in g[tag + "_cfg"]]
        b = float(g[tag + "_sigma_b"][1])
        X = t(g[tag + "_x"]).requires_grad_(True)
        new_X = O.ms_shift(X, torch.tensor(b), iters)
        (new_X * t(g[tag + "_w"])).sum().backward()
        assert float((new_X.detach() - t(g[tag + "_newx"])).abs().max()) < 1e-6
        assert float((X.grad - t(g[tag + "_grad"])).abs().max()) <= 1e-5 * float(np.abs(g[tag + "_grad"]).max())


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from sednet_b200.src import _lib
    _lib.load()
    return torch.device("cuda", 0)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1, 4])
def test_gradient_matches_reference_fixture(dev, golden, mode):
    """Relative to the largest gradient entry: 2e-4 (the forward states come from the FP32 FFMA kernel resp. the FP16-split
    tensor-core kernel, the backward recomputes the weights in FP32 with ex2.approx)."""
    from sednet_b200.src.mean_shift import MeanShift
    g = golden("ms_backward")
    for tag in CASES:
        n, d, _, iters = [int(v) for v in g[tag + "_cfg"]]
        b = float(g[tag + "_sigma_b"][1])
        X = t(g[tag + "_x"]).to(dev).requires_grad_(True)
        new_X, keys = MeanShift(prec_mode=mode).mean_shift_(X, b=torch.tensor(b), iterations=iters)
        assert keys is X and new_X.requires_grad
        (new_X * t(g[tag + "_w"]).to(dev)).sum().backward()
        gref = t(g[tag + "_grad"]).to(dev)
        assert float((new_X.detach() - t(g[tag + "_newx"]).to(dev)).abs().max()) < 2e-5
        assert float((X.grad - gref).abs().max()) <= 2e-4 * float(gref.abs().max()), tag


@pytest.mark.gpu
def test_gradient_matches_oracle_autograd_ragged_size(dev):
    """N = 1 500 (not a multiple of the 64-row tiles), 12 clusters, 5 iterations as in the triplet loss; the same through
    mean_shift(..., nms=False), whose bandwidth is a constant of the graph."""
    from sednet_b200.src.mean_shift import MeanShift
    N, d, iters = 1500, 128, 5
    lab = np.random.default_rng(3).integers(0, 12, N)
    x = t(synth.make_embedding(lab, d, 0.02, 11)).to(dev)
    W = torch.randn(N, d, device=dev, generator=torch.Generator(device="cuda").manual_seed(1))
    b = torch.tensor(0.2, device=dev)
    Xo = x.clone().requires_grad_(True)
    (O.ms_shift(Xo, b, iters) * W).sum().backward()
    Xk = x.clone().requires_grad_(True)
    new_X, _ = MeanShift(prec_mode=1).mean_shift_(Xk, b=b, iterations=iters)
    (new_X * W).sum().backward()
    scale = float(Xo.grad.abs().max())
    assert torch.isfinite(Xk.grad).all()
    assert float((Xk.grad - Xo.grad).abs().max()) <= 2e-4 * scale
    # end to end as the loss calls it: normalised network output -> mean_shift(nms=False)
    E = torch.randn(N, d, device=dev, generator=torch.Generator(device="cuda").manual_seed(2)).requires_grad_(True)
    np.random.seed(0)
    out, bw = MeanShift(prec_mode=1).mean_shift(torch.nn.functional.normalize(x + 0.01 * E, p=2, dim=1), 4000, 0.015,
                                                iterations=iters, nms=False)
    (out * W).sum().backward()
    assert torch.isfinite(E.grad).all() and float(E.grad.abs().max()) > 0 and not bw.requires_grad


@pytest.mark.gpu
def test_inference_path_is_untouched_without_grad(dev):
    """No autograd graph, same bits as before: requires_grad False or torch.no_grad() take the fused multi-iteration call."""
    from sednet_b200.src.mean_shift import MeanShift
    lab = np.random.default_rng(5).integers(0, 6, 700)
    x = t(synth.make_embedding(lab, 128, 0.02, 2)).to(dev)
    ms = MeanShift(prec_mode=1)
    a, _ = ms.mean_shift_(x, b=torch.tensor(0.2), iterations=4)
    with torch.no_grad():
        bq, _ = ms.mean_shift_(x.clone().requires_grad_(True), b=torch.tensor(0.2), iterations=4)
    c, _ = ms.mean_shift_(x.clone().requires_grad_(True), b=torch.tensor(0.2), iterations=4)
    assert not a.requires_grad and not bq.requires_grad and c.requires_grad
    assert torch.equal(a, bq)
    assert float((a - c.detach()).abs().max()) < 1e-6      # one-iteration calls: same kernel, states re-split from FP32
