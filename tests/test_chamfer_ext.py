"""The reference's chamfer extension (src/chamfer_distance/chamfer_distance.cu / .py) rebuilt on libsednet_b200.so.  The
extension is CUDA-only and JIT-compiled at import, so it cannot run in the CPU-only build container: the oracle
(oracle_metrics.chamfer_ext_*) restates the kernel text and is itself checked here against torch autograd."""
import numpy as np
import pytest
import torch

import oracle_metrics as OM


def _clouds(seed, B, n, m, dup=False):
    g = torch.Generator().manual_seed(seed)
    a, b = torch.randn((B, n, 3), generator=g), torch.randn((B, m, 3), generator=g)
    if dup:                       # exact ties: duplicated candidates -> the lowest index must win
        b[:, m // 2:] = b[:, : m - m // 2]
    return a, b


def test_oracle_chamfer_backward_equals_autograd():
    a, b = _clouds(1, 2, 300, 170)
    d1, d2, i1, i2 = OM.chamfer_ext_forward(a, b)
    g1, g2 = torch.rand_like(d1), torch.rand_like(d2)
    gx1, gx2 = OM.chamfer_ext_backward(a, b, g1, g2, i1, i2)
    A, Bt = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    l = (g1 * ((A - torch.stack([Bt[k][i1[k].long()] for k in range(2)])) ** 2).sum(-1)).sum() + \
        (g2 * ((Bt - torch.stack([A[k][i2[k].long()] for k in range(2)])) ** 2).sum(-1)).sum()
    l.backward()
    assert float((gx1 - A.grad).abs().max()) < 1e-5 and float((gx2 - Bt.grad).abs().max()) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("B,n,m,dup", [(2, 600, 300, False), (1, 2500, 4097, False), (3, 100, 1025, True), (1, 1, 7, False)])
def test_gpu_chamfer_extension_forward_backward(B, n, m, dup):
    from sednet_b200.src.chamfer_distance import ChamferDistance, ChamferIndex
    dev = torch.device("cuda", 0)
    a, b = _clouds(7 + n, B, n, m, dup)
    od1, od2, oi1, oi2 = OM.chamfer_ext_forward(a, b)
    A, Bt = a.to(dev).requires_grad_(True), b.to(dev).requires_grad_(True)
    d1, d2 = ChamferDistance()(A, Bt)
    i1, i2 = ChamferIndex()(A, Bt)
    assert i1.dtype == torch.int32 and d1.shape == (B, n) and d2.shape == (B, m)
    assert float((d1.detach().cpu() - od1).abs().max()) <= 1e-6 * float(od1.max()) + 1e-7
    assert float((d2.detach().cpu() - od2).abs().max()) <= 1e-6 * float(od2.max()) + 1e-7
    # indices: identical except where two candidates are closer than FP32 rounding (the kernel contracts x^2+y^2+z^2 into FMAs)
    for got, want, q, c in ((i1.cpu(), oi1, a, b), (i2.cpu(), oi2, b, a)):
        diff = got != want
        if diff.any():
            bb, jj = torch.nonzero(diff, as_tuple=True)
            dg = ((q[bb, jj] - c[bb, got[bb, jj].long()]).double() ** 2).sum(-1)
            dw = ((q[bb, jj] - c[bb, want[bb, jj].long()]).double() ** 2).sum(-1)
            assert float(((dg - dw).abs() / dw.clamp(min=1e-12)).max()) < 1e-6
            assert diff.float().mean() < 1e-3
    if dup:
        assert torch.equal(i1.cpu(), oi1) and int(i1.max()) < m - m // 2 + (m % 2)        # ties -> lowest index
    g1, g2 = torch.rand((B, n), generator=torch.Generator().manual_seed(3)), torch.rand((B, m), generator=torch.Generator().manual_seed(4))
    (d1 * g1.to(dev)).sum().backward(retain_graph=True)
    (d2 * g2.to(dev)).sum().backward()
    ogx1, ogx2 = OM.chamfer_ext_backward(a, b, g1, g2, i1.cpu(), i2.cpu())
    assert float((A.grad.cpu() - ogx1).abs().max()) <= 1e-5 * max(1.0, float(ogx1.abs().max()))
    assert float((Bt.grad.cpu() - ogx2).abs().max()) <= 1e-5 * max(1.0, float(ogx2.abs().max()))
