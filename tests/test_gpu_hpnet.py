"""GPU parity of the deterministic part of hpnet_process: compute_entropy and the cache-hit branch
against the values recorded from the unmodified reference (tests/golden/hpnet.npz), and mean-shift on the resulting
148-column embedding against the oracle."""
import os

import numpy as np
import pytest
import torch

import oracle as O
import oracle_hpnet as OH
from conftest import ROOT
from util import canon

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    from sednet_b200.src import _lib
    _lib.load()
    return torch.device("cuda", 0)


def test_compute_entropy_and_hpnet_golden(dev, tmp_path):
    from sednet_b200.src import smooth_normal_matrix as snm
    g = np.load(os.path.join(ROOT, "tests", "golden", "hpnet.npz"))
    for case in (0, 1):
        seed, n, chunk = [int(v) for v in g[f"c{case}_cfg"]]
        feat, v, types, edges = OH.hpnet_case(seed, n)
        e_feat = float(snm.compute_entropy(feat.to(dev), CHUNK=chunk))
        e_v = float(snm.compute_entropy(v.to(dev), CHUNK=chunk))
        # FP32 sums of 4e6 terms in the reference (torch.sum per chunk pair), FP64 here: 1e-5 relative
        assert abs(e_feat - g[f"c{case}_ent"][0]) < 1e-5 * g[f"c{case}_ent"][0], (e_feat, g[f"c{case}_ent"])
        assert abs(e_v - g[f"c{case}_ent"][1]) < 1e-5 * g[f"c{case}_ent"][1]
        # cache-hit branch: the reference's file names, in a scratch working directory
        cache = tmp_path / f"c{case}" / "src" / "normal_smooth_cache"
        cache.mkdir(parents=True)
        torch.save(v, str(cache / "Us_7_0.1_50.pt"))
        torch.save(torch.tensor(float(g[f"c{case}_ent"][1])), str(cache / "WUs_7_0.1_50.pt"))
        cwd = os.getcwd()
        os.chdir(str(tmp_path / f"c{case}"))
        try:
            emb = snm.hpnet_process(feat.to(dev), torch.zeros(1, n, 3, device=dev), torch.zeros(1, n, 3, device=dev), id=7,
                                    types=types.to(dev), edges=edges.to(dev), normal_smooth_w=0.5, CHUNK=chunk)
        finally:
            os.chdir(cwd)
        assert emb.shape == (1, n, 148)
        ref = g[f"c{case}_emb_sample"]
        assert np.abs(emb[0, ::50].cpu().numpy() - ref).max() < 2e-5 * np.abs(ref).max()
        assert abs(float(emb.double().sum()) - float(g[f"c{case}_emb_sum"])) < 1e-4 * abs(float(g[f"c{case}_emb_sum"]))
    # (no cache, no v: the branch that builds the spectral vectors -- tests/test_hpnet_spectral.py)
    with torch.no_grad():
        e_o = float(OH.compute_entropy(types.exp(), CHUNK=chunk))
    assert abs(float(snm.compute_entropy(types.exp().to(dev), CHUNK=chunk)) - e_o) < 1e-5 * e_o     # K = 6


def test_meanshift_on_148_column_embedding(dev):
    """The driver's flow after hpnet_process (generate_predictions_aug.py:371-384): L2-normalise the 148-column embedding
    and cluster it.  Rows wider than 128 run on the FFMA kernels (bandwidth, shift, nms): labels, bandwidth and shifted
    points against the oracle."""
    from sednet_b200.src.mean_shift import MeanShift
    feat, v, types, edges = OH.hpnet_case(5, 3000)
    with torch.no_grad():
        emb = OH.hpnet_combine(feat, v, torch.tensor(0.3), types, edges, 0.5, 600)
        X = torch.nn.functional.normalize(emb[0], p=2, dim=1).contiguous()
        rX, rc, rbw, rlab = O.mean_shift(X, 10000, 0.015, 20)
    assert X.shape[1] == 148 and MeanShift()._mode(148) == 3 and MeanShift()._mode(200) == 0
    for mode in (0, 3):                   # FFMA kernel (rows padded to 192), 192-wide tensor-core kernel
        newX, center, bw, labels = MeanShift(prec_mode=mode).mean_shift(X.to(dev), 10000, 0.015, 20)
        assert (canon(labels.cpu().numpy()) == canon(rlab.numpy())).all(), mode
        assert abs(float(bw) - float(rbw)) < 1e-4 * float(rbw)
        assert float((newX.cpu() - rX).abs().max()) < 1e-4 and center.shape == rc.shape, mode
        assert float((torch.linalg.norm(newX, dim=1) - 1).abs().max()) < 1e-5
    # epanechnikov kernel, tight planted clusters, a point count that is no multiple of the 128-row tiles
    from sednet_b200 import synth
    _, _, lab, _, _ = synth.make_cloud(2403, 2777, n_patches=6, min_pts=170)
    Xe = torch.from_numpy(synth.make_embedding(lab, 148, 0.01, 3))
    with torch.no_grad():
        onew, _, _, olab = O.mean_shift(Xe, 10000, 0.015, 20, "epa")
    for mode in (0, 3):
        newX, _, _, labels = MeanShift(prec_mode=mode).mean_shift(Xe.to(dev), 10000, 0.015, 20, kernel_type="epa")
        assert float((newX.cpu() - onew).abs().max()) < 1e-4, mode
        assert (canon(labels.cpu().numpy()) == canon(olab.numpy())).all() and (canon(olab.numpy()) == canon(lab)).all()


@pytest.mark.parametrize("d,n", [(132, 300), (192, 129), (160, 1000)])
def test_meanshift_tc192_edge_widths(dev, d, n):
    """The 192-wide tensor-core kernel at the ends of its range (132 and 192 columns), with fewer points than one key
    tile, one iteration and a batch of clouds with different bandwidths, against the FP32 FFMA kernel."""
    import ctypes as C
    from sednet_b200 import synth
    from sednet_b200.src import _lib
    B = 3
    labs = [np.random.default_rng(s).integers(0, 4, n) for s in range(B)]
    X = torch.stack([torch.from_numpy(synth.make_embedding(labs[b], d, 0.05, 7 + b)) for b in range(B)]).to(dev).contiguous()
    bw = torch.tensor([0.3, 0.4, 0.5], device=dev)
    for iters in (1, 6):
        ref, got, tmp = torch.empty_like(X), torch.empty_like(X), torch.empty_like(X)
        _lib.call("sed_ms_shift", _lib.ptr(X), _lib.ptr(bw), B, n, d, iters, 0, 0, _lib.ptr(ref), _lib.ptr(tmp), _lib.stream())
        _lib.call("sed_ms_shift", _lib.ptr(X), _lib.ptr(bw), B, n, d, iters, 0, 3, _lib.ptr(got), _lib.ptr(tmp), _lib.stream())
        # mode 3 drops the X_lo term of O (a 2^-12 relative rounding of X that averages out over a cluster: fewer than 100
        # points per cluster here); the bar for shifted points is 1e-4
        assert float((got - ref).abs().max()) < 1e-4, (d, n, iters)
        assert float((torch.linalg.norm(got, dim=2) - 1).abs().max()) < 1e-5
