"""GPU parity of the evaluation row: the device metrics against the values recorded from the unmodified
reference functions (tests/golden/metrics.npz) and against oracle/oracle_metrics.py.  Matching, pair lists and IoUs are
integer work: bit-exact; chamfer within 1e-6 relative."""
import os

import numpy as np
import pytest
import torch

import oracle_metrics as OM
from conftest import ROOT
from sednet_b200 import synth
from util import t

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    from sednet_b200.src import _lib
    _lib.load()
    return torch.device("cuda", 0)


def _case(case, seed, n):
    pts, gt, typ_gt, pred, typ_pred = synth.make_metric_case(seed, n)
    if case == 3:
        pts = (pts * 3.0).astype(np.float32)
    return pts, gt, typ_gt, pred, typ_pred


def test_siou_metrics_golden(dev):
    from sednet_b200.src import segment_utils as su
    g = np.load(os.path.join(ROOT, "tests", "golden", "metrics.npz"))
    *seeds, n = [int(v) for v in g["cfg"]]
    for case, seed in enumerate(seeds):
        pts, gt, typ_gt, pred, typ_pred = _case(case, seed, n)
        weights = su.to_one_hot(pred, int(np.unique(pred).shape[0])).float()
        for usecd in (0, 1):
            args = (gt.copy(), pred.copy(), typ_pred.copy(), typ_gt.copy(), weights)
            res = su.SIOU_matched_segments_usecd(*args, t(pts).to(dev)) if usecd else su.SIOU_matched_segments(*args)
            s_iou, p_iou, matching, pairs, recall = res
            ref = g[f"c{case}_u{usecd}"]
            assert s_iou == ref[0] and p_iou == ref[1] and recall == ref[2], (case, usecd, s_iou, p_iou, recall, ref)
            assert np.array_equal(np.array(pairs), g[f"c{case}_u{usecd}_pairs"])
            assert np.array_equal(matching[0][0], g[f"c{case}_u{usecd}_rows"]) and np.array_equal(matching[0][1], g[f"c{case}_u{usecd}_cols"])
        # in-place remap of the type arrays, as the reference does
        tp = typ_pred.copy()
        su.SIOU_matched_segments(gt.copy(), pred.copy(), tp, typ_gt.copy(), weights)
        assert not np.isin(tp, [0, 6, 7, 8]).any()
        logits = np.random.default_rng(seed).normal(size=(1, n, 10)).astype(np.float32)
        logits[0, np.arange(n), typ_pred] += 3.0
        I_gt = gt.copy()
        I_gt[:50] = -1
        r = su.compute_type_miou_abc(t(logits).to(dev), t(typ_gt.copy())[None].to(dev), t(pred.copy())[None].to(dev),
                                     t(I_gt)[None].to(dev))
        assert abs(float(r) - float(g[f"c{case}_abc"])) < 1e-7
        from sednet_b200.src.utils import chamfer_distance
        a, b = pts[pred == 0], pts[gt == 1]
        cd = float(chamfer_distance(t(a)[None].to(dev), t(b)[None].to(dev)))
        assert abs(cd - float(g[f"c{case}_cd"])) < 1e-6 * float(g[f"c{case}_cd"])


def test_metrics_at_full_size_vs_oracle(dev):
    """10 000 points, 20 segments: the device tables / masked chamfer pass against the oracle's host loops."""
    from sednet_b200.src import segment_utils as su
    pts, gt, typ_gt, pred, typ_pred = synth.make_metric_case(77, 10000)
    pts = (pts * 2.5).astype(np.float32)
    weights = su.to_one_hot(pred, int(np.unique(pred).shape[0])).float()
    got = su.SIOU_matched_segments_usecd(gt.copy(), pred.copy(), typ_pred.copy(), typ_gt.copy(), weights, t(pts).to(dev))
    ref = OM.siou_matched_segments(gt.copy(), pred.copy(), typ_pred.copy(), typ_gt.copy(), weights.cpu().numpy(), pts)
    assert got[0] == ref[0] and got[1] == ref[1] and got[4] == ref[4]
    assert np.array_equal(np.array(got[3]), np.array(ref[3]))
    tab = su.segment_tables(pred, gt, typ_pred, typ_gt, K=50, T=10)
    assert tab["confusion"].sum() == 10000 and np.array_equal(tab["npred"][:pred.max() + 1], np.bincount(pred))
    assert np.array_equal(tab["gt_first"][:gt.max() + 1], [int(np.flatnonzero(gt == c)[0]) for c in range(gt.max() + 1)])
    # two-cloud chamfer: row / column minima against a brute-force FP32 evaluation
    from sednet_b200.src import _lib
    a, b = t(pts[:3000])[None].to(dev), t(pts[3000:7100])[None].to(dev)
    ma, mb = torch.empty((1, 3000), device=dev), torch.empty((1, 4100), device=dev)
    _lib.call("sed_chamfer_min", _lib.ptr(a), _lib.ptr(b), 1, 3000, 4100, _lib.ptr(ma), _lib.ptr(mb), _lib.stream())
    d = torch.sum((a[0].cpu()[:, None, :] - b[0].cpu()[None, :, :]) ** 2, 2)
    assert torch.equal(ma[0].cpu(), d.min(1)[0]) and torch.equal(mb[0].cpu(), d.min(0)[0])


def test_driver_flow_end_to_end(dev):
    """The per-shape body of generate_predictions_aug.py written against the drop-in modules
    (examples/predict_like_driver.py) against the oracle chain: labels (canonicalised), types and metrics identical."""
    import importlib.util
    import oracle as O
    from util import canon
    spec = importlib.util.spec_from_file_location("predict_like_driver", os.path.join(ROOT, "examples", "predict_like_driver.py"))
    drv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(drv)
    n, k = 2000, 32
    pts, nrm, lab, typ, _ = synth.make_cloud(555, n, n_patches=6)
    sd_t, sd_i = synth.make_state_dict(0), synth.make_state_dict(1, randomize_gn=True)
    out = drv.predict_shape(drv.build_model(sd_t, k), drv.build_model(sd_i, k), t(pts)[None], t(nrm)[None], lab[None], typ[None])
    with torch.no_grad():
        ref = O.end_to_end({kk: t(v) for kk, v in sd_t.items()}, {kk: t(v) for kk, v in sd_i.items()}, t(pts)[None], t(nrm)[None], k)[0]
    assert (canon(out["cluster_ids"]) == canon(ref["labels"])).all()
    assert (out["pred_primitives"] == ref["types"]).mean() > 0.999
    rl = canon(ref["labels"])
    w = OM.one_hot(rl, int(np.unique(rl).shape[0]))
    o = OM.siou_matched_segments(lab.copy(), rl, ref["types"].copy(), typ.copy(), w, pts)
    # the label NUMBERING of the two paths may differ (canonical relabelling): the metrics do not depend on it
    assert abs(out["s_iou"] - o[0]) < 1e-12 and abs(out["s_recall"] - o[4]) < 1e-12
    params, status, seg_type = drv.stage2_fits(pts, nrm, lab, typ)
    assert (status[: int(lab.max()) + 1] == 0).all()
    # the same flow with HPNet_embed = True on given spectral vectors: 148-column embedding -> compute_entropy weights ->
    # tensor-core mean-shift at 192 columns, against the oracle chain
    import oracle_hpnet as OH
    g = torch.Generator().manual_seed(3)
    v = torch.randn((1, n, 12), generator=g)
    v = v / (torch.norm(v, dim=-1, keepdim=True) + 1e-16)
    out_h = drv.predict_shape(drv.build_model(sd_t, k), drv.build_model(sd_i, k), t(pts)[None], t(nrm)[None], lab[None], typ[None],
                              spectral_v=v, spectral_ent=torch.tensor(0.3), chunk=400)
    with torch.no_grad():
        sdt, sdi = {kk: t(vv) for kk, vv in sd_t.items()}, {kk: t(vv) for kk, vv in sd_i.items()}
        x6 = torch.cat([t(pts)[None], t(nrm)[None]], 2).permute(0, 2, 1)
        logp = O.sednet_forward(sdt, x6, k)[1]
        fo = O.sednet_forward(sdi, x6, k)
        emb = OH.hpnet_combine(fo[0].transpose(1, 2), v, torch.tensor(0.3), logp.transpose(1, 2), fo[3].transpose(1, 2), 0.5, 400)
        X = torch.nn.functional.normalize(emb[0], p=2, dim=1)
        _, rbw, rlab = O.guard_mean_shift(X, 0.015, 50)
    assert X.shape[1] == 148 and (canon(out_h["cluster_ids"]) == canon(rlab.numpy())).all()
    assert abs(out_h["bw"] - float(rbw)) < 1e-3 * float(rbw)

