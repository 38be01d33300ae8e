"""The per-shape body of the reference's inference driver (generate_predictions_aug.py:213-236, 365, 379-411, 424-437)
written against the drop-in modules of this repository: two network forwards, type argmax, normalised embedding,
guarded mean-shift, one-hot segment weights, the evaluation metrics, the text files for stage 2, and the stage-2 fits.

    python examples/predict_like_driver.py [n_points]        (needs a B200; synthetic cloud, random-initialised weights)
"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sednet_b200 import synth                                                      # noqa: E402
from sednet_b200.src.SEDNet import SEDNet                                          # noqa: E402  (src.SEDNet)
from sednet_b200.src.mean_shift import MeanShift                                   # noqa: E402  (src.mean_shift)
from sednet_b200.src.segment_utils import SIOU_matched_segments_usecd, to_one_hot  # noqa: E402  (src.segment_utils)
from sednet_b200.Fitting_patches_and_edges import wire                             # noqa: E402
from sednet_b200.Fitting_patches_and_edges.primitive_forward_v2 import fit_segments_batched_v2  # noqa: E402


def guard_mean_shift(ms, embedding, quantile, iterations, kernel_type="gaussian"):
    """generate_predictions_aug.py:25-35"""
    while True:
        _, center, bandwidth, cluster_ids = ms.mean_shift(embedding, 10000, quantile, iterations, kernel_type=kernel_type)
        if torch.unique(cluster_ids).shape[0] > 49:
            quantile *= 1.2
        else:
            break
    return center, bandwidth, cluster_ids


def build_model(state, k=64):
    m = SEDNet(embedding=True, emb_size=128, primitives=True, num_primitives=6, mode=5, num_channels=6,
               combine_label_prim=True, edge_module=True, late_fusion=True, nn_nb=k)      # :142-154
    m.load_state_dict({kk: torch.from_numpy(v) for kk, v in state.items()})
    return m.cuda().eval()


def predict_shape(model, model_inst, points_, normals_, labels, primitives_, quantile=0.015, iterations=50,
                  spectral_v=None, spectral_ent=None, normal_smooth_w=0.5, chunk=1000, hpnet_embed=False):
    """points_, normals_ (1,N,3) float32 host tensors; labels, primitives_ (1,N) numpy ground truth.  With hpnet_embed
    (HPNet_embed = True, the driver's default, generate_predictions_aug.py:58,371-378) the embedding goes through
    hpnet_process first: spectral_v (1,N,12) / spectral_ent are the shape's cached normal-smoothness eigenvectors, None
    builds them (farthest-50 table, factored affinity matrix, 10 LOBPCG steps).  Returns a dict."""
    hpnet_embed = hpnet_embed or spectral_v is not None
    points, normals = points_.cuda(), normals_.cuda()
    with torch.no_grad():
        _input = torch.cat([points, normals], 2)                                                   # :223
        primitives_log_prob = model(_input.permute(0, 2, 1), None, False)[1]                       # :224-226
        embedding, _, _, edges_pred = model_inst(_input.permute(0, 2, 1), None, False)             # :227-229
    pred_primitives = torch.max(primitives_log_prob[0], 0)[1].data.cpu().numpy()                   # :365
    if hpnet_embed:
        from sednet_b200.src.smooth_normal_matrix import hpnet_process                             # src.smooth_normal_matrix
        embedding = hpnet_process(embedding.transpose(1, 2), points, normals, types=primitives_log_prob.transpose(1, 2),
                                  edges=edges_pred.transpose(1, 2), normal_smooth_w=normal_smooth_w, CHUNK=chunk,
                                  v=None if spectral_v is None else spectral_v.cuda(), ent=spectral_ent)    # :372-376
        embedding = torch.nn.functional.normalize(embedding[0], p=2, dim=1).contiguous()           # :377
    else:
        embedding = torch.nn.functional.normalize(embedding[0].T, p=2, dim=1)                      # :380
    _, bw, cluster_ids = guard_mean_shift(MeanShift(), embedding, quantile, iterations)            # :382-384
    weights = to_one_hot(cluster_ids, np.unique(cluster_ids.data.cpu().numpy()).shape[0])          # :385-386
    cluster_ids = cluster_ids.data.cpu().numpy()
    s_iou, p_iou, _, _, s_recall = SIOU_matched_segments_usecd(labels[0], cluster_ids, pred_primitives.copy(),
                                                               primitives_[0].copy(), weights, points[0])   # :389-396
    return dict(cluster_ids=cluster_ids, pred_primitives=pred_primitives, edges=edges_pred[0].cpu().numpy(), bw=float(bw),
                s_iou=s_iou, p_iou=p_iou, s_recall=s_recall)


def stage2_fits(points, normals, inst, types):
    """What stage 2 does with the files of stage 1: one batched launch for every segment (stage-1 type ids)."""
    dev = torch.device("cuda")
    S = int(inst.max()) + 1
    seg_type = np.zeros((1, S), np.int32)
    for s in range(S):
        m = inst == s
        seg_type[0, s] = np.bincount(types[m]).argmax() if m.any() else 0
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    params, status = fit_segments_batched_v2(t(points)[None], t(normals)[None], t(inst.astype(np.int64))[None], t(seg_type),
                                             plane_filter_ratio=0.25)
    return params[0].cpu().numpy(), status[0].cpu().numpy(), seg_type[0]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    pts, nrm, lab, typ, _ = synth.make_cloud(1234, n)
    model, model_inst = build_model(synth.make_state_dict(0)), build_model(synth.make_state_dict(1, randomize_gn=True))
    out = predict_shape(model, model_inst, torch.from_numpy(pts)[None], torch.from_numpy(nrm)[None], lab[None], typ[None],
                        hpnet_embed=True)      # HPNet_embed = True is the driver's default (generate_predictions_aug.py:58)
    print(f"segments {len(np.unique(out['cluster_ids']))}  bw {out['bw']:.4f}  s_iou {out['s_iou']:.4f}  p_iou {out['p_iou']:.4f}  "
          f"recall {out['s_recall']:.4f}")
    with tempfile.TemporaryDirectory() as d:
        wire.write_stage1(d, 0, pts, nrm, out["cluster_ids"], out["pred_primitives"], out["edges"])
        back = wire.read_stage1(d, 0)
    params, status, seg_type = stage2_fits(back["points"], back["normals"], back["inst"], back["types"])
    print("stage-2 fits:", [(int(ty), int(st)) for ty, st in zip(seg_type, status)])


if __name__ == "__main__":
    main()
