"""Evaluation helpers of reference src/segment_utils.py with the reference's names and signatures: to_one_hot (:536-545),
relaxed_iou_fast (:609-627), primitive_type_segment_torch (:509-517), SIOU_matched_segments[_usecd] (:140-243),
mean_IOU_primitive_segment[_usecd] (:359-495), hungarian_matching / compute_type_miou_abc (:258-357).

The reference multiplies (N x 50) one-hot matrices and then scans the N points once per matched pair on the host.  Here
one device pass fills the small integer tables all of that derives from (sed_segment_tables), the chamfer terms of every
matched pair come from one masked nearest-point pass (sed_matched_chamfer), and the host only solves the 50 x 50
assignment (as the reference does) and reads the tables."""
import numpy as np
import torch

from . import _lib

try:                                       # the reference's solver when it is installed, scipy's otherwise
    from lapsolver import solve_dense
except Exception:                          # pragma: no cover
    from scipy.optimize import linear_sum_assignment as solve_dense

MAX_LABELS, MAX_TYPES = 64, 16


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


def _dev(a, dtype=torch.int64):
    t = a if isinstance(a, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(a))
    return _lib.require_cuda(t.cuda() if not t.is_cuda else t, dtype, "labels")


def segment_tables(pred, gt, type_pred=None, type_gt=None, K=50, T=10):
    """Device pass behind every metric below: dict of host numpy tables (see sed_segment_tables in the C header)."""
    p, g = _dev(pred).reshape(1, -1), _dev(gt).reshape(1, -1)
    N = p.shape[1]
    tp = _dev(type_pred).reshape(1, -1) if type_pred is not None else None
    tg = _dev(type_gt).reshape(1, -1) if type_gt is not None else None
    i32 = lambda *s: torch.empty(s, dtype=torch.int32, device=p.device)
    conf, npred, ngt, pt, gtt, first = i32(K, K), i32(K), i32(K), i32(K, T), i32(K, T), i32(K)
    _lib.call("sed_segment_tables", _lib.ptr(p), _lib.ptr(g), _lib.ptr(tp), _lib.ptr(tg), 1, N, K, T, _lib.ptr(conf),
              _lib.ptr(npred), _lib.ptr(ngt), _lib.ptr(pt), _lib.ptr(gtt), _lib.ptr(first), _lib.stream())
    return dict(confusion=conf.cpu().numpy().astype(np.int64), npred=npred.cpu().numpy().astype(np.int64),
                ngt=ngt.cpu().numpy().astype(np.int64), pred_types=pt.cpu().numpy(), gt_types=gtt.cpu().numpy(),
                gt_first=first.cpu().numpy())


def relaxed_iou_fast(pred, gt, max_clusters=50):
    """src/segment_utils.py:609-627: pred, gt (B,N,K) one-hot / soft memberships -> (B,K,K) relaxed IoU.  (Kept for API
    parity on arbitrary memberships; the SIOU functions below get the same matrix from the label confusion table.)"""
    dots = torch.matmul(pred.transpose(1, 2), gt)
    return dots / (torch.sum(pred, 1).unsqueeze(2) + torch.sum(gt, 1).unsqueeze(1) - dots + 1e-7)


def primitive_type_segment_torch(pred, weights):
    """src/segment_utils.py:509-517: pred (N,L) one-hot types, weights (N,K) -> (K,) type of every segment."""
    pred = _lib.require_cuda(pred, name="pred")
    weights = _lib.require_cuda(weights, name="weights")
    N, L = pred.shape
    K = weights.shape[1]
    types = torch.argmax(pred, 1).contiguous()          # the rows are one-hot (to_one_hot(primitives_pred, 10), :178)
    out = torch.empty((L, K), dtype=torch.float32, device=pred.device)
    _lib.call("sed_type_vote_weighted", _lib.ptr(types), _lib.ptr(weights), N, K, L, _lib.ptr(out), _lib.stream())
    return torch.max(out, 0)[1]


def _remap(a):
    a[a == 0] = 9
    a[a == 6] = 9
    a[a == 7] = 9
    a[a == 8] = 2


def _siou(target, pred_labels, primitives_pred, primitives, weights, points):
    _remap(primitives)                                   # in place, as the reference (:156-164)
    _remap(primitives_pred)
    tab = segment_tables(pred_labels, target, None, None, K=50, T=10)
    conf, npred, ngt = tab["confusion"], tab["npred"], tab["ngt"]
    dots = conf.astype(np.float32)
    cost = dots / (npred.astype(np.float32)[:, None] + ngt.astype(np.float32)[None, :] - dots + np.float32(1e-7))
    rids, cids = solve_dense(1.0 - cost)
    matching = [[rids, cids]]
    primitives_pred_hot = to_one_hot(primitives_pred, 10, weights.device.index).float()
    prim_pred = primitive_type_segment_torch(primitives_pred_hot, weights).cpu().numpy()
    keep = [(r, c) for r, c in zip(rids, cids) if ngt[c] > 0 and npred[r] > 0]
    recall_pos = 0
    if points is not None and keep:
        p2g, g2p = np.full((1, 50), -1, np.int32), np.full((1, 50), -1, np.int32)
        for r, c in keep:
            p2g[0, r], g2p[0, c] = c, r
        pts = _lib.require_cuda(points if isinstance(points, torch.Tensor) else torch.as_tensor(points), name="points")
        pl, gl = _dev(pred_labels), _dev(target)
        N = pl.shape[0]
        mp = torch.empty(N, dtype=torch.float32, device=pts.device)
        mg = torch.empty(N, dtype=torch.float32, device=pts.device)
        p2g_d, g2p_d = torch.from_numpy(p2g).to(pts.device), torch.from_numpy(g2p).to(pts.device)   # kept alive over the call
        _lib.call("sed_matched_chamfer", _lib.ptr(pts), _lib.ptr(pl), _lib.ptr(gl), _lib.ptr(p2g_d), _lib.ptr(g2p_d), 1, N, 50,
                  _lib.ptr(mp), _lib.ptr(mg), _lib.stream())
        for r, c in keep:                                # chamfer_distance(points[pred == r], points[gt == c]) / 2 (:473)
            cd = (torch.mean(mp[pl == r]) + torch.mean(mg[gl == c])) / 2.0
            if cd / 2 < 0.1:
                recall_pos += 1
    iou_b, recall_b, iou_b_prim, iou_b_prims = [], [], [], []
    for r, c in keep:
        if points is None and ngt[c] < 100:              # :389-390 (the _usecd variant keeps small gt segments)
            continue
        tp = conf[r, c]
        iou_b.append(tp / ((npred[r] + ngt[c] - tp) + 1e-8))
        recall_b.append(tp / (tp + (ngt[c] - tp) + 1e-8))
        gt_type = primitives[tab["gt_first"][c]]
        iou_b_prim.append(gt_type == prim_pred[r])
        iou_b_prims.append([gt_type, prim_pred[r]])
    recall = np.mean(recall_b) if points is None else float(recall_pos) / np.unique(target).shape[0]
    return np.mean(iou_b), np.mean(iou_b_prim), matching, iou_b_prims, np.mean(recall)


def SIOU_matched_segments(target, pred_labels, primitives_pred, primitives, weights):
    """src/segment_utils.py:140-191 -> (segment_iou, primitive_iou, matching, [[gt type, pred type]], segment_recall).
    target, pred_labels, primitives_pred, primitives: (N,) numpy int arrays; weights (N,K) CUDA tensor."""
    return _siou(target, pred_labels, primitives_pred, primitives, weights, None)


def SIOU_matched_segments_usecd(target, pred_labels, primitives_pred, primitives, weights, points):
    """src/segment_utils.py:194-243: as above with every gt segment kept and the recall counted over the matched pairs
    whose chamfer distance / 2 is below 0.1; points (N,3) CUDA tensor."""
    return _siou(target, pred_labels, primitives_pred, primitives, weights, points)


def hungarian_matching(W_pred, W_gt):
    """src/segment_utils.py:258-277 on (N,K), (N,K') memberships (numpy): maximal relaxed-IoU assignment."""
    dot = W_pred.T @ W_gt
    den = W_pred.sum(0)[:, None] + W_gt.sum(0)[None, :] - dot
    return solve_dense(-(dot / np.maximum(den, 1e-10)))


def compute_type_miou_abc(type_per_point, T_gt, cluster_pred, I_gt):
    """src/segment_utils.py:300-357: type_per_point (1,N,K) scores or (1,N) labels, T_gt, cluster_pred, I_gt (1,N) int64
    tensors -> share of Hungarian-matched (pred, gt) instances whose modal point types agree (0-d tensor)."""
    assert type_per_point.shape[0] == 1
    T_pred = torch.argmax(type_per_point, dim=-1) if type_per_point.dim() == 3 else type_per_point
    for a in (T_pred, T_gt):                             # in place on label inputs, as the reference (:315-323)
        a[a == 6] = 0
        a[a == 7] = 0
        a[a == 9] = 0
        a[a == 8] = 2
    Kp, Kg = int(cluster_pred.max()) + 1, int(I_gt.max()) + 1
    K = max(Kp, Kg)
    if K > MAX_LABELS:
        raise RuntimeError(f"more than {MAX_LABELS} instance labels")
    tab = segment_tables(cluster_pred[0], I_gt[0], T_pred[0], T_gt[0], K=K, T=MAX_TYPES)
    dot = tab["confusion"][:Kp, :Kg].astype(np.float64)
    den = tab["npred"][:Kp, None] + tab["ngt"][None, :Kg] - dot
    pred_ind, gt_ind = solve_dense(-(dot / np.maximum(den, 1e-10)))
    agree = cnt = 0
    for p, g in zip(pred_ind, gt_ind):
        if tab["ngt"][g] == 0 or tab["npred"][p] == 0:   # torch.mode of an empty selection raises -> pair skipped
            continue
        agree += int(np.argmax(tab["gt_types"][g]) == np.argmax(tab["pred_types"][p]))
        cnt += 1
    return (torch.tensor([float(agree)], device=T_gt.device) / cnt)[0]
