"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

CPU restatement of the per-shape evaluation the reference driver runs after clustering:
    relaxed_iou_fast                          src/segment_utils.py:609-627
    primitive_type_segment_torch              :509-517
    SIOU_matched_segments[_usecd]             :140-243
    mean_IOU_primitive_segment[_usecd]        :359-495
    compute_type_miou_abc, hungarian_matching :258-277, :300-357
    chamfer_distance                          src/utils.py:273-296
Hungarian solver: the reference calls lapsolver.solve_dense (not installed here; oracle/ref_shim.py maps it to
scipy.optimize.linear_sum_assignment, which this file uses as well).  Pinned by oracle/make_golden_metrics.py against the
unmodified reference functions (tests/golden/metrics.npz).
"""
import numpy as np
import torch
from scipy.optimize import linear_sum_assignment


def remap_types(a):
    """0, 6, 7 -> 9 (closed spline), 8 -> 2 (open spline), in place (:156-164)."""
    a[a == 0] = 9
    a[a == 6] = 9
    a[a == 7] = 9
    a[a == 8] = 2
    return a


def one_hot(labels, maxx=50):
    out = np.zeros((len(labels), maxx), np.float32)
    out[np.arange(len(labels)), np.asarray(labels).astype(np.int64)] = 1.0
    return out


def relaxed_iou_fast(pred, gt):
    """(N,K) one-hot each -> (K,K) float32: dots / (|p| + |g| - dots + 1e-7) (:609-627)."""
    dots = pred.T @ gt
    return dots / (pred.sum(0)[:, None] + gt.sum(0)[None, :] - dots + np.float32(1e-7))


def chamfer_distance(pred, gt):
    """src/utils.py:273-296 for one pair of clouds (n,3), (m,3): (mean of row minima + mean of column minima) / 2."""
    p, g = torch.as_tensor(pred, dtype=torch.float32), torch.as_tensor(gt, dtype=torch.float32)
    diff = torch.sum((p[None, :, :] - g[:, None, :]) ** 2, 2)          # (m, n)
    return float((torch.mean(torch.min(diff, 0)[0]) + torch.mean(torch.min(diff, 1)[0])) / 2.0)


def primitive_type_segment(pred_hot, weights):
    """argmax over types of sum_n pred_hot[n, l] * weights[n, k] (:509-517) -> (K,)"""
    return np.argmax(pred_hot.T.astype(np.float64) @ weights.astype(np.float64), 0)


def siou_matched_segments(target, pred_labels, primitives_pred, primitives, weights, points=None, min_gt=100):
    """SIOU_matched_segments (:140-191; points None, gt segments below 100 points dropped, recall = tp / (tp + fn)) and
    SIOU_matched_segments_usecd (:194-243; every gt segment, recall = share of matched pairs with chamfer / 2 < 0.1
    over the number of gt labels).  Mutates the two type arrays like the reference.  Returns
    (segment_iou, primitive_iou, (rows, cols), [[gt type, pred type], ...], recall)."""
    remap_types(primitives)
    remap_types(primitives_pred)
    cost = relaxed_iou_fast(one_hot(pred_labels), one_hot(target))
    rows, cols = linear_sum_assignment(1.0 - cost)
    prim_pred = primitive_type_segment(one_hot(primitives_pred, 10), np.asarray(weights, np.float32))
    ious, recalls, hits, pairs = [], [], [], []
    recall_pos = 0
    for r, c in zip(rows, cols):
        pi, gi = pred_labels == r, target == c
        if gi.sum() == 0 or pi.sum() == 0:
            continue
        if points is None and gi.sum() < min_gt:
            continue
        tp = np.sum(pi & gi)
        ious.append(tp / (np.sum(pi | gi) + 1e-8))
        if points is None:
            recalls.append(tp / (tp + np.sum(~pi & gi) + 1e-8))
        elif chamfer_distance(points[pi], points[gi]) / 2 < 0.1:
            recall_pos += 1
        g_type = primitives[gi][0]
        p_type = prim_pred[r]
        hits.append(g_type == p_type)
        pairs.append([g_type, p_type])
    recall = np.mean(recalls) if points is None else float(recall_pos) / np.unique(target).shape[0]
    return np.mean(ious), np.mean(hits), (rows, cols), pairs, recall


def compute_type_miou_abc(type_per_point, T_gt, cluster_pred, I_gt):
    """:300-357 with (N,) arrays (the reference carries a leading batch axis of 1): share of Hungarian-matched (pred,
    gt) instance pairs whose modal types agree.  Mutates T_gt / a 1-D type_per_point like the reference."""
    T_pred = np.argmax(type_per_point, -1) if type_per_point.ndim == 2 else type_per_point
    for a in (T_pred, T_gt):
        a[a == 6] = 0
        a[a == 7] = 0
        a[a == 9] = 0
        a[a == 8] = 2
    Wp = one_hot(cluster_pred, int(cluster_pred.max()) + 1)
    Wg = one_hot(I_gt + 1, int(I_gt.max()) + 2)[:, 1:] if I_gt.min() == -1 else one_hot(I_gt, int(I_gt.max()) + 1)
    dot = Wp.T @ Wg
    den = Wp.sum(0)[:, None] + Wg.sum(0)[None, :] - dot
    pred_ind, gt_ind = linear_sum_assignment(-(dot / np.maximum(den, 1e-10)))
    agree = cnt = 0
    for p, g in zip(pred_ind, gt_ind):
        a, b = T_gt[I_gt == g], T_pred[cluster_pred == p]
        if a.size == 0 or b.size == 0:         # torch.mode of an empty tensor raises -> the pair is skipped (:342-355)
            continue
        agree += int(np.bincount(a).argmax() == np.bincount(b).argmax())
        cnt += 1
    return np.float32(agree) / np.float32(cnt)


# --------------------------------------------------------------------------
# the chamfer extension -- src/chamfer_distance/chamfer_distance.cu (CUDA only: it cannot be built for or run in the
# CPU-only build container, so this restatement is NOT pinned by an execution of the reference; it follows the kernel text)
# --------------------------------------------------------------------------
def chamfer_ext_forward(xyz1, xyz2):
    """ChamferDistanceKernel (:6-158), both directions (:149-150): for every point of xyz1 (B,n,3) the squared distance to and
    the index of its nearest point of xyz2 (B,m,3) -- differences candidate - query, x^2 + y^2 + z^2, strict `<` scan in index
    order, i.e. the lowest index on ties -- and the same with the roles swapped.  Returns dist1, dist2, idx1, idx2."""
    import torch

    def one(q, c):
        d = ((c[:, None, :, :] - q[:, :, None, :]) ** 2).sum(-1)          # (B, n, m)
        dist, idx = torch.min(d, dim=2)                                    # first minimum = lowest index
        return dist, idx.to(torch.int32)

    d1, i1 = one(xyz1, xyz2)
    d2, i2 = one(xyz2, xyz1)
    return d1, d2, i1, i2


def chamfer_ext_backward(xyz1, xyz2, g1, g2, idx1, idx2):
    """ChamferDistanceGradKernel (:161-187), both directions: grad_xyz1[j] += 2 g1[j] (p1_j - p2_idx1[j]), the opposite sign
    scattered onto p2_idx1[j]; then the same for (xyz2, g2, idx2)."""
    import torch
    gx1, gx2 = torch.zeros_like(xyz1), torch.zeros_like(xyz2)
    for b in range(xyz1.shape[0]):
        v = 2 * g1[b][:, None] * (xyz1[b] - xyz2[b][idx1[b].long()])
        gx1[b] += v
        gx2[b].index_add_(0, idx1[b].long(), -v)
        v = 2 * g2[b][:, None] * (xyz2[b] - xyz1[b][idx2[b].long()])
        gx2[b] += v
        gx1[b].index_add_(0, idx2[b].long(), -v)
    return gx1, gx2
