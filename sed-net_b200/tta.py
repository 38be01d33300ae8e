"""Test-time augmentation of the type prediction, batched.

The reference driver (generate_predictions_aug.py:238-362) runs every augmented copy of the shape as its own B = 1
forward, one after the other: 2 extra forwards for `multi_vote`, 5 for `fold5drop`, 12 for both.  The copies of one
mode have the same number of points, so here they go through the fused forward as ONE batch per point count
(B = 3 at N; B = 5 at N - drop; B = 2 at N + B = 10 at N - drop), on the same kernels as the plain forward.

`model` is a sednet_b200.src.SEDNet.SEDNet (type network); points, normals are (1,N,3) CUDA tensors as in the driver.
Each function returns the driver's `primitives_log_prob` (1,P,N) after the mode's accumulation.
"""
import torch


def _inp(points, normals):
    return torch.cat([points, normals], 2).permute(0, 2, 1).contiguous()          # (B,6,N), :223-225


@torch.no_grad()
def multi_vote(model, points, normals):
    """generate_predictions_aug.py:238-262: mean of the log-probabilities at scales 1, 1.15, 0.85 (points only)."""
    x = torch.cat([_inp(points, normals), _inp(points * 1.15, normals), _inp(points * 0.85, normals)], 0)
    lp = model(x, None, False)[1]
    return ((lp[0:1] + lp[1:2]) + lp[2:3]) / 3


def _fold_inputs(points, normals, drop):
    N = points.shape[1]
    folds, keeps = [], []
    for i in range(N // drop):
        keep = torch.ones(N, dtype=torch.bool, device=points.device)
        keep[i * drop:(i + 1) * drop] = False
        keeps.append(keep)
        folds.append(_inp(points[:, keep], normals[:, keep]))
    return torch.cat(folds, 0), keeps


@torch.no_grad()
def fold5drop(model, points, normals, drop_out_num=2000):
    """generate_predictions_aug.py:264-304: the full forward plus, for every fold i, the forward of the cloud without
    points [i*drop, (i+1)*drop) added onto the points it kept."""
    lp = model(_inp(points, normals), None, False)[1].clone()
    folds, keeps = _fold_inputs(points, normals, drop_out_num)
    lpb = model(folds, None, False)[1]
    total = torch.zeros_like(lp)
    for i, keep in enumerate(keeps):
        total[0][:, keep] += lpb[i]
    return lp + total


@torch.no_grad()
def fold5drop_multi_vote(model, points, normals, drop_out_num=2000):
    """generate_predictions_aug.py:307-362: the fold accumulation (2000-point folds in the driver) for the cloud as is and
    rotated by diag(-1, 1, -1), summed."""
    R = torch.tensor([[-1.0, 0, 0], [0, 1.0, 0], [0, 0, -1.0]], device=points.device).unsqueeze(0)
    views = [(points, normals), (torch.bmm(points, R), torch.bmm(normals, R))]
    full = model(torch.cat([_inp(p, n) for p, n in views], 0), None, False)[1]
    fi = [_fold_inputs(p, n, drop_out_num) for p, n in views]
    lpb = model(torch.cat([f for f, _ in fi], 0), None, False)[1]
    nf = fi[0][0].shape[0]
    out = None
    for v in range(2):
        total = torch.zeros_like(full[v:v + 1])
        for i, keep in enumerate(fi[v][1]):
            total[0][:, keep] += lpb[v * nf + i]
        cur = full[v:v + 1] + total
        out = cur if out is None else out + cur
    return out
