"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

CPU restatement of the reference driver's test-time-augmentation loops (generate_predictions_aug.py:238-362) on top of
oracle.sednet_forward: one B = 1 forward per augmented copy, sequentially, exactly as the driver does.  (The driver is a
script, not an importable function: the loops are restated line by line; the forward they call is the oracle's, which
tests/test_oracle_golden.py pins against the reference's SEDNet.)"""
import torch

import oracle as O


def _lp(sd, points, normals, k):
    x = torch.cat([points, normals], 2).permute(0, 2, 1)
    return O.sednet_forward(sd, x, k)[1]


def multi_vote(sd, points, normals, k):
    lp = _lp(sd, points, normals, k)                                   # :226-228
    lp_big = _lp(sd, points * 1.15, normals, k)                        # :239-249
    lp_small = _lp(sd, points * 0.85, normals, k)                      # :251-260
    return (lp + lp_big + lp_small) / 3                                # :262


def fold5drop(sd, points, normals, k, drop_out_num):
    lp = _lp(sd, points, normals, k)
    N = points.shape[1]
    total = torch.zeros_like(lp).flatten()                             # :267
    batch = []
    for i in range(N // drop_out_num):                                 # :269-297
        index = torch.ones(points.shape, dtype=torch.bool)
        index[:, i * drop_out_num:(i + 1) * drop_out_num, :] = False
        pd = points[index].reshape((1, N - drop_out_num, 3))
        nd = normals[index].reshape((1, N - drop_out_num, 3))
        batch.append(_lp(sd, pd, nd, k))
    batch = torch.cat(batch, 0)
    for i in range(N // drop_out_num):                                 # :299-302
        index = torch.ones(lp.shape, dtype=torch.bool)
        index[:, :, i * drop_out_num:(i + 1) * drop_out_num] = False
        total[index.flatten()] += batch[i].flatten()
    return lp + total.reshape(lp.shape)                                # :304


def fold5drop_multi_vote(sd, points, normals, k, drop=2000):
    """:307-362 with the fold size as a parameter (the driver hard-codes 2000 for 10 000 points)."""
    N = points.shape[1]
    angles = [torch.eye(3).unsqueeze(0), torch.tensor([[-1.0, 0, 0], [0, 1.0, 0], [0, 0, -1.0]]).unsqueeze(0)]
    tot = None
    for R in angles:
        nc, pc = torch.bmm(normals, R), torch.bmm(points, R)
        cur = _lp(sd, pc, nc, k)
        total = torch.zeros_like(cur).flatten()
        for i in range(N // drop):
            index = torch.ones(points.shape, dtype=torch.bool)
            index[:, i * drop:(i + 1) * drop, :] = False
            b = _lp(sd, pc[index].reshape((1, N - drop, 3)), nc[index].reshape((1, N - drop, 3)), k)
            index = torch.ones(cur.shape, dtype=torch.bool)
            index[:, :, i * drop:(i + 1) * drop] = False
            total[index.flatten()] += b.flatten()
        cur = cur + total.reshape(cur.shape)
        tot = cur if tot is None else tot + cur
    return tot
