"""Data-parallel sharding of a batch of shapes across ranks (one process per GPU).

Shapes are independent (the reference's driver already loops over shapes one at a time,
generate_predictions_aug.py:213; GroupNorm has no batch statistics), so ranks take contiguous slices of the batch
and the only exchange is one all-gather of fixed-size per-shape records at the end (NCCL on GPUs, gloo in the CPU
tests)."""
import torch
import torch.distributed as dist

RECORD_FIELDS = ("shape_id", "n_labels", "n_fitted", "mean_residual", "bandwidth", "label_checksum")


def shard_range(total, rank, world):
    """Contiguous slice [lo, hi) of `total` shapes owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


MAX_LABELS = 512    # upper bound of a label value (the nms keeps at most 512 centres)


def canonical_labels(labels):
    """(B,N) int64 labels -> the same partitions numbered by first occurrence (device ops, no host sync).  Which of several
    numerically identical converged points becomes a centre -- and with it the numbering -- may change with the batch a cloud
    is processed in (the mean-shift kernel's CTA decomposition depends on the batch); the partition does not."""
    B, N = labels.shape
    pos = torch.arange(N, device=labels.device).expand(B, N)
    first = torch.full((B, MAX_LABELS), N, dtype=torch.int64, device=labels.device).scatter_reduce(1, labels, pos, "amin")
    rank = first.argsort(dim=1, stable=True).argsort(dim=1, stable=True)
    return rank.gather(1, labels)


def make_records(shape_ids, n_labels, status, residual, bw, labels):
    """(B, len(RECORD_FIELDS)) float64 records from the pipeline outputs of the local shard (tensors on one device);
    label_checksum is a position-weighted sum of the canonically numbered labels (equal partitions <=> equal checksums,
    up to collisions)."""
    fitted = (status != 1)
    nf = fitted.sum(1)
    mean_res = (residual * fitted).sum(1) / nf.clamp(min=1)
    lab = canonical_labels(labels)
    chk = (lab.to(torch.float64) * (torch.arange(labels.shape[1], device=labels.device, dtype=torch.float64) % 97 + 1)).sum(1)
    cols = [shape_ids, n_labels, nf, mean_res, bw, chk]
    return torch.stack([c.to(torch.float64) for c in cols], 1)


def gather_records(rec, max_per_rank):
    """All-gather of per-shape records; every rank pads its shard to `max_per_rank` rows (shape_id -1 = padding).
    Returns the (n_shapes, F) table sorted by shape id on every rank."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    pad = torch.full((max_per_rank, rec.shape[1]), -1.0, dtype=rec.dtype, device=rec.device)
    pad[: rec.shape[0]] = rec
    if world == 1:
        out = pad
    else:
        out = torch.empty((world * max_per_rank, rec.shape[1]), dtype=rec.dtype, device=rec.device)
        dist.all_gather_into_tensor(out, pad)
    out = out[out[:, 0] >= 0]
    return out[torch.argsort(out[:, 0])]
