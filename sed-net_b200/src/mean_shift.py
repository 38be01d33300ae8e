"""MeanShift with the reference's method names and return values (reference src/mean_shift.py:11-185), running on
the fused sm_100a kernels: bandwidth = streaming K-th-nearest select, shift = flash-style Gaussian density step,
nms = on-device histogram + ordered compaction (no N x N matrix, no np.unique round trip)."""
import numpy as np
import torch

from . import _lib

MAX_CENTERS = 512


class MeanShift:
    def __init__(self, prec_mode=None):
        """prec_mode of sed_ms_shift (include/sednet_b200.h): 0 FP32 FFMA, 1 / 3 tcgen05 FP16 hi/lo split (3+2 / 3+1 MMAs),
        2 plain FP16; None = $SEDNET_B200_MS_PREC if set, otherwise automatic: the FP32-faithful tensor-core mode 3
        (embeddings up to 192 wide, zero-padded to 128 or 192; identical labels, <= 4e-6 from mode 0), else the FFMA kernel."""
        import os
        env = os.environ.get("SEDNET_B200_MS_PREC")
        self.prec_mode = prec_mode if prec_mode is not None else (int(env) if env is not None else None)

    def _mode(self, d):
        if self.prec_mode is not None:
            return self.prec_mode
        return 3 if d <= 192 and d % 4 == 0 else 0

    # -- src/mean_shift.py:19-43
    def mean_shift(self, X, num_samples, quantile, iterations, kernel_type="gaussian", bw=None, nms=True):
        X = _lib.require_cuda(X, name="X")
        if bw is None:
            with torch.no_grad():
                bw = self.compute_bandwidth(X, num_samples, quantile)
                bw = torch.clamp(bw, min=0.003)
        new_X, _ = self.mean_shift_(X, b=bw, iterations=iterations, kernel_type=kernel_type)
        if nms:
            with torch.no_grad():
                _, indices, new_labels = self.nms(new_X, X, b=bw)
            center = new_X[indices]
            return new_X, center, bw, new_labels
        return new_X, bw

    # -- src/mean_shift.py:45-79
    def mean_shift_(self, X, b, iterations=10, kernel_type="gaussian"):
        X = _lib.require_cuda(X, name="X")
        N, d = X.shape
        bw = torch.as_tensor(b, dtype=torch.float32, device=X.device).reshape(1).contiguous()
        out, tmp = torch.empty_like(X), torch.empty_like(X)
        kt = 0 if kernel_type == "gaussian" else 1
        _lib.call("sed_ms_shift", _lib.ptr(X), _lib.ptr(bw), 1, N, d, int(iterations), kt, self._mode(d),
                  _lib.ptr(out), _lib.ptr(tmp), _lib.stream())
        return out, X

    # -- src/mean_shift.py:81-96 (class-level guard: 5000 samples, quantile doubled)
    def guard_mean_shift(self, embedding, quantile, iterations, kernel_type="gaussian"):
        while True:
            _, center, bandwidth, cluster_ids = self.mean_shift(embedding, 5000, quantile, iterations,
                                                                kernel_type=kernel_type)
            if torch.unique(cluster_ids).shape[0] > 49:
                quantile *= 2
            else:
                break
        return center, bandwidth, cluster_ids

    # -- src/mean_shift.py:98-113
    def kernel(self, X, kernel_type, bw):
        if kernel_type == "gaussian":
            return torch.exp(torch.clamp(-X / (bw ** 2) / 2, min=-75, max=75))
        return torch.relu(3 / 4 * (1 - X / (bw ** 2)))

    # -- src/mean_shift.py:115-137
    def compute_bandwidth(self, X, num_samples, quantile):
        X = _lib.require_cuda(X, name="X")
        N, d = X.shape
        if num_samples < N:  # random row subset, np RNG as in the reference (:125-128)
            L = np.arange(N)
            np.random.shuffle(L)
            X = X[torch.as_tensor(L[:num_samples], device=X.device)].contiguous()
            N = num_samples
        K = int(quantile * num_samples)
        kth = torch.empty(N, dtype=torch.float32, device=X.device)
        bw = torch.empty(1, dtype=torch.float32, device=X.device)
        _lib.call("sed_ms_bandwidth", _lib.ptr(X), 1, N, d, K, 0.0, _lib.ptr(kth), _lib.ptr(bw), _lib.stream())
        return bw[0]

    # -- src/mean_shift.py:139-179
    def nms(self, centers, X, b):
        centers = _lib.require_cuda(centers, name="centers")
        X = _lib.require_cuda(X, name="X")
        N, d = X.shape
        dev = X.device
        bw = torch.as_tensor(b, dtype=torch.float32, device=dev).reshape(1).contiguous()
        labels = torch.empty(N, dtype=torch.int64, device=dev)
        ids = torch.empty(MAX_CENTERS, dtype=torch.int32, device=dev)
        counts = torch.empty(2, dtype=torch.int32, device=dev)
        cen = torch.empty((MAX_CENTERS, d), dtype=torch.float32, device=dev)
        ws = torch.empty(_lib.load().sed_ms_nms_workspace_bytes(1, N), dtype=torch.uint8, device=dev)
        _lib.call("sed_ms_nms", _lib.ptr(centers), _lib.ptr(X), _lib.ptr(bw), 1, N, d, MAX_CENTERS, _lib.ptr(labels),
                  _lib.ptr(ids), _lib.ptr(counts[0:1]), _lib.ptr(counts[1:2]), _lib.ptr(cen), _lib.ptr(ws),
                  _lib.stream())
        n = int(counts[0].item())  # the reference returns tensors sized by the cluster count: one host sync
        if n < 0:
            raise RuntimeError(f"mean-shift produced more than {MAX_CENTERS} cluster centres")
        return cen[:n].clone(), ids[:n].to(torch.int64), labels

    # -- src/mean_shift.py:181-185
    def pdist(self, x, y):
        x = x.unsqueeze(1)
        y = y.unsqueeze(0)
        return torch.sum((x - y) ** 2, 2)
