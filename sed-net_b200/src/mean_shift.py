"""MeanShift with the reference's method names and return values (reference src/mean_shift.py:11-185), running on
the fused sm_100a kernels: bandwidth = streaming K-th-nearest select, shift = flash-style Gaussian density step,
nms = on-device histogram + ordered compaction (no N x N matrix, no np.unique round trip)."""
import numpy as np
import torch

from . import _lib

MAX_CENTERS = 512


class _MeanShiftIterations(torch.autograd.Function):
    """new_X = mean_shift_(X, b, iterations) with the backward of csrc/meanshift_bwd.cu.  The forward runs the same kernels as
    inference, one iteration per call so that the positions entering every iteration are kept ((iterations + 1) x N x d
    floats; the reference's autograd keeps ~6 N x N tensors per iteration).  The bandwidth is a constant, as in the reference
    (mean_shift computes it under no_grad); a row whose weights all underflow passes no gradient."""

    @staticmethod
    def forward(ctx, X, bw, iterations, mode):
        N, d = X.shape
        Xc = X.detach().contiguous()
        states = [Xc]
        tmp = torch.empty_like(Xc)
        for _ in range(iterations):
            out = torch.empty_like(Xc)
            _lib.call("sed_ms_shift_from", _lib.ptr(Xc), _lib.ptr(states[-1]), _lib.ptr(bw), 1, N, d, 1, 0, mode,
                      _lib.ptr(out), _lib.ptr(tmp), _lib.stream())
            states.append(out)
        ctx.save_for_backward(bw, *states[:-1])
        ctx.shape = (N, d)
        return states[-1].clone()

    @staticmethod
    def backward(ctx, grad_out):
        bw, *states = ctx.saved_tensors
        N, d = ctx.shape
        X = states[0]
        G = grad_out.contiguous().float()
        dX = torch.zeros_like(X)
        ws = torch.empty(_lib.load().sed_ms_shift_backward_workspace_bytes(1, N), dtype=torch.uint8, device=X.device)
        for Q in reversed(states):
            dQ = torch.empty_like(X)
            _lib.call("sed_ms_shift_backward_step", _lib.ptr(Q), _lib.ptr(X), _lib.ptr(G), _lib.ptr(bw), 1, N, d,
                      _lib.ptr(dQ), _lib.ptr(dX), _lib.ptr(ws), _lib.stream())
            G = dQ
        return dX + G, None, None, None


class MeanShift:
    def __init__(self, prec_mode=None):
        """prec_mode of sed_ms_shift (include/sednet_b200.h): 0 FP32 FFMA (reference operation order), 4 tcgen05 with every
        operand of both legs split into FP16 hi/lo (3 + 3 MMAs: FP32-faithful), 1 the same with single-FP16 weights (3 + 2 MMAs:
        ~2e-6 per iteration against FP32), 3 the same exponent but a single FP16 pass for the weighted mean (3 + 1 MMAs: faster;
        one FP16 rounding of X per term), 2 plain FP16.
        None = $SEDNET_B200_MS_PREC if set, otherwise the most faithful tensor-core mode for the row width: mode 4 up to 128
        columns (partitions identical to the FP32 oracle's wherever the FP32 FFMA kernel's are; modes 1 and 3 are opt-in: faster,
        identical labels on separated clusters, up to 5e-4 off and a flipped point on heavily overlapping ones); mode 3 for
        129..192 columns (the only tensor-core kernel at that width: the 148-column hpnet embedding,
        parity-tested against the oracle in tests/test_gpu_hpnet.py); mode 0 beyond.  Mode 3 is opt-in for 128 columns."""
        import os
        env = os.environ.get("SEDNET_B200_MS_PREC")
        self.prec_mode = prec_mode if prec_mode is not None else (int(env) if env is not None else None)

    def _mode(self, d):
        if self.prec_mode is not None:
            return self.prec_mode
        if d % 4 == 0 and d <= 128:
            return 4
        return 3 if d % 4 == 0 and d <= 192 else 0

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
        if torch.is_grad_enabled() and X.requires_grad and int(iterations) > 0:
            # training (src/segment_loss.py:50-56): the reference differentiates these iterations with autograd
            if kernel_type != "gaussian":
                raise NotImplementedError("the backward kernels cover the gaussian kernel (the one the triplet loss uses)")
            return _MeanShiftIterations.apply(X, bw.detach(), int(iterations), self._mode(d)), X
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

    # -- src/mean_shift.py:98-113 (the N x N kernel matrix itself; mean_shift_ never forms it here, so this is API
    #    parity for callers that want the matrix: plain tensor ops on X's device)
    def kernel(self, X, kernel_type, bw):
        dist = 2.0 - 2.0 * X @ torch.transpose(X, 1, 0)
        if kernel_type == "gaussian":
            return torch.exp(torch.clamp(-dist / (bw ** 2) / 2, min=-75, max=75))
        if kernel_type == "epa":
            return torch.nn.functional.relu(3 / 4 * (1 - dist / (bw ** 2)))
        raise ValueError(f"unknown kernel_type {kernel_type!r} (the reference defines 'gaussian' and 'epa')")

    # -- src/mean_shift.py:115-137
    def compute_bandwidth(self, X, num_samples, quantile):
        X = _lib.require_cuda(X, name="X")
        N, d = X.shape
        L = np.arange(N)
        np.random.shuffle(L)     # always drawn, as in the reference (:125-127): NumPy's global stream stays aligned
        if num_samples < N:      # random row subset; with num_samples >= N the reference only permutes the rows,
            X = X[torch.as_tensor(L[:num_samples], device=X.device)].contiguous()   # which the mean does not see
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
