"""End-to-end inference step over a batch of clouds (two forwards, type argmax, normalise, guarded mean-shift,
per-segment type vote, primitive fits, residuals) behind the C-ABI pipeline handle of libsednet_b200.so.

Reference flow: generate_predictions_aug.py:213-236,365,379-387 + Fitting_patches_and_edges/residual_utils.py:210-331.
``run_host`` is the host-buffer entry (H2D and D2H copies inside the call); ``run_device`` keeps inputs and results
resident in HBM.
"""
import ctypes as C

import numpy as np
import torch

from .src import _lib


def _host_table(sd):
    keep, arr = [], (C.c_void_p * len(_lib.PARAM_KEYS))()
    for i, k in enumerate(_lib.PARAM_KEYS):
        v = sd[k]
        a = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
        a = np.ascontiguousarray(a, dtype=np.float32)
        keep.append(a)
        arr[i] = a.ctypes.data
    return arr, keep


class Pipeline:
    def __init__(self, max_batch, n_points, k=64, max_segments=64):
        if not torch.cuda.is_available():
            raise RuntimeError("sednet_b200.Pipeline needs a CUDA device (no CPU fallback)")
        self.B, self.N, self.k, self.S = max_batch, n_points, k, max_segments
        self._h = C.c_void_p()
        _lib.call("sed_pipeline_create", max_batch, n_points, k, max_segments, C.byref(self._h))
        pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
        B, N, S = max_batch, n_points, max_segments
        self.out = dict(labels=pin((B, N), torch.int64), pred_type=pin((B, N), torch.int32),
                        seg_type=pin((B, S), torch.int32), params=pin((B, S, 8), torch.float32),
                        status=pin((B, S), torch.int32), residual=pin((B, S), torch.float32),
                        bw=pin((B,), torch.float32), n_labels=pin((B,), torch.int32))

    def close(self):
        if self._h:
            _lib.load().sed_pipeline_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_weights(self, sd_type, sd_inst):
        """Two state_dicts (name -> array/tensor; keys as in the reference's SEDNet): type network and instance
        network (generate_predictions_aug.py:142-170,191-198)."""
        t, k1 = _host_table(sd_type)
        i, k2 = _host_table(sd_inst)
        _lib.call("sed_pipeline_set_weights", self._h, t, i)

    def run_host(self, points, normals, quantile=0.015, iterations=50, prec_mode=4):
        """points, normals: (B,N,3) float32 HOST tensors (pinned for full copy speed). Returns dict of host tensors
        (views into pinned result buffers, valid until the next call)."""
        B = points.shape[0]
        assert points.shape == (B, self.N, 3) and normals.shape == (B, self.N, 3) and B <= self.B
        assert points.dtype == torch.float32 and not points.is_cuda and points.is_contiguous()
        assert normals.dtype == torch.float32 and not normals.is_cuda and normals.is_contiguous()
        o = self.out
        _lib.call("sed_pipeline_run_host", self._h, C.c_void_p(points.data_ptr()), C.c_void_p(normals.data_ptr()), B,
                  float(quantile), int(iterations), int(prec_mode), _lib.ptr(o["labels"]), _lib.ptr(o["pred_type"]),
                  _lib.ptr(o["seg_type"]), _lib.ptr(o["params"]), _lib.ptr(o["status"]), _lib.ptr(o["residual"]),
                  _lib.ptr(o["bw"]), _lib.ptr(o["n_labels"]), _lib.stream())
        return {k: v[:B] for k, v in o.items()}

    def run_device(self, points, normals, quantile=0.015, iterations=50, prec_mode=4):
        """points, normals: (B,N,3) float32 CUDA tensors; results stay on the device (see ``device_tensor``)."""
        points = _lib.require_cuda(points, name="points")
        normals = _lib.require_cuda(normals, name="normals")
        B = points.shape[0]
        assert points.shape == (B, self.N, 3) and B <= self.B
        _lib.call("sed_pipeline_run_device", self._h, _lib.ptr(points), _lib.ptr(normals), B, float(quantile),
                  int(iterations), int(prec_mode), _lib.stream())

    def run_forward(self, points, normals):
        """First half of run_device: both networks; leaves X = normalised embedding (B,N,128) in the handle
        (``device_tensor_view("X")`` is writable: a driver may post-process the embedding before clustering)."""
        points = _lib.require_cuda(points, name="points")
        normals = _lib.require_cuda(normals, name="normals")
        B = points.shape[0]
        assert points.shape == (B, self.N, 3) and B <= self.B
        _lib.call("sed_pipeline_run_forward", self._h, _lib.ptr(points), _lib.ptr(normals), B, _lib.stream())
        self._d = 128

    def set_cluster_embedding(self, X):
        """Replace the embedding the clustering half reads: X (B,N,d) float32 CUDA, unit rows, d a multiple of 4 up to 192 (the
        reference's hpnet_process concatenation is 148 wide, generate_predictions_aug.py:371-380).  run_forward resets the
        handle to the network's own 128-wide embedding."""
        X = _lib.require_cuda(X, name="X")
        B, N, d = X.shape
        assert N == self.N and B <= self.B
        _lib.call("sed_pipeline_set_cluster_width", self._h, int(d))
        self._d = int(d)
        self.device_tensor_view("X", width=d)[:B].copy_(X)

    def run_cluster(self, points, normals, quantile=0.015, iterations=50, prec_mode=None):
        """Second half: guarded mean-shift of the handle's X, type vote, fits, residuals (results on the device).
        prec_mode None = 4 for rows up to 128 wide, 3 for wider ones (the tensor-core kernel of 129..192 columns; modes 1 and 4
        would take the FP32 FFMA kernel there, ~30x slower)."""
        if prec_mode is None:
            prec_mode = 4 if getattr(self, "_d", 128) <= 128 else 3
        points = _lib.require_cuda(points, name="points")
        normals = _lib.require_cuda(normals, name="normals")
        B = points.shape[0]
        assert points.shape == (B, self.N, 3) and B <= self.B
        _lib.call("sed_pipeline_run_cluster", self._h, _lib.ptr(points), _lib.ptr(normals), B, float(quantile),
                  int(iterations), int(prec_mode), _lib.stream())

    STAGES = ("graph1", "forwards", "bandwidth", "shift", "nms", "fit")

    def stage_ms(self):
        """Device time of each stage of the last run (CUDA events on the run's stream) and the guard retries."""
        ms, r = (C.c_float * 6)(), C.c_int()
        _lib.call("sed_pipeline_stage_ms", self._h, ms, C.byref(r))
        return dict(zip(self.STAGES, [float(v) for v in ms])), int(r.value)

    _SHAPES = dict(labels=("BN", torch.int64), pred_type=("BN", torch.int32), seg_type=("BS", torch.int32),
                   params=("BS8", torch.float32), status=("BS", torch.int32), residual=("BS", torch.float32),
                   bw=("B", torch.float32), n_labels=("B", torch.int32), n_centers=("B", torch.int32),
                   X=("BNd", torch.float32), shifted=("BNd", torch.float32), embedding=("BdN", torch.float32),
                   log_prob=("B6N", torch.float32), type_log_prob=("B6N", torch.float32))

    def device_tensor_view(self, name, width=None):
        """Zero-copy torch view of a named device buffer of the handle (valid for the handle's lifetime).  X / shifted are
        dense (B,N,d) at the current clustering width (128 unless set_cluster_embedding changed it; `width` overrides)."""
        code, dt = self._SHAPES[name]
        dims = dict(B=self.B, N=self.N, S=self.S, d=width or getattr(self, "_d", 128))
        shape = tuple(dims[c] if c in dims else int(c) for c in code)
        p = _lib.load().sed_pipeline_device_ptr(self._h, name.encode())
        if not p:
            raise KeyError(name)
        typestr = {torch.float32: "<f4", torch.int32: "<i4", torch.int64: "<i8"}[dt]

        class _Buf:
            __cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (int(p), False), "version": 2}
        return torch.as_tensor(_Buf(), device=torch.device("cuda", torch.cuda.current_device()))

    def device_tensor(self, name):
        """Copy of a named device buffer (debug / parity checks)."""
        torch.cuda.synchronize()
        return self.device_tensor_view(name).clone()


def launches(reset=False):
    """Number of kernels the library has launched (since the last reset)."""
    return int(_lib.load().sed_launch_count(1 if reset else 0))
