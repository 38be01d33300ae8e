"""ctypes binding of libsednet_b200.so (C ABI: include/sednet_b200.h).

The library is the product: if it is missing, or a tensor is not a contiguous CUDA tensor of the expected dtype,
the call raises -- nothing here falls back to PyTorch or the CPU.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SEDNET_B200_LIB") or os.path.join(os.path.dirname(_HERE), "libsednet_b200.so")   # override: A/B builds

c_f32p, c_f64p, c_i32p, c_i64p, c_vp = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
I, F, L, D = C.c_int, C.c_float, C.c_int64, C.c_double

# name -> (restype, argtypes); order and meaning as in include/sednet_b200.h
SIGNATURES = {
    "sed_version": (I, []),
    "sed_error_string": (C.c_char_p, [I]),
    "sed_knn_l2": (I, [c_f32p, I, I, I, I, c_vp, I, c_vp]),
    "sed_knn_pn": (I, [c_f32p, I, I, I, F, c_vp, I, c_vp]),
    "sed_graph_feature": (I, [c_f32p, c_i64p, I, I, I, I, c_f32p, c_vp]),
    "sed_sednet_workspace_bytes": (L, [I, I, I]),
    "sed_sednet_forward": (I, [C.POINTER(C.c_void_p), c_f32p, I, I, I, F, F, I, I, c_f32p, c_f32p, c_f32p, c_f32p,
                               c_f32p, c_vp, L, c_vp]),
    "sed_sednet_forward_g1": (I, [C.POINTER(C.c_void_p), c_f32p, c_i32p, I, I, I, F, F, I, I, c_f32p, c_f32p, c_f32p,
                                  c_f32p, c_f32p, c_vp, L, c_vp]),
    "sed_encoder_forward": (I, [C.POINTER(C.c_void_p), c_f32p, I, I, I, F, c_f32p, c_f32p, c_vp, L, c_vp]),
    "sed_edgeconv_workspace_bytes": (L, [I, I, I]),
    "sed_edgeconv_forward": (I, [c_f32p, L, c_i32p, c_f32p, c_f32p, c_f32p, I, I, I, I, I, I, F, F, c_f32p, L, c_vp,
                                 c_vp]),
    "sed_pointwise_forward": (I, [c_f32p, L, c_f32p, I, c_f32p, c_f32p, c_f32p, I, c_f32p, L, c_f64p, c_f32p, I, I, I, I,
                                  c_vp]),
    "sed_normalize_transpose": (I, [c_f32p, I, I, I, c_f32p, c_vp]),
    "sed_ms_bandwidth": (I, [c_f32p, I, I, I, I, F, c_f32p, c_f32p, c_vp]),
    "sed_ms_shift": (I, [c_f32p, c_f32p, I, I, I, I, I, I, c_f32p, c_f32p, c_vp]),
    "sed_ms_shift_from": (I, [c_f32p, c_f32p, c_f32p, I, I, I, I, I, I, c_f32p, c_f32p, c_vp]),
    "sed_ms_shift_backward_workspace_bytes": (L, [I, I]),
    "sed_ms_shift_backward_step": (I, [c_f32p, c_f32p, c_f32p, c_f32p, I, I, I, c_f32p, c_f32p, c_vp, c_vp]),
    "sed_ms_nms_workspace_bytes": (L, [I, I]),
    "sed_ms_nms": (I, [c_f32p, c_f32p, c_f32p, I, I, I, I, c_i64p, c_i32p, c_i32p, c_i32p, c_f32p, c_vp, c_vp]),
    "sed_one_hot": (I, [c_i64p, I, I, c_f32p, c_vp]),
    "sed_fit_segments": (I, [c_f32p, c_f32p, c_f32p, c_i64p, c_i32p, I, I, I, I, c_f32p, c_i32p, c_vp]),
    "sed_fit_segments_v2": (I, [c_f32p, c_f32p, c_f32p, c_i64p, c_i32p, I, I, I, I, D, c_f32p, c_i32p, c_vp]),
    "sed_three_nn": (I, [c_f32p, c_f32p, I, I, I, c_f32p, c_i32p, c_vp]),
    "sed_inst_edges": (I, [c_i32p, c_i64p, I, I, c_vp, c_vp]),
    "sed_face_face_map": (I, [c_f32p, c_i64p, c_i32p, c_i64p, I, I, I, c_vp, c_vp]),
    "sed_compute_entropy": (I, [c_f32p, I, I, I, c_f32p, c_vp]),
    "sed_far_idx": (I, [c_f32p, I, I, I, c_i32p, c_vp]),
    "sed_affinity_normal_build": (I, [c_f32p, c_i32p, I, I, I, F, c_f32p, c_f32p, c_vp]),
    "sed_affinity_workspace_bytes": (L, [I, I]),
    "sed_affinity_prepare": (I, [c_i32p, c_f32p, I, I, c_vp, c_vp]),
    "sed_affinity_matmul": (I, [c_i32p, c_f32p, c_f32p, c_vp, c_f32p, I, I, I, c_f32p, c_vp]),
    "sed_segment_tables": (I, [c_i64p, c_i64p, c_i64p, c_i64p, I, I, I, I, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p,
                               c_vp]),
    "sed_type_vote_weighted": (I, [c_i64p, c_f32p, I, I, I, c_f32p, c_vp]),
    "sed_chamfer_min": (I, [c_f32p, c_f32p, I, I, I, c_f32p, c_f32p, c_vp]),
    "sed_chamfer_forward": (I, [c_f32p, c_f32p, I, I, I, c_f32p, c_f32p, c_i32p, c_i32p, c_vp]),
    "sed_chamfer_backward": (I, [c_f32p, c_f32p, c_f32p, c_f32p, c_i32p, c_i32p, I, I, I, c_f32p, c_f32p, c_vp]),
    "sed_matched_chamfer": (I, [c_f32p, c_i64p, c_i64p, c_i32p, c_i32p, I, I, I, c_f32p, c_f32p, c_vp]),
    "sed_lstsq3": (I, [c_f32p, c_f32p, I, c_f32p, c_i32p, c_vp]),
    "sed_svd3": (I, [c_f32p, I, c_f32p, c_f32p, c_vp]),
    "sed_svd3_backward": (I, [c_f32p, c_f32p, c_f32p, c_f32p, I, c_f32p, c_vp]),
    "sed_residual_segments": (I, [c_f32p, c_i64p, c_i32p, c_f32p, c_i32p, I, I, I, I, c_f32p, c_vp]),
    "sed_primitive_distance": (I, [c_f32p, I, I, c_f32p, I, c_f32p, c_vp]),
    "sed_segment_types": (I, [c_f32p, c_i64p, I, I, I, I, c_i32p, c_i32p, c_i32p, c_vp]),
    "sed_pipeline_create": (I, [I, I, I, I, C.POINTER(C.c_void_p)]),
    "sed_pipeline_destroy": (None, [c_vp]),
    "sed_pipeline_set_weights": (I, [c_vp, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "sed_pipeline_run_host": (I, [c_vp, c_vp, c_vp, I, D, I, I, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "sed_pipeline_run_device": (I, [c_vp, c_f32p, c_f32p, I, D, I, I, c_vp]),
    "sed_pipeline_run_forward": (I, [c_vp, c_f32p, c_f32p, I, c_vp]),
    "sed_pipeline_run_cluster": (I, [c_vp, c_f32p, c_f32p, I, D, I, I, c_vp]),
    "sed_pipeline_set_cluster_width": (I, [c_vp, I]),
    "sed_pipeline_device_ptr": (c_vp, [c_vp, C.c_char_p]),
    "sed_pipeline_stage_ms": (I, [c_vp, C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "sed_launch_count": (L, [I]),
}

# order of enum sed_param -> state_dict key (src/SEDNet.py parameter names)
PARAM_KEYS = [
    "encoder.conv1.0.weight", "encoder.bn1.weight", "encoder.bn1.bias",
    "encoder.conv2.0.weight", "encoder.bn2.weight", "encoder.bn2.bias",
    "encoder.conv3.0.weight", "encoder.bn3.weight", "encoder.bn3.bias",
    "encoder.mlp1.weight", "encoder.mlp1.bias", "encoder.bnmlp1.weight", "encoder.bnmlp1.bias",
    "conv1.weight", "conv1.bias", "bn1.weight", "bn1.bias",
    "conv2.weight", "conv2.bias", "bn2.weight", "bn2.bias",
    "mlp_prim_prob1.weight", "mlp_prim_prob1.bias", "bn_prim_prob1.weight", "bn_prim_prob1.bias",
    "mlp_prim_prob2.weight", "mlp_prim_prob2.bias",
    "edge_module.0.weight", "edge_module.0.bias", "edge_module.1.weight", "edge_module.1.bias",
    "edge_module.2.weight", "edge_module.2.bias",
    "mlp_seg_prob1.weight", "mlp_seg_prob1.bias", "bn_seg_prob1.weight", "bn_seg_prob1.bias",
    "asis.0.weight", "asis.0.bias", "asis.1.weight", "asis.1.bias",
    "prim_encoding.0.weight", "prim_encoding.0.bias",
    "mlp_seg_prob2.weight", "mlp_seg_prob2.bias",
]

_lib = None


def load():
    """dlopen the library and declare every signature. Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python sed-net_b200/build.py` "
                               "(there is no CPU or PyTorch fallback for the hot path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = load().sed_error_string(int(rc)).decode()
        raise RuntimeError(f"libsednet_b200 {what} failed: {msg} (code {rc})")


def call(name, *args):
    check(getattr(load(), name)(*args), name)


def require_cuda(t, dtype=torch.float32, name="tensor"):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the sm_100a kernels are the only implementation of this path")
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def param_table(state, device, prefix="", count=None):
    """SED_P_COUNT device pointers (ctypes array) + the tensors that must stay alive.  With `prefix`, keys are looked
    up without it (an encoder module's own parameter names) and only the first `count` entries are filled."""
    keep = []
    arr = (C.c_void_p * len(PARAM_KEYS))()
    for i, k in enumerate(PARAM_KEYS[:count]):
        t = state[k[len(prefix):]].detach()
        if t.device != device or t.dtype != torch.float32 or not t.is_contiguous():
            t = t.to(device=device, dtype=torch.float32).contiguous()
        keep.append(t)
        arr[i] = t.data_ptr()
    return arr, keep
