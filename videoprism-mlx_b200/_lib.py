"""ctypes binding of libvideoprism_b200.so (the C ABI in include/videoprism_b200.h)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvideoprism_b200.so")

VP_OK, VP_ERR_INVALID, VP_ERR_KEY, VP_ERR_INCOMPLETE, VP_ERR_CUDA, VP_ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5
VP_F32, VP_BF16, VP_I32, VP_U8 = 0, 1, 2, 3
VP_KIND_ENCODER, VP_KIND_CLIP, VP_KIND_CLASSIFIER = 0, 1, 2
VP_FLAG_CHECK_FP32 = 1


class VpConfig(C.Structure):
    _fields_ = [
        ("kind", C.c_int), ("patch_size", C.c_int),
        ("pos_emb_t", C.c_int), ("pos_emb_h", C.c_int), ("pos_emb_w", C.c_int),
        ("model_dim", C.c_int), ("num_spatial_layers", C.c_int), ("num_temporal_layers", C.c_int),
        ("num_heads", C.c_int), ("mlp_dim", C.c_int), ("atten_logit_cap", C.c_float),
        ("num_auxiliary_layers", C.c_int), ("num_unimodal_layers", C.c_int), ("vocabulary_size", C.c_int),
        ("num_classes", C.c_int), ("text_norm_policy", C.c_int),
    ]


_P = C.c_void_p
_I = C.c_int
_PROTOS = {
    "vp_create": (_I, [C.POINTER(VpConfig), C.POINTER(_P)]),
    "vp_create_on_device": (_I, [C.POINTER(VpConfig), _I, C.POINTER(_P)]),
    "vp_create_ex": (_I, [C.POINTER(VpConfig), _I, C.c_uint, C.POINTER(_P)]),
    "vp_handle_flags": (_I, [_P]),
    "vp_handle_device": (_I, [_P]),
    "vp_destroy": (None, [_P]),
    "vp_last_error": (C.c_char_p, [_P]),
    "vp_set_weight": (_I, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), _I]),
    "vp_num_weights": (_I, [_P]),
    "vp_weight_key": (C.c_char_p, [_P, _I]),
    "vp_weight_ndim": (_I, [_P, _I]),
    "vp_weight_dim": (C.c_int64, [_P, _I, _I]),
    "vp_finalize": (_I, [_P]),
    "vp_encoder_forward": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _I, _P]),
    "vp_encoder_forward_host": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P]),
    "vp_encoder_forward_u8": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _I, _P]),
    "vp_encoder_forward_host_u8": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P]),
    "vp_encoder_forward_host_async": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _I, _P, C.POINTER(C.c_uint64)]),
    "vp_wait": (_I, [_P, C.c_uint64]),
    "vp_clip_video_forward_u8": (_I, [_P, _P, _I, _I, _I, _I, _P, _I, _P, _P, _P, _P, _P]),
    "vp_clip_video_forward_host_async": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _I, _P, _P, C.POINTER(C.c_uint64)]),
    "vp_clip_video_forward": (_I, [_P, _P, _I, _I, _I, _I, _P, _I, _P, _P, _P, _P, _P]),
    "vp_clip_text_forward": (_I, [_P, _P, _P, _I, _I, _I, _P, _P]),
    "vp_clip_video_forward_host": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "vp_clip_text_forward_host": (_I, [_P, _P, _P, _I, _I, _I, _P, _P]),
    "vp_classifier_forward": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "vp_similarity": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "vp_workspace_bytes": (C.c_size_t, [_P, _I, _I, _I, _I]),
    "vp_release_workspace": (_I, [_P]),
    "vp_kernel_launches": (C.c_int64, [_P]),
    "vp_device_sm_count": (_I, []),
    "vp_trace": (_I, [_P, _I]),
    "vp_trace_report": (_I, [_P, C.c_char_p, _I]),
    "vp_gemm_bf16": (_I, [_P, _I, _P, _I, _P, _I, _I, _I, _I, _P, _I, _P, _I, _P, _P, _I, _I, _P]),
    "vp_gemm_bf16_ln": (_I, [_P, _I, _P, _I, _P, _I, _I, _I, _I, _P, _I, _P, _I, _P, _I, _P, _I, _P, _P]),
    "vp_gemm_stats_slots": (_I, [_I]),
    "vp_row_stats": (_I, [_P, _I, _P, _I, _I, _P]),
    "vp_fold_ln_weight": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, C.c_float, _P]),
    "vp_layernorm": (_I, [_P, _I, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "vp_patchify": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "vp_resize_frames_u8": (_I, [_P, _I, _I, _I, _P, _I, _I, _P]),
    "vp_attention": (_I, [_P, _P, _P, _I, _P, _I, _I, _I, _I, _I, _I, C.c_float, _P, _I, _P]),
}

EXPORTED_SYMBOLS = tuple(_PROTOS)
_lib = None


def lib() -> C.CDLL:
    """Loads the shared library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python videoprism-mlx_b200/build.py` "
                "(or __graft_entry__.build()).  There is no CPU / PyTorch fallback for this path.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


class VpError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"videoprism_b200 error {code}: {msg}")
        self.code = code


def check(code: int, handle=None) -> None:
    if code == VP_OK:
        return
    msg = lib().vp_last_error(handle)
    msg = msg.decode() if msg else ""
    if code == VP_ERR_INVALID:
        raise ValueError(msg or "invalid argument")
    if code == VP_ERR_KEY:
        raise KeyError(msg)
    raise VpError(code, msg)
