"""ctypes binding of libsvt_b200.so (include/svt_b200.h).  Fails loudly: there is no fallback path."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SVT_B200_LIB points the binding at another build of the same library (A/B measurements of two builds on one box,
# tools/ab_step.py); the default is the in-tree build next to this file.
LIB_PATH = os.environ.get("SVT_B200_LIB") or os.path.join(_HERE, "libsvt_b200.so")

MAX_CONV_LAYERS = 8


class EncoderConfig(C.Structure):
    _fields_ = [
        ("hidden_size", C.c_int), ("num_layers", C.c_int), ("num_heads", C.c_int), ("ffn_size", C.c_int),
        ("num_conv_layers", C.c_int), ("conv_dim", C.c_int),
        ("conv_kernel", C.c_int * MAX_CONV_LAYERS), ("conv_stride", C.c_int * MAX_CONV_LAYERS),
        ("conv_bias", C.c_int), ("feat_norm_layer", C.c_int), ("stable_layer_norm", C.c_int),
        ("pos_conv_kernel", C.c_int), ("pos_conv_groups", C.c_int), ("layer_norm_eps", C.c_float),
        ("normalize_wav", C.c_int), ("output_norm", C.c_int), ("feat_proj_norm", C.c_int),
        ("pos_conv_layers", C.c_int), ("rel_pos_buckets", C.c_int), ("rel_pos_max_distance", C.c_int),
        ("pos_conv_batch_norm", C.c_int),
    ]


class FusionConfig(C.Structure):
    _fields_ = [("d_model", C.c_int), ("nhead", C.c_int), ("d_ffn", C.c_int), ("alpha", C.c_float)]


class VideoConfig(C.Structure):
    _fields_ = [("embed_dim", C.c_int), ("num_layers", C.c_int), ("num_heads", C.c_int), ("ffn_size", C.c_int),
                ("conv_pos", C.c_int), ("conv_pos_groups", C.c_int), ("layer_norm_eps", C.c_float),
                ("input_norm", C.c_int), ("output_norm", C.c_int)]


class SvtError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libsvt_b200 error {code}: {msg}")
        self.code = code


_lib = None

_P = C.c_void_p
_SIGS = {
    "svt_version": (C.c_int, []),
    "svt_last_error": (C.c_char_p, []),
    "svt_device_count": (C.c_int, []),
    "svt_debug_launch_count": (C.c_longlong, []),
    "svt_set_option": (C.c_int, [C.c_char_p, C.c_int]),
    "svt_debug_attention_trace": (None, [_P]),
    "svt_encoder_create": (C.c_int, [C.POINTER(EncoderConfig), C.POINTER(_P)]),
    "svt_encoder_destroy": (None, [_P]),
    "svt_encoder_set_tensor": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int, C.c_int]),
    "svt_encoder_set_head": (C.c_int, [_P, _P, _P, C.c_int]),
    "svt_encoder_finalize": (C.c_int, [_P]),
    "svt_encoder_set_norm_per_clip": (C.c_int, [_P, C.c_int]),
    "svt_encoder_num_frames": (C.c_int, [_P, C.c_int]),
    "svt_encoder_workspace_bytes": (C.c_size_t, [_P, C.c_int, C.c_int]),
    "svt_encoder_forward": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, C.c_size_t, _P, _P, _P]),
    "svt_encoder_forward_host": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, C.c_size_t, _P, _P, _P, _P]),
    "svt_pipeline_create": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, C.c_size_t, _P, _P, C.POINTER(_P)]),
    "svt_pipeline_destroy": (None, [_P]),
    "svt_pipeline_submit": (C.c_int, [_P, _P, _P, C.POINTER(C.c_longlong)]),
    "svt_pipeline_wait": (C.c_int, [_P, C.c_longlong]),
    "svt_fusion_create": (C.c_int, [C.POINTER(FusionConfig), C.POINTER(_P)]),
    "svt_fusion_destroy": (None, [_P]),
    "svt_fusion_set_tensor": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int, C.c_int]),
    "svt_fusion_finalize": (C.c_int, [_P]),
    "svt_fusion_workspace_bytes": (C.c_size_t, [_P, C.c_int, C.c_int]),
    "svt_fusion_forward": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, _P, C.c_size_t, _P, _P]),
    "svt_video_create": (C.c_int, [C.POINTER(VideoConfig), C.POINTER(_P)]),
    "svt_video_destroy": (None, [_P]),
    "svt_video_set_tensor": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int, C.c_int]),
    "svt_video_finalize": (C.c_int, [_P]),
    "svt_video_workspace_bytes": (C.c_size_t, [_P, C.c_int, C.c_int]),
    "svt_video_forward": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, C.c_size_t, _P, _P]),
    "svt_frame_postproc": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "svt_frame2note": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_double, C.c_double, C.c_double, _P, C.c_int,
                                 C.POINTER(C.c_int)]),
    "svt_op_gemm": (C.c_int, [_P, C.c_longlong, C.c_int, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int,
                              C.c_int, _P]),
    "svt_wavlm_relative_bucket": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "svt_video_transform_u8": (C.c_int, [_P, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _P, _P]),
    "svt_op_gemm_rowln": (C.c_int, [_P, C.c_longlong, C.c_int, _P, _P, _P, _P, C.c_float, C.c_int, _P, C.c_int, C.c_int, C.c_int, _P]),
    "svt_op_gemm_ln": (C.c_int, [_P, _P, _P, _P, _P, C.c_float, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "svt_op_row_stats_cast": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P]),
    "svt_op_posconv": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "svt_op_pack_posconv": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "svt_op_attention": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "svt_op_layer_norm": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_float, C.c_int, _P]),
    "svt_op_linear_small": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, C.c_int, _P, _P]),
    "svt_op_conv0": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, _P, C.c_int, _P, C.c_int, _P, _P]),
}
EXPORTED_SYMBOLS = tuple(_SIGS)


def lib():
    """Load the shared library (once).  Raises if it has not been built -- never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m svt_speechbrain_b200.build` "
                "(there is no CPU / PyTorch fallback for this package)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        # SVT_B200_OPTIONS="name=value,..." applies svt_set_option switches at load time (A/B runs of one build)
        for item in filter(None, os.environ.get("SVT_B200_OPTIONS", "").split(",")):
            name, _, value = item.partition("=")
            if L.svt_set_option(name.strip().encode(), int(value)) != 0:
                raise ValueError(f"SVT_B200_OPTIONS: {L.svt_last_error().decode()}")
        _lib = L
    return _lib


def check(status: int) -> None:
    if status != 0:
        raise SvtError(status, lib().svt_last_error().decode("utf-8", "replace"))


def ptr(t):
    """Raw address of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def current_stream_ptr():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(t, what: str):
    if not t.is_cuda:
        raise RuntimeError(
            f"{what}: expected a CUDA tensor -- svt_speechbrain_b200 is a B200 (sm_100a) implementation "
            "with no CPU fallback")
