"""ctypes binding of the C ABI declared in include/s3prl_b200.h.

PyTorch is used only as the owner of device memory and streams: every call below passes raw
``data_ptr()`` addresses and the current ``cudaStream_t`` to the shared library. There is no CPU
fallback: on a host without a CUDA device, anything that computes raises ``S3BError``.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Optional

import os

# S3B_LIB_PATH: A/B runs of another build of the same C ABI on one box (tools/ab_step.sh); never a CPU fallback
_LIB_PATH = Path(os.environ.get("S3B_LIB_PATH") or Path(__file__).resolve().parent / "_lib" / "libs3prl_b200.so")

EXPORTED_SYMBOLS = [
    "s3b_version",
    "s3b_last_error",
    "s3b_device_count",
    "s3b_model_create",
    "s3b_model_set_tensor",
    "s3b_model_finalize",
    "s3b_model_destroy",
    "s3b_num_frames",
    "s3b_valid_frames",
    "s3b_num_outputs",
    "s3b_default_lanes",
    "s3b_forward",
    "s3b_forward_host",
    "s3b_forward_ex",
    "s3b_forward_host_ex",
    "s3b_wavlm_buckets",
    "s3b_peer_create",
    "s3b_peer_connect",
    "s3b_peer_slot",
    "s3b_peer_push",
    "s3b_peer_wait",
    "s3b_peer_destroy",
    "s3b_profile_enable",
    "s3b_profile_read",
    "s3b_launch_count",
    "s3b_weighted_sum",
    "s3b_weighted_sum_backward",
    "s3b_linear_f32",
    "s3b_layernorm_f32",
    "s3b_gemm_bench",
    "s3b_attention_f32",
    "s3b_fbank",
    "s3b_fbank_num_frames",
    "s3b_trimmed_lengths",
    "s3b_melspec",
]


class S3BError(RuntimeError):
    pass


class S3BConfig(C.Structure):
    """Mirror of ``struct s3b_config``."""

    _fields_ = [
        ("family", C.c_int32),
        ("extractor_layer_norm", C.c_int32),
        ("conv_bias", C.c_int32),
        ("layer_norm_first", C.c_int32),
        ("normalize_wav", C.c_int32),
        ("num_layers", C.c_int32),
        ("embed_dim", C.c_int32),
        ("ffn_dim", C.c_int32),
        ("num_heads", C.c_int32),
        ("pos_conv_kernel", C.c_int32),
        ("pos_conv_groups", C.c_int32),
        ("relative_position", C.c_int32),
        ("num_buckets", C.c_int32),
        ("max_distance", C.c_int32),
        ("gru_rel_pos", C.c_int32),
        ("no_feature_layer_norm", C.c_int32),
        ("pred_heads", C.c_int32),
        ("pos_conv_depth", C.c_int32),
        ("reserved", C.c_int32 * 5),
    ]


class S3BForwardOpts(C.Structure):
    """Mirror of ``struct s3b_forward_opts``."""

    _fields_ = [
        ("struct_size", C.c_int32),
        ("lanes", C.c_int32),
        ("layer_stride", C.c_int64),
        ("ffn_out", C.c_void_p),
        ("ffn_layer_stride", C.c_int64),
        ("last_residual", C.c_void_p),
        ("reserved", C.c_int64 * 4),
    ]

    def __init__(self, **kw):
        super().__init__(**kw)
        self.struct_size = C.sizeof(S3BForwardOpts)


FAMILY_HUBERT, FAMILY_WAV2VEC2, FAMILY_WAVLM, FAMILY_DISTILLER = 0, 1, 2, 3

_lib: Optional[C.CDLL] = None


def lib_path() -> Path:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load the shared library (building is `__graft_entry__.build()` / `python -m s3prl_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise S3BError(
            f"{_LIB_PATH} is missing: build it with `python s3prl_b200/build.py` (needs nvcc). "
            "s3prl_b200 has no CPU fallback."
        )
    lib = C.CDLL(str(_LIB_PATH))
    vp, i32, i64, f32p = C.c_void_p, C.c_int32, C.c_int64, C.c_void_p
    lib.s3b_version.restype = C.c_int
    lib.s3b_last_error.restype = C.c_char_p
    lib.s3b_device_count.restype = C.c_int
    lib.s3b_model_create.argtypes = [C.POINTER(S3BConfig), C.POINTER(vp)]
    lib.s3b_model_set_tensor.argtypes = [vp, C.c_char_p, vp, C.POINTER(i64), i32]
    lib.s3b_model_finalize.argtypes = [vp]
    lib.s3b_model_destroy.argtypes = [vp]
    lib.s3b_model_destroy.restype = None
    lib.s3b_num_frames.argtypes = [vp, i64]
    lib.s3b_num_frames.restype = i64
    lib.s3b_num_outputs.argtypes = [vp]
    lib.s3b_num_outputs.restype = i32
    lib.s3b_default_lanes.argtypes = [vp, i32, i64]
    lib.s3b_default_lanes.restype = i32
    lib.s3b_valid_frames.argtypes = [vp, C.POINTER(i64), i32, i64, C.POINTER(i32)]
    lib.s3b_forward.argtypes = [vp, C.POINTER(vp), C.POINTER(i64), i32, i64, f32p, vp]
    lib.s3b_forward_host.argtypes = [vp, C.POINTER(vp), C.POINTER(i64), i32, i64, f32p]
    lib.s3b_forward_ex.argtypes = [vp, C.POINTER(vp), C.POINTER(i64), i32, i64, f32p, vp, C.POINTER(S3BForwardOpts)]
    lib.s3b_forward_host_ex.argtypes = [vp, C.POINTER(vp), C.POINTER(i64), i32, i64, f32p, f32p]
    lib.s3b_wavlm_buckets.argtypes = [i32, i32, C.POINTER(i32), i32, C.POINTER(i32)]
    lib.s3b_peer_create.argtypes = [i32, i32, i64, i32, C.POINTER(vp), vp]
    lib.s3b_peer_connect.argtypes = [vp, vp]
    lib.s3b_peer_slot.argtypes = [vp, C.c_uint32]
    lib.s3b_peer_slot.restype = vp
    lib.s3b_peer_push.argtypes = [vp, f32p, i32, i64, f32p, C.c_uint32, vp]
    lib.s3b_peer_wait.argtypes = [vp, C.c_uint32, vp]
    lib.s3b_peer_destroy.argtypes = [vp]
    lib.s3b_peer_destroy.restype = None
    lib.s3b_profile_enable.argtypes = [vp, i32]
    lib.s3b_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(i64), i32]
    lib.s3b_launch_count.argtypes = [vp]
    lib.s3b_launch_count.restype = i64
    lib.s3b_weighted_sum.argtypes = [f32p, i32, i64, f32p, f32p, vp]
    lib.s3b_weighted_sum_backward.argtypes = [f32p, i32, i64, f32p, f32p, vp]
    lib.s3b_linear_f32.argtypes = [f32p, f32p, f32p, f32p, i64, i32, i32, i32, f32p, vp]
    lib.s3b_gemm_bench.argtypes = [i64, i32, i32, i32, i32, i32, C.POINTER(C.c_float)]
    lib.s3b_layernorm_f32.argtypes = [f32p, i64, i32, f32p, f32p, i32, f32p, vp]
    lib.s3b_attention_f32.argtypes = [f32p, f32p, f32p, C.POINTER(i32), i32, i32, i32, f32p, vp]
    lib.s3b_fbank_num_frames.argtypes = [i64]
    lib.s3b_fbank_num_frames.restype = i64
    lib.s3b_fbank.argtypes = [C.POINTER(vp), C.POINTER(i64), i32, f32p, vp]
    lib.s3b_trimmed_lengths.argtypes = [C.POINTER(vp), C.POINTER(i64), i32, C.POINTER(i64)]
    lib.s3b_melspec.argtypes = [C.POINTER(vp), C.POINTER(i64), i32, i64, i32, C.POINTER(i32), C.POINTER(i32), i32, f32p, vp]
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().s3b_last_error()
        raise S3BError(msg.decode() if msg else f"s3prl_b200 call failed with status {status}")


def require_gpu() -> None:
    if load().s3b_device_count() < 1:
        raise S3BError("no CUDA device visible: s3prl_b200 has no CPU fallback")


def current_stream_ptr() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream
