"""ctypes binding of libvlmclip_b200.so (include/vlmclip.h).

There is deliberately no fallback: if the library is missing or a call fails, the caller gets an exception.
PyTorch is used only as the owner of device memory and streams; every entry point receives raw device
pointers (`tensor.data_ptr()`) and the current CUDA stream handle.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libvlmclip_b200.so"
_lib = None

ACT_NONE, ACT_QUICK_GELU, ACT_GELU_ERF, ACT_RELU = 0, 1, 2, 3
POST_RESIDUAL_LN, POST_RESIDUAL, POST_BLEND_L2, POST_PLAIN = 0, 1, 2, 3

_p, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> (restype, argtypes); mirrors include/vlmclip.h one to one (tests/test_abi.py checks the header)
PROTOTYPES = {
    "vlmclip_abi_version": (_i, []),
    "vlmclip_last_error": (C.c_char_p, []),
    "vlmclip_launch_count": (_i64, []),
    "vlmclip_gemm_bf16": (_i, [_p, _i64, _p, _i64, _p, _i64, _p, _p, _i64, _p, _p, _p, _i, _f, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "vlmclip_gemm_bf16_res2": (_i, [_p, _i64, _p, _i64, _p, _i64, _i64, _p, _p, _p, _p, _f, _i, _i, _i, _p]),
    "vlmclip_gemm_bf16_splitk": (_i, [_p, _i64, _p, _i64, _p, _i64, _i64, _i, _i, _i, _i, _p]),
    "vlmclip_sum_planes_f32": (_i, [_p, _i64, _i, _p, _i64, _p]),
    "vlmclip_gemm_bf16_atb_splitk": (_i, [_p, _i64, _p, _i64, _p, _i64, _i64, _i, _i, _i, _i, _p]),
    "vlmclip_colsum_bf16_slices": (_i, [_i]),
    "vlmclip_colsum_bf16": (_i, [_p, _i64, _p, _p, _i, _i, _p]),
    "vlmclip_layernorm_bf16": (_i, [_p, _i64, _p, _i64, _p, _p, _p, _i, _i, _f, _p]),
    "vlmclip_layernorm_bf16_f32out": (_i, [_p, _i64, _p, _i64, _p, _p, _i, _i, _f, _p]),
    "vlmclip_ln_partials_to_stats": (_i, [_p, _p, _i, _i, _f, _p]),
    "vlmclip_row_stats_bf16": (_i, [_p, _i64, _p, _i, _i, _f, _p]),
    "vlmclip_im2col_patches": (_i, [_p, _i, _p, _i, _i, _i, _i, _p]),
    "vlmclip_preprocess_patches": (_i, [_p, _i64, _i, _i, _i, _p, _p, _f, _f, _f, _f, _f, _f, _p, _i, _i, _i, _i, _p]),
    "vlmclip_mean_pool": (_i, [_p, _p, _i, _i, _i, _p]),
    "vlmclip_mean_pool_bwd": (_i, [_p, _p, _i, _i, _i, _p]),
    "vlmclip_vision_embed_ln": (_i, [_p, _i, _p, _p, _p, _p, _p, _p, _i, _i, _i, _f, _p]),
    "vlmclip_text_embed": (_i, [_p, _p, _i, _p, _p, _p, _i, _i, _i, _i, _p]),
    "vlmclip_attention_fwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _f, _p]),
    "vlmclip_attention_fwd_workspace": (_i64, [_i, _i, _i]),
    "vlmclip_attention_fwd_ws": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _f, _p]),
    "vlmclip_attention_1q": (_i, [_p, _i64, _p, _p, _i64, _i64, _p, _i, _i, _i, _f, _p]),
    "vlmclip_encoder_fwd": (_i, [_p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _i, _i, _p]),
    "vlmclip_adapter_fwd": (_i, [_p, _i, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _f, _p]),
    "vlmclip_adapter_bwd_workspace": (_i64, [_i, _i, _i]),
    "vlmclip_adapter_bwd": (_i, [_p, _i, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p,
                                 _i, _i, _i, _i, _i, _f, _f, _p]),
    "vlmclip_linear_f32": (_i, [_p, _i64, _p, _p, _p, _i, _i, _i, _p]),
    "vlmclip_linear_f32_dgrad": (_i, [_p, _p, _p, _i, _i, _i, _p]),
    "vlmclip_clip_loss_workspace": (_i64, [_i, _i]),
    "vlmclip_clip_loss": (_i, [_p, _p, _f, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "vlmclip_clip_loss_state_size": (_i64, [_i, _i, _i]),
    "vlmclip_clip_loss_counters": (_i64, [_i]),
    "vlmclip_scale_f32": (_i, [_p, _p, _p, _i64, _p]),
    "vlmclip_clip_loss_bwd_workspace": (_i64, [_i, _i, _i]),
    "vlmclip_clip_loss_fwd": (_i, [_p, _p, _f, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "vlmclip_clip_loss_bwd": (_i, [_p, _p, _p, _i, _i, _f, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "vlmclip_class_head": (_i, [_p, _p, _f, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "vlmclip_class_head_bwd": (_i, [_p, _p, _p, _f, _p, _p, _i, _i, _i, _p]),
    "vlmclip_l2norm_rows": (_i, [_p, _p, _p, _p, _i, _i, _p]),
    "vlmclip_l2norm_rows_bwd": (_i, [_p, _p, _p, _i, _i, _p]),
    "vlmclip_adamw_clip_step": (_i, [_p, _p, _p, _p, _i64, _p, _f, _f, _f, _f, _f, _p, _p, _p, _p]),
    "vlmclip_gather_rows_bf16_to_f32": (_i, [_p, _i64, _p, _i, _i, _p]),
    "vlmclip_gather_rows2_bf16_to_f32": (_i, [_p, _p, _i64, _p, _i, _i, _p]),
    "vlmclip_layernorm_f32": (_i, [_p, _i64, _p, _p, _p, _p, _i, _i, _f, _p]),
    "vlmclip_layernorm_f32_bwd": (_i, [_p, _p, _i64, _p, _p, _p, _p, _p, _p, _i, _i, _p]),
    "vlmclip_gelu_f32": (_i, [_p, _p, _i64, _p]),
    "vlmclip_gelu_f32_bwd": (_i, [_p, _p, _p, _i64, _p]),
    "vlmclip_fma_mask_f32": (_i, [_p, _p, _p, _p, _i64, _p]),
    "vlmclip_attn1q_f32_fwd": (_i, [_p, _p, _p, _i64, _p, _p, _p, _i, _i, _i, _f, _p]),
    "vlmclip_attn1q_f32_bwd": (_i, [_p, _p, _p, _p, _i64, _p, _p, _p, _p, _p, _p, _i, _i, _i, _f, _p]),
    "vlmclip_transpose_to_bf16": (_i, [_p, _i, _i64, _p, _i64, _i, _i, _i, _i, _i, _i, _p]),
    "vlmclip_cast_f32_to_bf16": (_i, [_p, _p, _i64, _p]),
    "vlmclip_add_bf16_into_f32": (_i, [_p, _p, _i64, _p]),
    "vlmclip_rowsum_bf16": (_i, [_p, _i64, _p, _i, _i, _p]),
    "vlmclip_colsum_f32": (_i, [_p, _i64, _p, _i, _i64, _p]),
    "vlmclip_quick_gelu_bf16": (_i, [_p, _p, _i64, _p]),
    "vlmclip_quick_gelu_bwd_bf16": (_i, [_p, _p, _p, _i64, _p]),
    "vlmclip_layernorm_bwd_workspace": (_i64, [_i, _i]),
    "vlmclip_layernorm_bwd": (_i, [_p, _i, _i64, _p, _i64, _p, _p, _p, _p, _i64, _p, _p, _p, _i, _i, _f, _p]),
    "vlmclip_vision_embed": (_i, [_p, _p, _p, _p, _i, _i, _i, _p]),
    "vlmclip_embed_scatter_add": (_i, [_p, _i64, _p, _p, _i64, _i, _i, _p]),
    "vlmclip_attention_bwd_workspace": (_i64, [_i, _i, _i]),
    "vlmclip_attention_bwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _p]),
    "vlmclip_linear_f32_wgrad": (_i, [_p, _p, _i64, _p, _i, _i, _i, _p]),
}


class LayerPtrs(C.Structure):
    """vlmclip_layer_t of include/vlmclip.h: device pointers of one frozen encoder layer (LayerNorm folded)."""
    _fields_ = [(n, C.c_void_p) for n in ("qkv_w", "qkv_b", "qkv_c", "out_w", "out_b", "fc1_w", "fc1_b", "fc1_c",
                                           "fc2_w", "fc2_b")]


class NativeError(RuntimeError):
    pass


def lib_path() -> Path:
    return _LIB_PATH


def load():
    """dlopen the library (building it is `__graft_entry__.build()`'s job, not this function's)."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise NativeError(
            f"{_LIB_PATH} is missing: build it with `python -m vlm_clip_b200.build` "
            "(the CUDA library is the only compute path; there is no fallback)"
        )
    lib = C.CDLL(os.fspath(_LIB_PATH))
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    msg = load().vlmclip_last_error().decode("utf-8", "replace")
    if rc < 0:
        raise ValueError(f"{what}: {msg}")
    raise NativeError(f"{what}: CUDA error {rc}: {msg}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  The tensor must be on a CUDA device."""
    if t is None:
        return None
    if not t.is_cuda:
        raise NativeError("vlm_clip_b200 kernels need CUDA tensors (there is no CPU path)")
    return C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count() -> int:
    return int(load().vlmclip_launch_count())
