"""ctypes binding of libhd_b200.so — the C-ABI boundary declared in include/hd_b200.h.

There is no CPU fallback: if the shared library is missing or a kernel call fails, this raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_int, c_void_p, c_float, c_size_t, c_char_p, c_longlong

import torch

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libhd_b200.so")

_lib = None

P = c_void_p
I = c_int
F = c_float
LL = c_longlong

# name -> (restype, argtypes)
_SIGNATURES = {
    "hd_last_error": (c_char_p, []),
    "hd_version": (I, []),
    "hd_launch_count": (LL, []),
    "hd_set_pdl": (None, [I]),
    "hd_trace_dump": (None, []),
    "hd_conv2d_igemm": (I, [P, P, P, P, P, P, P, P, I, I, I, I, I, I, I, I, I, I, I, I, P]),
    "hd_conv2d_igemm_affine": (I, [P, P, P, P, P, P, P, I, I, I, I, I, I, I, I, I, P]),
    "hd_conv2d_igemm_dual": (I, [P, P, P, P, P, P, I, I, I, I, I, I, I, I, I, P]),
    "hd_bn_fold_all": (I, [P, I, P, P]),
    "hd_set_conv_variant": (None, [I]),
    "hd_set_conv_debug": (None, [I]),
    "hd_conv2d_wgrad": (I, [P, P, P, P, I, I, I, I, I, I, I, I, I, P]),
    "hd_conv2d_wgrad_sync": (I, [P, P, P, P, I, I, I, I, I, I, I, I, I, P, P]),
    "hd_conv2d_wgrad_ksplit": (I, [I, I, I, I]),
    "hd_conv2d_wgrad_workspace_bytes": (c_size_t, [I, I, I, I, I]),
    "hd_pack_all_weights": (I, [P, I, LL, P]),
    "hd_pack_conv_weight": (I, [P, P, I, I, I, I, I, I, P]),
    "hd_nchw_f32_to_nhwc_bf16": (I, [P, P, I, I, I, I, I, P]),
    "hd_nhwc_bf16_to_nchw_f32": (I, [P, P, I, I, I, I, I, P]),
    "hd_stem_im2col": (I, [P, P, I, I, I, P]),
    "hd_stem_pack_weight": (I, [P, P, I, P]),
    "hd_stem_unfold": (I, [P, P, I, I, I, P]),
    "hd_conv2d_igemm_vtaps": (I, [P, P, P, P, P, P, I, I, I, I, I, I, I, I, I, P, P, P, I, P]),
    "hd_head_backward": (I, [P, LL, P, I, P, P, P, P, P, I, I, I, I, P]),
    "hd_bn_finalize": (I, [P, P, F, P, P, P, P, P, F, F, I, P, P, P, P, I, P]),
    "hd_bn_act": (I, [P, P, P, P, LL, I, I, P]),
    "hd_bn_add_relu": (I, [P, P, P, P, P, P, P, LL, I, P]),
    "hd_bn_add_relu_mask": (I, [P, P, P, P, P, P, P, P, LL, I, P]),
    "hd_bn_bwd_reduce_fin_mask": (I, [P, P, P, P, LL, I, P, P]),
    "hd_bn_bwd_apply_mask": (I, [P, P, P, P, P, P, LL, I, P]),
    "hd_maxpool2": (I, [P, P, I, I, I, I, P]),
    "hd_upsample2_add": (I, [P, P, P, I, I, I, I, P]),
    "hd_bn_bwd_reduce": (I, [P, P, P, P, P, P, P, P, P, P, P, LL, I, P]),
    "hd_bn_bwd_reduce_fin": (I, [P, P, P, P, P, P, P, P, P, LL, I, P, P]),
    "hd_bn_bwd_finalize": (I, [P, P, F, P, P, P, P, P, P, I, I, P]),
    "hd_bn_bwd_fused_small": (I, [P, P, P, P, P, P, P, P, LL, I, P, P, P]),
    "hd_bn_bwd_apply": (I, [P, P, P, P, P, P, P, P, P, P, P, P, P, LL, I, P]),
    "hd_maxpool2_bwd": (I, [P, P, P, P, P, I, I, I, I, P]),
    "hd_bn_add_relu_pool2": (I, [P, P, P, P, P, P, P, P, I, I, I, I, P]),
    "hd_maxpool2_bwd_idx": (I, [P, P, P, P, P, I, I, I, I, P]),
    "hd_bn_bwd_reduce_pool_fin": (I, [P, P, P, P, P, P, P, P, P, I, I, I, I, P, P]),
    "hd_bn_bwd_apply_pool": (I, [P, P, P, P, P, P, P, P, P, P, P, P, I, I, I, I, P]),
    "hd_conv2d_igemm_halo_eligible": (I, [I, I, I, I, I]),
    "hd_conv2d_igemm_bwdstat": (I, [P, P, P, I, I, I, I, I, I, P, P, P, P, P, P]),
    "hd_sum2x2": (I, [P, P, I, I, I, I, P]),
    "hd_add": (I, [P, P, P, P, LL, P]),
    "hd_colsum": (I, [P, P, LL, I, I, P]),
    "hd_loss_forward": (I, [P, LL, P, LL, P, LL, P, P, P, P, I, I, I, I, F, F, F, F, F, I, I, P, P, P]),
    "hd_loss_backward": (I, [P, LL, P, LL, P, LL, P, P, P, P, I, I, I, I, F, F, F, F, F, I, I, P, P,
                             P, LL, P, LL, P, LL, P]),
    "hd_encode_targets": (I, [P, P, I, I, I, I, I, I, I, P, P, P, P, P, P]),
    "hd_normalize_u8": (I, [P, P, I, I, I, P, P, P]),
    "hd_decode_scratch_bytes": (c_size_t, [I, I, I, I, I]),
    "hd_decode_scratch_init": (I, [P, I, I, P]),
    "hd_decode_nms": (I, [P, LL, LL, P, LL, LL, P, LL, LL, I, I, I, I, I, I, F, F, F, I, I, I, P, P, P, P, P, P]),
}


def exported_symbols() -> list[str]:
    return sorted(_SIGNATURES)


def register(name: str, restype, argtypes) -> None:
    _SIGNATURES[name] = (restype, argtypes)
    if _lib is not None:
        fn = getattr(_lib, name)
        fn.restype, fn.argtypes = restype, argtypes


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the sm_100a CUDA extension has not been built. "
                "Run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
                "There is no CPU fallback for this path.")
        _lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(_lib, name)  # AttributeError if the .so is stale
            fn.restype, fn.argtypes = restype, argtypes
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().hd_last_error()
        raise RuntimeError(f"libhd_b200 {what} failed (code {rc}): {msg.decode() if msg else ''}")


def ptr(t) -> c_void_p | None:
    if t is None:
        return None
    return c_void_p(t.data_ptr())


def stream(device=None) -> c_void_p:
    """The current torch stream of `device` (default: the current device)."""
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t: torch.Tensor, name: str = "tensor") -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must live on a CUDA device: this path has no CPU implementation")
