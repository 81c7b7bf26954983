"""Thin per-kernel Python wrappers over the C ABI (include/hd_b200.h).

These are what the parity tests call; the nn.Module path (hourglass.py) drives the same kernels
through the native network executor. Every function enqueues on the current CUDA stream, allocates
outputs with torch, never synchronises and has no CPU fallback.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import ptr, stream, check

BF16 = torch.bfloat16


def _block_n(cout: int) -> int:
    if cout > 64:
        return 128
    if cout > 16:
        return 64
    return 16


def pack_weight(w: torch.Tensor, mode: int = 0, rows_pad: int | None = None, k_pad: int | None = None) -> torch.Tensor:
    """OIHW fp32 -> [taps, rows_pad, k_pad] bf16. mode 0: forward operand, mode 1: dgrad operand."""
    _lib.require_cuda(w, "weight")
    w = w.detach().contiguous().float()
    cout, cin, kh, kw = w.shape
    assert kh == kw
    rows, kdim = (cout, cin) if mode == 0 else (cin, cout)
    rows_pad = rows_pad or _block_n(rows)
    k_pad = k_pad or ((kdim + 63) // 64) * 64
    out = torch.empty((kh * kw, rows_pad, k_pad), dtype=BF16, device=w.device)
    check(_lib.lib().hd_pack_conv_weight(ptr(w), ptr(out), cout, cin, kh, rows_pad, k_pad, mode, stream()),
          "pack_conv_weight")
    return out


def to_nhwc(x: torch.Tensor, c_pad: int | None = None) -> torch.Tensor:
    _lib.require_cuda(x, "x")
    x = x.contiguous().float()
    n, c, h, w = x.shape
    c_pad = c_pad or c
    y = torch.empty((n, h, w, c_pad), dtype=BF16, device=x.device)
    check(_lib.lib().hd_nchw_f32_to_nhwc_bf16(ptr(x), ptr(y), n, c, h, w, c_pad, stream()), "nchw_to_nhwc")
    return y


def to_nchw(x: torch.Tensor, c: int | None = None) -> torch.Tensor:
    _lib.require_cuda(x, "x")
    n, h, w, cs = x.shape
    c = c or cs
    y = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    check(_lib.lib().hd_nhwc_bf16_to_nchw_f32(ptr(x), ptr(y), n, c, h, w, cs, stream()), "nhwc_to_nchw")
    return y


def conv2d_igemm(x: torch.Tensor, w_packed: torch.Tensor, cout: int, ksize: int, bias: torch.Tensor | None = None,
                 addend: torch.Tensor | None = None, stats: torch.Tensor | None = None,
                 out: torch.Tensor | None = None, head_out: torch.Tensor | None = None, stack_idx: int = 0,
                 out2: torch.Tensor | None = None) -> torch.Tensor:
    """x: NHWC bf16 (C % 64 == 0). w_packed from pack_weight. stats: fp32 [2, cout] accumulated in place.

    head_out: fp32 (B, S, cout, H, W) logits tensor; when given the kernel writes slice [:, stack_idx] (NCHW fp32).
    """
    _lib.require_cuda(x, "x")
    n, h, w, cin = x.shape
    block_n = w_packed.shape[1]
    assert w_packed.shape[2] == cin and w_packed.shape[0] == ksize * ksize
    if head_out is not None:
        out_t, mode, num_stack, out_cs = head_out, 1, head_out.shape[1], 0
    else:
        if out is None:
            out = torch.empty((n, h, w, cout), dtype=BF16, device=x.device)
        out_t, mode, num_stack, out_cs = out, 0, 1, out.shape[3]
    check(_lib.lib().hd_conv2d_igemm(
        ptr(x), ptr(w_packed), ptr(out_t), ptr(out2), ptr(bias), ptr(addend),
        ptr(stats[0]) if stats is not None else None, ptr(stats[1]) if stats is not None else None,
        n, h, w, cin, cout, block_n, ksize, mode, out_cs, out2.shape[3] if out2 is not None else 0,
        stack_idx, num_stack, stream()), "conv2d_igemm")
    return out_t


def conv2d_wgrad(x: torch.Tensor, dy: torch.Tensor, cin_real: int, ksize: int, grad: torch.Tensor | None = None,
                 accumulate: bool = False) -> torch.Tensor:
    """x: NHWC bf16 [N,H,W,cin], dy: NHWC bf16 [N,H,W,128] -> grad OIHW fp32 [128, cin_real, k, k]."""
    _lib.require_cuda(x, "x")
    n, h, w, cin = x.shape
    cout = dy.shape[3]
    if grad is None:
        grad = torch.empty((cout, cin_real, ksize, ksize), dtype=torch.float32, device=x.device)
        accumulate = False
    nbytes = _lib.lib().hd_conv2d_wgrad_workspace_bytes(n, h, w, cin, ksize)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=x.device)
    check(_lib.lib().hd_conv2d_wgrad(ptr(x), ptr(dy), ptr(grad), ptr(ws), n, h, w, cin, cin_real, cout, ksize,
                                     1 if accumulate else 0, stream()), "conv2d_wgrad")
    return grad
