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


def conv2d_igemm_affine(x: torch.Tensor, w_packed: torch.Tensor, cout: int, ksize: int, scale: torch.Tensor | None,
                        shift: torch.Tensor | None, relu: bool, bias: torch.Tensor | None = None,
                        addend: torch.Tensor | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
    """Eval-mode `Convolution` / `Residual` tail in one launch: relu?((conv + bias) * scale + shift + addend), NHWC bf16."""
    _lib.require_cuda(x, "x")
    n, h, w, cin = x.shape
    block_n = w_packed.shape[1]
    assert w_packed.shape[2] == cin and w_packed.shape[0] == ksize * ksize
    if out is None:
        out = torch.empty((n, h, w, cout), dtype=BF16, device=x.device)
    check(_lib.lib().hd_conv2d_igemm_affine(ptr(x), ptr(w_packed), ptr(out), ptr(bias), ptr(addend), ptr(scale), ptr(shift),
                                            1 if relu else 0, n, h, w, cin, cout, block_n, ksize, out.shape[3], stream()),
          "conv2d_igemm_affine")
    return out


def conv2d_igemm_dual(x: torch.Tensor, w_packed: torch.Tensor, x2: torch.Tensor, w2_packed: torch.Tensor, cout: int,
                      ksize: int, addend: torch.Tensor | None = None) -> torch.Tensor:
    """conv_{k x k}(x; w) + conv_{1x1}(x2; w2) (+ addend) in one pass (64 output channels, large maps): the fused
    dgrad of a `Residual` with a skip convolution."""
    _lib.require_cuda(x, "x")
    n, h, w, cin = x.shape
    out = torch.empty((n, h, w, cout), dtype=BF16, device=x.device)
    check(_lib.lib().hd_conv2d_igemm_dual(ptr(x), ptr(w_packed), ptr(x2), ptr(w2_packed), ptr(out), ptr(addend), n, h, w,
                                          cin, x2.shape[3], cout, w_packed.shape[1], ksize, cout, stream()),
          "conv2d_igemm_dual")
    return out


def conv2d_wgrad(x: torch.Tensor, dy: torch.Tensor, cin_real: int, ksize: int, grad: torch.Tensor | None = None,
                 accumulate: bool = False, stem_perm: bool = False) -> torch.Tensor:
    """x: NHWC bf16 [N,H,W,cin], dy: NHWC bf16 [N,H,W,128] -> grad OIHW fp32 [128, cin_real, k, k]."""
    _lib.require_cuda(x, "x")
    n, h, w, cin = x.shape
    cout = dy.shape[3]
    if grad is None:
        shape = (cout, 3, 7, 7) if stem_perm else (cout, cin_real, ksize, ksize)
        grad = torch.empty(shape, dtype=torch.float32, device=x.device)
        accumulate = False
    nbytes = _lib.lib().hd_conv2d_wgrad_workspace_bytes(n, h, w, cin, ksize)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=x.device)
    check(_lib.lib().hd_conv2d_wgrad(ptr(x), ptr(dy), ptr(grad), ptr(ws), n, h, w, cin, cin_real, cout, ksize,
                                     1 if accumulate else 0, 1 if stem_perm else 0, stream()), "conv2d_wgrad")
    return grad


# ------------------------------------------------------------------------------------------------ stem
def stem_im2col(x: torch.Tensor) -> torch.Tensor:
    """(B,3,H,W) fp32 NCHW image -> (B,H/2,W/2,192) bf16 patch matrix of the 7x7 stride-2 stem."""
    _lib.require_cuda(x, "x")
    x = x.contiguous().float()
    n, c, h, w = x.shape
    assert c == 3
    out = torch.empty((n, h // 2, w // 2, 192), dtype=BF16, device=x.device)
    check(_lib.lib().hd_stem_im2col(ptr(x), ptr(out), n, h, w, stream()), "stem_im2col")
    return out


def stem_pack_weight(w: torch.Tensor) -> torch.Tensor:
    w = w.detach().contiguous().float()
    out = torch.empty((1, 64, 192), dtype=BF16, device=w.device)
    check(_lib.lib().hd_stem_pack_weight(ptr(w), ptr(out), w.shape[0], stream()), "stem_pack_weight")
    return out


# ------------------------------------------------------------------------------------------------ BatchNorm & friends
def bn_finalize(stats, count, gamma, beta, running_mean=None, running_var=None, nbt=None, momentum=0.1, eps=1e-5,
                training=True):
    """stats: fp32 [2, C] (sum, sum of squares). Returns bnp fp32 [4, C]: scale, shift, mean, rstd."""
    C = gamma.numel()
    bnp = torch.empty((4, C), dtype=torch.float32, device=gamma.device)
    check(_lib.lib().hd_bn_finalize(ptr(stats[0]) if stats is not None else None,
                                    ptr(stats[1]) if stats is not None else None, float(count), ptr(gamma),
                                    ptr(beta), ptr(running_mean), ptr(running_var), ptr(nbt), momentum, eps,
                                    1 if training else 0, ptr(bnp[0]), ptr(bnp[1]), ptr(bnp[2]), ptr(bnp[3]), C,
                                    stream()), "bn_finalize")
    return bnp


def bn_act(y, bnp, relu=True):
    z = torch.empty_like(y)
    C = y.shape[-1]
    check(_lib.lib().hd_bn_act(ptr(y), ptr(bnp[0]), ptr(bnp[1]), ptr(z), y.numel() // C, C, 1 if relu else 0,
                               stream()), "bn_act")
    return z


def bn_add_relu(y2, bnp2, skip, bnp_s=None):
    out = torch.empty_like(y2)
    C = y2.shape[-1]
    check(_lib.lib().hd_bn_add_relu(ptr(y2), ptr(bnp2[0]), ptr(bnp2[1]), ptr(skip),
                                    ptr(bnp_s[0]) if bnp_s is not None else None,
                                    ptr(bnp_s[1]) if bnp_s is not None else None, ptr(out), y2.numel() // C, C,
                                    stream()), "bn_add_relu")
    return out


def bn_add_relu_pool2(y2, bnp2, ys, bnp_s):
    """maxpool2(relu(bn(y2) + bn_s(ys))) without materialising the un-pooled tensor: (pooled bf16, argmax uint8)."""
    n, h, w, c = y2.shape
    pooled = torch.empty((n, h // 2, w // 2, c), dtype=BF16, device=y2.device)
    idx = torch.empty((n, h // 2, w // 2, c), dtype=torch.uint8, device=y2.device)
    check(_lib.lib().hd_bn_add_relu_pool2(ptr(y2), ptr(bnp2[0]), ptr(bnp2[1]), ptr(ys), ptr(bnp_s[0]), ptr(bnp_s[1]),
                                          ptr(pooled), ptr(idx), n, h, w, c, stream()), "bn_add_relu_pool2")
    return pooled, idx


def maxpool2_bwd_idx(idx, dpool, add1=None, add2=None):
    n, ho, wo, c = dpool.shape
    dx = torch.empty((n, 2 * ho, 2 * wo, c), dtype=BF16, device=dpool.device)
    check(_lib.lib().hd_maxpool2_bwd_idx(ptr(idx), ptr(dpool), ptr(add1), ptr(add2), ptr(dx), n, 2 * ho, 2 * wo, c,
                                         stream()), "maxpool2_bwd_idx")
    return dx


def maxpool2(x):
    n, h, w, c = x.shape
    y = torch.empty((n, h // 2, w // 2, c), dtype=BF16, device=x.device)
    check(_lib.lib().hd_maxpool2(ptr(x), ptr(y), n, h, w, c, stream()), "maxpool2")
    return y


def upsample2_add(up1, low):
    n, h, w, c = up1.shape
    out = torch.empty_like(up1)
    check(_lib.lib().hd_upsample2_add(ptr(up1), ptr(low), ptr(out), n, h, w, c, stream()), "upsample2_add")
    return out


def bn_bwd(dout, out, y, bnp, gamma, ys=None, bnp_s=None, gamma_s=None, want_g=False, remask=False):
    """Backward of relu(bn(y) [+ bn_s(ys) | + x]) w.r.t. y (and ys): returns dy, dys, g, (dgamma, dbeta), (dgamma_s, dbeta_s)."""
    C = y.shape[-1]
    npix = y.numel() // C
    sums = torch.zeros((3, C), dtype=torch.float32, device=y.device)
    L = _lib.lib()
    # remask: rebuild the ReLU mask from y*scale+shift (+ ys*scale_s+shift_s for a two-branch tail) instead of reading `out`
    out_arg = None if remask else out
    check(L.hd_bn_bwd_reduce_fin(ptr(dout), ptr(out_arg), ptr(bnp[0]), ptr(bnp[1]),
                                 ptr(bnp_s[0]) if ys is not None else None, ptr(bnp_s[1]) if ys is not None else None,
                                 ptr(y), ptr(ys), ptr(sums), npix, C, None, stream()), "bn_bwd_reduce")
    coef = torch.empty((3, C), dtype=torch.float32, device=y.device)
    dgamma = torch.empty(C, dtype=torch.float32, device=y.device)
    dbeta = torch.empty(C, dtype=torch.float32, device=y.device)
    check(L.hd_bn_bwd_finalize(ptr(sums[0]), ptr(sums[1]), float(npix), ptr(gamma), ptr(bnp[2]), ptr(bnp[3]),
                               ptr(coef), ptr(dgamma), ptr(dbeta), 0, C, stream()), "bn_bwd_finalize")
    coef_s = dgs = dbs = dys = None
    if ys is not None:
        coef_s = torch.empty((3, C), dtype=torch.float32, device=y.device)
        dgs = torch.empty(C, dtype=torch.float32, device=y.device)
        dbs = torch.empty(C, dtype=torch.float32, device=y.device)
        check(L.hd_bn_bwd_finalize(ptr(sums[0]), ptr(sums[2]), float(npix), ptr(gamma_s), ptr(bnp_s[2]),
                                   ptr(bnp_s[3]), ptr(coef_s), ptr(dgs), ptr(dbs), 0, C, stream()),
              "bn_bwd_finalize")
        dys = torch.empty_like(ys)
    dy = torch.empty_like(y)
    g = torch.empty_like(y) if want_g else None
    check(L.hd_bn_bwd_apply(ptr(dout), ptr(out_arg), ptr(bnp[0]), ptr(bnp[1]),
                            ptr(bnp_s[0]) if ys is not None else None, ptr(bnp_s[1]) if ys is not None else None,
                            ptr(y), ptr(coef), ptr(dy), ptr(ys),
                            ptr(coef_s), ptr(dys), ptr(g),
                            npix, C, stream()), "bn_bwd_apply")
    return dy, dys, g, (dgamma, dbeta), (dgs, dbs)


def maxpool2_bwd(x, dpool, add1=None, add2=None):
    n, h, w, c = x.shape
    dx = torch.empty_like(x)
    check(_lib.lib().hd_maxpool2_bwd(ptr(x), ptr(dpool), ptr(add1), ptr(add2), ptr(dx), n, h, w, c, stream()),
          "maxpool2_bwd")
    return dx


def sum2x2(dout):
    n, h, w, c = dout.shape
    dlow = torch.empty((n, h // 2, w // 2, c), dtype=BF16, device=dout.device)
    check(_lib.lib().hd_sum2x2(ptr(dout), ptr(dlow), n, h, w, c, stream()), "sum2x2")
    return dlow


def add(a, b, c=None):
    out = torch.empty_like(a)
    check(_lib.lib().hd_add(ptr(a), ptr(b), ptr(c), ptr(out), a.numel(), stream()), "add")
    return out


def colsum(x, C=None):
    cs = x.shape[-1]
    C = C or cs
    out = torch.zeros(C, dtype=torch.float32, device=x.device)
    check(_lib.lib().hd_colsum(ptr(x), ptr(out), x.numel() // cs, C, cs, stream()), "colsum")
    return out


def head_backward(dlogits, feat, wp, cout, extra=None):
    """dlogits: fp32 (B, cout, H, W) view with channel stride H*W. Returns dfeat (NHWC bf16), dW [cout,128], dbias."""
    n, h, w, _ = feat.shape
    assert dlogits.stride(1) == h * w and dlogits.stride(3) == 1
    dfeat = torch.empty_like(feat)
    dw = torch.zeros((cout, 128), dtype=torch.float32, device=feat.device)
    db = torch.zeros(cout, dtype=torch.float32, device=feat.device)
    check(_lib.lib().hd_head_backward(ptr(dlogits), dlogits.stride(0), ptr(extra),
                                      extra.shape[-1] if extra is not None else 0, ptr(feat), ptr(wp), ptr(dfeat),
                                      ptr(dw), ptr(db), n, h, w, cout, stream()), "head_backward")
    return dfeat, dw, db
