"""ORACLE (test infrastructure, never shipped on the product path).

CPU fp32 restatement of the reference's stacked-hourglass forward pass as a pure function of a
state_dict, used as the checker for the sm_100a kernels. Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package.

Follows /root/reference/hourglass.py:
  Convolution  :94-108   act(bn(conv(x))), pad=(k-1)//2
  Residual     :111-127  relu(conv2(conv1(x)) + skip(x))
  Hourglass    :130-156  up1(x) + upsample2x(low3(low2(low1(maxpool2x2(x)))))
  PreLayer     :159-173  conv7x7s2+BN+ReLU, Residual(64,128), maxpool, Residual, Residual
  Neck         :176-186  conv1x1+bias+BN+ReLU, Residual
  Head         :189-195  conv1x1+bias
  StackedHourglass.forward :223-237

Parity is pinned: tests/test_oracle_golden.py checks this restatement against outputs of the unmodified
reference modules run in the build container (tests/golden/make_golden.py, fixtures in tests/golden/).

`emulate_bf16=True` additionally rounds every tensor the CUDA path stores in bf16 (conv operands and the
materialised activations) to bf16 at the same points, keeping fp32 accumulation and fp32 BN statistics;
that is the numerics contract of the B200 path (DESIGN.md "numerics").
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def _q(t: torch.Tensor, on: bool) -> torch.Tensor:
    """bf16 storage rounding with a straight-through gradient."""
    if not on:
        return t
    return t + (t.to(torch.bfloat16).to(t.dtype) - t).detach()


class _Ctx:
    def __init__(self, sd: Dict[str, torch.Tensor], training: bool, emulate_bf16: bool,
                 new_stats: Optional[Dict[str, torch.Tensor]], taps: Optional[Dict[str, torch.Tensor]]):
        self.sd = sd
        self.training = training
        self.q = emulate_bf16
        self.new_stats = new_stats
        self.taps = taps


def _conv(ctx: _Ctx, prefix: str, x: torch.Tensor, k: int, stride: int = 1) -> torch.Tensor:
    w = _q(ctx.sd[prefix + ".convolution.weight"], ctx.q)
    b = ctx.sd.get(prefix + ".convolution.bias")
    return F.conv2d(x, w, b, stride=stride, padding=(k - 1) // 2)


def _bn(ctx: _Ctx, prefix: str, y: torch.Tensor) -> torch.Tensor:
    """BatchNorm2d(affine, track_running_stats), eps 1e-5, momentum 0.1 (hourglass.py:103).

    Train mode: biased batch variance normalises; the unbiased one goes to running_var.
    In bf16 emulation the statistics come from the fp32 conv output and are applied to its bf16-rounded copy,
    like the conv-epilogue statistics of the CUDA path.
    """
    g, b = ctx.sd[prefix + ".bn.weight"], ctx.sd[prefix + ".bn.bias"]
    if ctx.training:
        mean = y.mean(dim=(0, 2, 3))
        var = y.var(dim=(0, 2, 3), unbiased=False)
        if ctx.new_stats is not None:
            n = y.numel() // y.shape[1]
            rm, rv = ctx.sd[prefix + ".bn.running_mean"], ctx.sd[prefix + ".bn.running_var"]
            ctx.new_stats[prefix + ".bn.running_mean"] = ((1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean).detach()
            ctx.new_stats[prefix + ".bn.running_var"] = (
                (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * var * (n / max(n - 1, 1))).detach()
    else:
        mean, var = ctx.sd[prefix + ".bn.running_mean"], ctx.sd[prefix + ".bn.running_var"]
    scale = g * torch.rsqrt(var + BN_EPS)
    shift = b - mean * scale
    yq = _q(y, ctx.q)
    return yq * scale[None, :, None, None] + shift[None, :, None, None]


def _conv_bn_act(ctx: _Ctx, prefix: str, x: torch.Tensor, k: int, relu: bool, stride: int = 1) -> torch.Tensor:
    y = _bn(ctx, prefix, _conv(ctx, prefix, x, k, stride))
    if relu:
        y = F.relu(y)
    return _q(y, ctx.q)


def _residual(ctx: _Ctx, prefix: str, x: torch.Tensor) -> torch.Tensor:
    z1 = _conv_bn_act(ctx, prefix + ".conv1", x, 3, relu=True)
    y2 = _bn(ctx, prefix + ".conv2", _conv(ctx, prefix + ".conv2", z1, 3))
    if (prefix + ".skip.convolution.weight") in ctx.sd:
        s = _bn(ctx, prefix + ".skip", _conv(ctx, prefix + ".skip", x, 1))
    else:
        s = x
    out = _q(F.relu(y2 + s), ctx.q)
    if ctx.taps is not None:
        ctx.taps[prefix] = out.detach()
    return out


def _hourglass(ctx: _Ctx, prefix: str, x: torch.Tensor, depth: int) -> torch.Tensor:
    up1 = _residual(ctx, prefix + ".up1", x)
    low = F.max_pool2d(x, 2, 2)
    low = _residual(ctx, prefix + ".low1", low)
    if depth > 1:
        low = _hourglass(ctx, prefix + ".low2", low, depth - 1)
    else:
        low = _residual(ctx, prefix + ".low2", low)
    low = _residual(ctx, prefix + ".low3", low)
    out = _q(up1 + F.interpolate(low, scale_factor=2, mode="nearest"), ctx.q)
    if ctx.taps is not None:
        ctx.taps[prefix] = out.detach()
    return out


def num_stacks(sd: Dict[str, torch.Tensor]) -> int:
    s = 0
    while f"head_lst.{s}.layer.convolution.weight" in sd:
        s += 1
    return s


def stacked_hourglass_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, training: bool = False,
                              emulate_bf16: bool = False, new_stats: Optional[Dict[str, torch.Tensor]] = None,
                              taps: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
    """x: (B,3,H,W) fp32, H and W multiples of 64 -> (B, S, num_cls+4, H/4, W/4) raw logits."""
    ctx = _Ctx(sd, training, emulate_bf16, new_stats, taps)
    # the stem reads the fp32 image directly on the CUDA path too (no bf16 rounding of the image)
    stem_w = "pre_layer.layers.0"
    y = F.conv2d(x, sd[stem_w + ".convolution.weight"], sd[stem_w + ".convolution.bias"], stride=2, padding=3)
    h = _q(F.relu(_bn(ctx, stem_w, y)), ctx.q)
    h = _residual(ctx, "pre_layer.layers.1", h)
    h = F.max_pool2d(h, 2, 2)
    h = _residual(ctx, "pre_layer.layers.3", h)
    h = _residual(ctx, "pre_layer.layers.4", h)
    preds: List[torch.Tensor] = []
    S = num_stacks(sd)
    for i in range(S):
        hg = _hourglass(ctx, f"hourglass_lst.{i}", h, 4)
        feat = _conv_bn_act(ctx, f"neck_lst.{i}.layers.1", hg, 1, relu=True)
        feat = _residual(ctx, f"neck_lst.{i}.layers.2", feat)
        pred = _conv(ctx, f"head_lst.{i}.layer", feat, 1)
        preds.append(pred)
        if i < S - 1:
            mf = _conv(ctx, f"merge_feature.{i}", feat, 1)
            mp = _conv(ctx, f"merge_prediction.{i}", _q(pred, ctx.q), 1)
            h = _q(_q(h + mf, ctx.q) + mp, ctx.q)
    return torch.stack(preds, dim=1)
