"""Training losses — drop-in for the reference's loss.py (`LossCalculator`, loss.py:6-40).

Same constructor, same call signature `(phm, poff, psize, ghm, goff, gsize, mask) -> differentiable scalar`, same
`.log` dict (keys hm / offset / size / total, lists of floats) and `get_log(length)` string, but the ~25 elementwise
launches + 7 reductions + 4 `.item()` syncs of the reference collapse into one fused sm_100a forward kernel and one
backward kernel (csrc/loss.cu), and the per-call host syncs disappear: loss values stay on the device and are only
copied to the host when `.log` / `get_log()` is read.

`forward_logits` is the fused entry the native train step uses: it takes the raw head output of one stack and folds
the caller-side split + sigmoid (train.py:107-111) into the kernel as well.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from ._lib import ptr, stream, check

_KEYS = ("hm", "offset", "size", "total")


def _plane_view(t: torch.Tensor) -> torch.Tensor:
    """fp32 (B, C, H, W) tensor whose channel planes are dense (stride C -> H*W, H -> W, W -> 1)."""
    if t.dtype != torch.float32:
        t = t.float()
    _, _, h, w = t.shape
    if t.stride(3) != 1 or t.stride(2) != w or t.stride(1) != h * w:
        t = t.contiguous()
    return t


class _FusedLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hm, off, size, ghm, goff, gsize, mask, cfg):
        alpha, beta, w_hm, w_off, w_size, from_logits, sigmoid_reg = cfg
        _lib.require_cuda(hm, "prediction")
        hm, off, size = _plane_view(hm), _plane_view(off), _plane_view(size)
        ghm, goff, gsize, mask = (t.float().contiguous() for t in (ghm, goff, gsize, mask))
        B, C, H, W = hm.shape
        buf = torch.empty(10, dtype=torch.float32, device=hm.device)   # [0:5] results, [5:10] reduction scratch
        with torch.cuda.device(hm.device):
            check(_lib.lib().hd_loss_forward(ptr(hm), hm.stride(0), ptr(off), off.stride(0), ptr(size), size.stride(0),
                                             ptr(ghm), ptr(goff), ptr(gsize), ptr(mask), B, C, H, W, alpha, beta, w_hm,
                                             w_off, w_size, int(from_logits), int(sigmoid_reg), ptr(buf[5:]), ptr(buf),
                                             stream(hm.device)), "loss_forward")
        ctx.save_for_backward(hm, off, size, ghm, goff, gsize, mask, buf)
        ctx.cfg = cfg
        values = buf[:4]
        ctx.mark_non_differentiable(values)
        return buf[3].clone(), values

    @staticmethod
    def backward(ctx, g_total, _g_values):
        hm, off, size, ghm, goff, gsize, mask, buf = ctx.saved_tensors
        alpha, beta, w_hm, w_off, w_size, from_logits, sigmoid_reg = ctx.cfg
        B, C, H, W = hm.shape
        d_hm, d_off, d_size = torch.empty_like(hm), torch.empty_like(off), torch.empty_like(size)
        g = g_total.detach().float().contiguous()
        with torch.cuda.device(hm.device):
            check(_lib.lib().hd_loss_backward(ptr(hm), hm.stride(0), ptr(off), off.stride(0), ptr(size), size.stride(0),
                                              ptr(ghm), ptr(goff), ptr(gsize), ptr(mask), B, C, H, W, alpha, beta, w_hm,
                                              w_off, w_size, int(from_logits), int(sigmoid_reg), ptr(buf), ptr(g),
                                              ptr(d_hm), d_hm.stride(0), ptr(d_off), d_off.stride(0), ptr(d_size),
                                              d_size.stride(0), stream(hm.device)), "loss_backward")
        return d_hm, d_off, d_size, None, None, None, None, None


class LossCalculator(nn.Module):
    def __init__(self, hm_weight, offset_weight, size_weight, focal_alpha, focal_beta):
        super().__init__()
        self._log = {k: [] for k in _KEYS}
        self._pending = []            # device tensors (4,) not yet copied to the host
        self.hm_weight = hm_weight
        self.offset_weight = offset_weight
        self.size_weight = size_weight
        self.focal_alpha = focal_alpha
        self.focal_beta = focal_beta

    # -- `.log` keeps the reference's type (dict of float lists; pickled into checkpoints, train.py:82,197)
    @property
    def log(self):
        self._flush()
        return self._log

    @log.setter
    def log(self, value):
        self._pending = []
        self._log = value

    def _flush(self):
        if self._pending:
            vals = torch.stack(self._pending).cpu().tolist()      # one D2H copy for all pending steps
            self._pending = []
            for row in vals:
                for k, v in zip(_KEYS, row):
                    self._log[k].append(v)

    def _cfg(self, from_logits, sigmoid_reg):
        return (float(self.focal_alpha), float(self.focal_beta), float(self.hm_weight), float(self.offset_weight),
                float(self.size_weight), bool(from_logits), bool(sigmoid_reg))

    def _record(self, values):
        self._pending.append(values.detach())
        if len(self._pending) >= 4096:
            self._flush()

    def forward(self, phm, poff, psize, ghm, goff, gsize, mask):
        total, values = _FusedLoss.apply(phm, poff, psize, ghm, goff, gsize, mask, self._cfg(False, False))
        self._record(values)
        return total

    def forward_logits(self, logits, ghm, goff, gsize, mask, num_cls=None, normalized_coord=False):
        """logits: (B, num_cls+4, H, W) raw head output of one stack (a view of the network output is fine)."""
        C = logits.shape[1] - 4 if num_cls is None else num_cls
        hm, off, size = logits[:, :C], logits[:, C:C + 2], logits[:, C + 2:C + 4]
        total, values = _FusedLoss.apply(hm, off, size, ghm, goff, gsize, mask, self._cfg(True, normalized_coord))
        self._record(values)
        return total

    def get_log(self, length=100):
        log = self.log
        parts = []
        for key in _KEYS:
            if len(log[key]) < length:
                length = len(log[key])
            parts.append('%s: %5.2f' % (key, sum(log[key][-length:]) / length))
        return ', '.join(parts)
