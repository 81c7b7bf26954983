"""ORACLE (test infrastructure): CenterNet training losses restated on CPU.

Follows /root/reference/loss.py:
  FocalLoss.forward     :58-69   eps=1e-7 inside both logs, mask (B,1,h,w) broadcast over classes,
                                 per-sample sum -> batch mean -> / clamp(mask.sum(), 1)
  NormedL1Loss.forward  :46-50   |pred*mask - gt*mask| summed per sample -> batch mean -> / clamp(mask.sum(), 1)
  LossCalculator.forward:18-32   total = hm*w_hm + offset*w_off + size*w_size
and the caller-side head activation of /root/reference/train.py:105-111 (split [C,2,2], sigmoid on the heatmap,
sigmoid on offset/size only with --normalized-coord).
"""
from __future__ import annotations

import torch


def focal_loss(pred, gt, mask, alpha=2.0, beta=4.0, eps=1e-7):
    neg_w = torch.pow(1.0 - gt, beta)
    pos = torch.log(pred + eps) * torch.pow(1.0 - pred, alpha) * mask
    neg = torch.log(1.0 - pred + eps) * torch.pow(pred, alpha) * neg_w * (1.0 - mask)
    pos = pos.sum(dim=(1, 2, 3)).mean()
    neg = neg.sum(dim=(1, 2, 3)).mean()
    num_pos = mask.sum().clamp(1, 1e30)
    return -(pos + neg) / num_pos


def normed_l1(pred, gt, mask):
    per = torch.abs(pred * mask - gt * mask).sum(dim=(1, 2, 3)).mean()
    return per / mask.sum().clamp(1, 1e30)


def losses_from_logits(logits, ghm, goff, gsize, mask, num_cls=2, hm_weight=1.0, offset_weight=1.0,
                       size_weight=0.1, alpha=2.0, beta=4.0, normalized_coord=False):
    """logits: (B, num_cls+4, h, w) raw head output of ONE stack. Returns (hm, offset, size, total)."""
    phm, poff, psize = logits.split([num_cls, 2, 2], dim=1)
    phm = torch.sigmoid(phm)
    if normalized_coord:
        poff, psize = torch.sigmoid(poff), torch.sigmoid(psize)
    hm = focal_loss(phm, ghm, mask, alpha, beta)
    off = normed_l1(poff, goff, mask)
    size = normed_l1(psize, gsize, mask)
    return hm, off, size, hm * hm_weight + off * offset_weight + size * size_weight
