"""Inference wrapper — drop-in for the reference's `Prediction` module (evaluate.py:114-182).

Same constructor and `forward(x) -> (box_lst, cls_lst, score_lst)` contract (lists indexed per image, score-descending
tensors), but the per-image / per-stack Python loops with ~40 launches and >= 4 host syncs per stack become one
fused kernel launch for the whole batch (csrc/decode.cu: sigmoid, peak test, top-k, gather, threshold, cross-stack
class-agnostic NMS) followed by a single device->host read of the per-image box counts.
"""
from __future__ import annotations

import torch

from . import _lib
from .transform import _decode_call


class Prediction(torch.nn.Module):
    def __init__(self, network, topk, scale_factor, conf_th, nms, nms_th, normalized_coord=False):
        super().__init__()
        self.network = network
        self.topk = topk
        self.scale_factor = scale_factor
        self.conf_th = conf_th
        self.nms = nms
        self.nms_th = nms_th
        self.normalized_coord = normalized_coord

    def decode(self, batch_output):
        """batch_output: (B, S, num_cls+4, H, W) raw logits -> three lists of per-image tensors."""
        if self.nms == 'soft-nms':
            raise NotImplementedError('soft-nms is a CPU-only O(N^2) Python loop in the reference (evaluate.py:184-243) '
                                      'and is outside the B200 hot path; use nms="nms"')
        if self.nms != 'nms':
            raise NotImplementedError('Not expected nms algorithm: %s' % self.nms)
        _lib.require_cuda(batch_output, "network output")
        out = batch_output.detach().float().contiguous()
        B, S, O, H, W = out.shape
        C = O - 4
        hw = H * W
        heat, off, wh = out, out[:, :, C:], out[:, :, C + 2:]
        strides = ((S * O * hw, O * hw),) * 3
        boxes, clss, scores, counts = _decode_call(heat, off, wh, strides, B, S, C, H, W, self.topk,
                                                   self.scale_factor, self.conf_th, self.nms_th,
                                                   self.normalized_coord, True, True)
        n = counts.tolist()                                   # the only host sync
        return ([boxes[b, :n[b]] for b in range(B)], [clss[b, :n[b]] for b in range(B)],
                [scores[b, :n[b]] for b in range(B)])

    def forward(self, x):
        ''' x: input tensor (b, c, h, w) '''
        with torch.no_grad():
            batch_output = self.network(x)       # b, n, num_cls+4, h, w
        return self.decode(batch_output)
