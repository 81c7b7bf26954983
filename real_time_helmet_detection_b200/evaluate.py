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


class _Captured:
    __slots__ = ("graph", "x", "boxes", "clss", "scores", "counts")


class Prediction(torch.nn.Module):
    """`cuda_graph=True` (an addition to the reference signature): the eval forward (~50 launches on
    two streams) and the two decode launches are recorded once per input shape into a CUDA graph and replayed, so a
    batch-1 prediction costs one graph launch + one copy-in + the count read-back instead of ~55 host-side launches."""

    def __init__(self, network, topk, scale_factor, conf_th, nms, nms_th, normalized_coord=False, cuda_graph=False):
        super().__init__()
        self.network = network
        self.topk = topk
        self.scale_factor = scale_factor
        self.conf_th = conf_th
        self.nms = nms
        self.nms_th = nms_th
        self.normalized_coord = normalized_coord
        self.cuda_graph = cuda_graph
        self._graphs = {}

    def _decode_device(self, batch_output):
        """Enqueue-only part of decode(): (boxes (B,S*k,4), classes (B,S*k) i64, scores (B,S*k), counts (B,) i32)."""
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
        return _decode_call(heat, off, wh, strides, B, S, C, H, W, self.topk, self.scale_factor, self.conf_th,
                            self.nms_th, self.normalized_coord, True, True)

    @staticmethod
    def _to_lists(boxes, clss, scores, counts):
        n = counts.tolist()                                   # the only host sync
        B = len(n)
        return ([boxes[b, :n[b]] for b in range(B)], [clss[b, :n[b]] for b in range(B)],
                [scores[b, :n[b]] for b in range(B)])

    def decode(self, batch_output):
        """batch_output: (B, S, num_cls+4, H, W) raw logits -> three lists of per-image tensors."""
        return self._to_lists(*self._decode_device(batch_output))

    def _capture(self, x):
        if self.network.training:
            raise RuntimeError("Prediction(cuda_graph=True) records the eval-mode forward: call network.eval() first")
        c = _Captured()
        c.x = x.detach().clone()
        cur = torch.cuda.current_stream(x.device)
        side = torch.cuda.Stream(device=x.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side), torch.no_grad():       # eager warm-up: workspace, streams, pinned job-table slots
            for _ in range(2):
                self._decode_device(self.network(c.x))
        cur.wait_stream(side)
        torch.cuda.synchronize(x.device)
        c.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(c.graph), torch.no_grad():
            c.boxes, c.clss, c.scores, c.counts = self._decode_device(self.network(c.x))
        return c

    def forward(self, x):
        ''' x: input tensor (b, c, h, w) '''
        if self.cuda_graph:
            _lib.require_cuda(x, "Prediction input")
            key = (tuple(x.shape), x.dtype, x.device)
            c = self._graphs.get(key)
            if c is None:
                c = self._graphs[key] = self._capture(x)
            c.x.copy_(x)
            c.graph.replay()
            # the graph's output buffers are overwritten by the next replay: hand out copies, like the fresh tensors
            # the reference returns
            return self._to_lists(c.boxes.clone(), c.clss.clone(), c.scores.clone(), c.counts)
        with torch.no_grad():
            batch_output = self.network(x)       # b, n, num_cls+4, h, w
        return self.decode(batch_output)
