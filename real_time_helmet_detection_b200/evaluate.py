"""Inference wrapper — drop-in for the reference's `Prediction` module (evaluate.py:114-182).

Same constructor and `forward(x) -> (box_lst, cls_lst, score_lst)` contract (lists indexed per image, score-descending
tensors), but the per-image / per-stack Python loops with ~40 launches and >= 4 host syncs per stack become one
fused kernel launch for the whole batch (csrc/decode.cu: sigmoid, peak test, top-k, gather, threshold, cross-stack
class-agnostic NMS) followed by a single device->host read of the per-image box counts.
"""
from __future__ import annotations

import os
import pickle
import time

import numpy as np
import torch

from . import _lib
from .transform import _decode_buffers, _decode_call, _decode_scratch


class _Captured:
    # `workspace` pins the network's HBM arena the graph's kernels were recorded against: the graph bakes raw pointers
    # into it, so it must neither be freed nor silently swapped while the graph can still be replayed
    __slots__ = ("graph", "x", "boxes", "clss", "scores", "counts", "workspace")


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
        self._scratch = {}         # decode scratch per (device, B, S, C, H, W): initialised once, self-cleaning afterwards

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
        key = (out.device, B, S, C, H, W)
        scratch = self._scratch.get(key)
        if scratch is None:
            scratch = self._scratch[key] = _decode_scratch(out.device, B, S, C, H, W)
        bufs = _decode_buffers(out.device, B, S, C, H, W, self.topk, scratch=scratch)
        return _decode_call(heat, off, wh, strides, B, S, C, H, W, self.topk, self.scale_factor, self.conf_th,
                            self.nms_th, self.normalized_coord, True, True, bufs=bufs)

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
        # a PRIVATE arena per graph: the recorded kernels keep raw pointers into it, so it must outlive every other
        # forward of the network (other shapes, eager calls, training steps drop and re-allocate the network's own)
        self.network._workspace, self.network._ws_key = None, None
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
        c.workspace = self.network._workspace
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


# ---------------------------------------------------------------------------------------------------------------------
# Evaluation loop + on-disk formats of the reference (evaluate.py:40-112): `prediction_results.pickle`
# ({image filename: float64 array (n, 6) = class, score, xmin, ymin, xmax, ymax in ORIGINAL image pixels}) and one
# `txt/<name>.txt` per image ("%d %f %d %d %d %d" per detection) - the input of the mAP scorer (mAP/main.py).
def resize_box_to_original_scale(boxes, original_size, transformed_size):
    """(n,4) xyxy boxes in network-input pixels -> original image pixels (evaluate.py:101-112), vectorised."""
    rw = original_size[0] / transformed_size[0]
    rh = original_size[1] / transformed_size[1]
    return np.asarray(boxes, dtype=np.float64).reshape(-1, 4) * np.array([rw, rh, rw, rh])


def evaluate_step(dataloader, predictor, device, args):
    """Same contract as the reference's `evaluate_step` (evaluate.py:59-99). One prediction call per batch; the
    rescale to the original size and the (class, score, box) packing happen on the device, so every image costs one
    small device->host copy instead of three."""
    predictor.eval()
    results = {}
    t_data = t_fwd = 0.0
    n_batches = 0
    tic = time.time()
    for image, _gt_heatmap, _gt_offset, _gt_size, _gt_mask, gt_dict in dataloader:
        t_data += time.time() - tic
        tic = time.time()
        box_lst, cls_lst, score_lst = predictor(image.to(device))
        t_fwd += time.time() - tic
        n_batches += 1
        for b in range(image.shape[0]):
            ann = gt_dict[b]['annotation']
            rw = int(ann['size']['width']) / args.imsize
            rh = int(ann['size']['height']) / args.imsize
            boxes, clss, scores = box_lst[b], cls_lst[b], score_lst[b]
            if boxes.shape[0] != 0:
                scale = torch.tensor([rw, rh, rw, rh], dtype=torch.float64, device=boxes.device)
                packed = torch.cat([clss.to(torch.float64)[:, None], scores.to(torch.float64)[:, None],
                                    boxes.to(torch.float64) * scale], dim=1)
                results[ann['filename']] = packed.cpu().numpy()
            else:
                results[ann['filename']] = np.zeros((0, 6))
        tic = time.time()
    n_batches = max(n_batches, 1)
    print('%s: Evaluation, Time(ms) [data: %6.2f, forward: %6.2f]'
          % (time.ctime(), t_data / n_batches * 1000, t_fwd / n_batches * 1000))
    return results


def save_predictions(predictions, save_path):
    """The two artefacts `evaluate.py:40-54` writes under --save-path."""
    os.makedirs(os.path.join(save_path, 'txt'), exist_ok=True)
    with open(os.path.join(save_path, 'prediction_results.pickle'), 'wb') as f:
        pickle.dump(predictions, f, protocol=pickle.HIGHEST_PROTOCOL)
    for filename, pred in predictions.items():
        with open(os.path.join(save_path, 'txt', os.path.splitext(filename)[0] + '.txt'), 'w') as f:
            for row in pred:
                f.write('%d %f %d %d %d %d\n' % (row[0], row[1], row[2], row[3], row[4], row[5]))
