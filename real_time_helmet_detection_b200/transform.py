"""CenterNet decode — drop-in for the reference's transform.py (`hm2box`, transform.py:73-110).

`hm2box` keeps the reference signature and return types ((k',4) fp32 boxes, (k',) int64 classes, (k',) fp32 scores,
score-descending) but runs as ONE fused kernel (csrc/decode.cu): 3x3 peak test, joint top-k, gather, box assembly
and the confidence threshold, with a single device->host read (the survivor count, needed for the output shapes).

`box2hm` is the host-side ground-truth encoder used by the dataloader (transform.py:4-70); it is data preparation, not
part of the GPU hot path, and is provided so the reference's data.py keeps importing from one module.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import ptr, stream, check


def _decode_scratch(dev, B, S, C, H, W):
    """Scratch of the decode kernels with its candidate counters zeroed (needed once per buffer: every call leaves them
    zero again, csrc/decode.cu)."""
    scratch = torch.empty((_lib.lib().hd_decode_scratch_bytes(B, S, C, H, W),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().hd_decode_scratch_init(ptr(scratch), B, S, stream(dev)), "decode_scratch_init")
    return scratch


def _decode_buffers(dev, B, S, C, H, W, topk, scratch=None):
    n = S * topk
    return (torch.empty((B, n, 4), dtype=torch.float32, device=dev), torch.empty((B, n), dtype=torch.int64, device=dev),
            torch.empty((B, n), dtype=torch.float32, device=dev), torch.empty((B,), dtype=torch.int32, device=dev),
            scratch if scratch is not None else _decode_scratch(dev, B, S, C, H, W))


def _decode_call(heat, off, wh, strides, B, S, C, H, W, topk, scale_factor, conf_th, nms_th, normalized,
                 apply_sigmoid, do_nms, bufs=None):
    L = _lib.lib()
    if topk > C * H * W:
        raise RuntimeError("selected index k out of range")
    boxes, clss, scores, counts, scratch = bufs if bufs is not None else _decode_buffers(heat.device, B, S, C, H, W, topk)
    (bh, sh), (bo, so), (bw, sw) = strides
    with torch.cuda.device(heat.device):
        check(L.hd_decode_nms(ptr(heat), bh, sh, ptr(off), bo, so, ptr(wh), bw, sw, B, S, C, H, W, int(topk),
                              float(scale_factor), float(conf_th), float(nms_th), int(bool(normalized)),
                              int(bool(apply_sigmoid)), int(bool(do_nms)), ptr(scratch), ptr(boxes), ptr(clss),
                              ptr(scores), ptr(counts), stream(heat.device)), "decode_nms")
    return boxes, clss, scores, counts


def hm2box(heatmap, offset, wh, scale_factor=4, topk=10, conf_th=0.3, normalized=False):
    _lib.require_cuda(heatmap, "heatmap")
    heatmap, offset, wh = (t.detach().float().contiguous() for t in (heatmap, offset, wh))
    C, H, W = heatmap.shape[-3:]
    boxes, clss, scores, counts = _decode_call(heatmap, offset, wh, ((0, 0), (0, 0), (0, 0)), 1, 1, C, H, W, topk,
                                               scale_factor, conf_th, 0.0, normalized, False, False)
    n = int(counts.item())
    return boxes[0, :n], clss[0, :n], scores[0, :n]


def box2hm(boxes, labels, imsize, scale_factor=4, num_cls=2, normalized=False):
    """Boxes (input pixels) -> (heat (C,h,w), offset (2,h,w), size (2,h,w), mask (1,h,w)) float32 numpy arrays."""
    width, height = imsize[0] // scale_factor, imsize[1] // scale_factor
    heat = np.zeros((num_cls, height, width), dtype=np.float32)
    offset = np.zeros((2, height, width), dtype=np.float32)
    size = np.zeros((2, height, width), dtype=np.float32)
    mask = np.zeros((1, height, width), dtype=np.float32)
    if boxes is None:
        return heat, offset, size, mask
    for box, label in zip(boxes, labels):
        if box is None:
            continue
        x0, y0, x1, y1 = [v / scale_factor for v in box]
        cx, cy = (x0 + x1) / 2, (y0 + y1) / 2
        ix, iy = int(cx), int(cy)
        mask[:, iy, ix] = 1.0
        dx, dy = cx - ix, cy - iy
        bw, bh = x1 - x0, y1 - y0
        if normalized:
            dx, dy = dx / scale_factor, dy / scale_factor
            bw, bh = bw / width, bh / height
        offset[:, iy, ix] = (dx, dy)
        size[:, iy, ix] = (bw, bh)
        radius = float(np.hypot(cx - x0, cy - y0))
        _splat_gaussian(heat[label], ix, iy, radius)
    return heat, offset, size, mask


def _splat_gaussian(plane, ix, iy, radius):
    r = int(radius)
    h, w = plane.shape
    ys, xs = np.ogrid[-r:r + 1, -r:r + 1]
    sigma = radius / 3
    with np.errstate(divide="ignore", invalid="ignore"):
        g = np.exp(-(xs * xs + ys * ys) / (2 * sigma * sigma))
    l, rr = min(ix, r), min(w - ix, r + 1)
    t, b = min(iy, r), min(h - iy, r + 1)
    dst = plane[iy - t:iy + b, ix - l:ix + rr]
    np.maximum(dst, g[r - t:r + b, r - l:r + rr], out=dst)
