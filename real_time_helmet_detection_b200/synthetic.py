"""Seeded synthetic inputs of the benchmark workloads: there is no dataset on the GPU box.

`synthetic_targets` pushes 1-5 random boxes per image through the host-side GT encoder (transform.box2hm);
`synthetic_head` builds the decode input of config 5 (a -6 logit floor with Gaussian blobs).
"""
from __future__ import annotations

import numpy as np

from .transform import box2hm


def synthetic_boxes(batch, imsize=512, num_cls=2, max_boxes=5):
    """The box lists behind `synthetic_targets`: per image (boxes [[x0,y0,x1,y1], ...], labels [...])."""
    out = []
    for b in range(batch):
        rs = np.random.RandomState(b)
        nb = rs.randint(1, max_boxes + 1)
        boxes, labels = [], []
        for _ in range(nb):
            x0, y0 = rs.uniform(0, 0.7 * imsize, 2)
            bw, bh = rs.uniform(0.05, 0.3, 2) * imsize
            boxes.append([x0, y0, min(x0 + bw, imsize - 1), min(y0 + bh, imsize - 1)])
            labels.append(int(rs.randint(0, num_cls)))
        out.append((boxes, labels))
    return out


def synthetic_targets(batch, imsize=512, num_cls=2, scale_factor=4, max_boxes=5):
    outs = [[], [], [], []]
    for boxes, labels in synthetic_boxes(batch, imsize, num_cls, max_boxes):
        for lst, arr in zip(outs, box2hm(boxes, labels, (imsize, imsize), scale_factor, num_cls)):
            lst.append(arr)
    return tuple(np.stack(o) for o in outs)


def synthetic_head(S=1, H=128, W=128, num_cls=2, seed=0, blobs=60):
    rs = np.random.RandomState(seed)
    out = np.zeros((1, S, num_cls + 4, H, W), np.float32)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    for s in range(S):
        heat = (-6 + 0.5 * rs.randn(num_cls, H, W)).astype(np.float32)
        for _ in range(blobs):
            c = rs.randint(0, num_cls)
            cy, cx = rs.uniform(0, H - 1), rs.uniform(0, W - 1)
            amp, sig = rs.uniform(4, 10), rs.uniform(1, 4)
            blob = (-6 + (amp + 6) * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * sig * sig))).astype(np.float32)
            heat[c] = np.maximum(heat[c], blob)
        out[0, s, :num_cls] = heat
        out[0, s, num_cls:num_cls + 2] = rs.uniform(0, 1, (2, H, W)).astype(np.float32)
        out[0, s, num_cls + 2:] = rs.uniform(4, 24, (2, H, W)).astype(np.float32)
    return out
