"""ORACLE (test infrastructure): ground-truth encoder restated in numpy.

Follows /root/reference/transform.py:4-70 (`box2hm`, `gaussian2D`, `draw_gaussian`): boxes in input pixels are
divided by scale_factor; the integer centre cell gets mask=1, the fractional centre offset and the box size
(optionally normalised); each class heat-map receives max(current, gaussian) with radius = half the box diagonal
(in cells), sigma = radius/3, the patch clipped at the map border. Used to synthesise realistic GT for the loss
parity tests and the benchmark (the reference's dataloader is out of scope).
"""
from __future__ import annotations

import numpy as np


def encode_boxes(boxes, labels, imsize, scale_factor=4, num_cls=2, normalized=False):
    w, h = imsize[0] // scale_factor, imsize[1] // scale_factor
    heat = np.zeros((num_cls, h, w), np.float32)
    off = np.zeros((2, h, w), np.float32)
    size = np.zeros((2, h, w), np.float32)
    mask = np.zeros((1, h, w), np.float32)
    for box, label in zip(boxes or [], labels or []):
        if box is None:
            continue
        x0, y0, x1, y1 = (float(v) / scale_factor for v in box)
        cx, cy = (x0 + x1) / 2, (y0 + y1) / 2
        ix, iy = int(cx), int(cy)
        mask[:, iy, ix] = 1.0
        ox, oy = cx - ix, cy - iy
        sx, sy = x1 - x0, y1 - y0
        if normalized:
            ox, oy = ox / scale_factor, oy / scale_factor
            sx, sy = sx / w, sy / h
        off[:, iy, ix] = (ox, oy)
        size[:, iy, ix] = (sx, sy)
        radius = ((cx - x0) ** 2 + (cy - y0) ** 2) ** 0.5
        r = int(radius)
        # (2*int(radius)+1)^2 gaussian patch with sigma = radius/3 (float radius), as the reference builds it
        ax = np.arange(-r, r + 1, dtype=np.float64)
        sig = radius / 3
        with np.errstate(divide="ignore", invalid="ignore"):
            g = np.exp(-(ax[None, :] ** 2 + ax[:, None] ** 2) / (2 * sig * sig))
        left, right = min(ix, r), min(w - ix, r + 1)
        top, bottom = min(iy, r), min(h - iy, r + 1)
        dst = heat[label][iy - top:iy + bottom, ix - left:ix + right]
        np.maximum(dst, g[r - top:r + bottom, r - left:r + right].astype(np.float32), out=dst)
    return heat, off, size, mask


def synthetic_targets(batch, imsize=512, num_cls=2, scale_factor=4, max_boxes=5):
    """Seeded synthetic GT as specified in BASELINE.md / SURVEY.md 8(d): 1-5 random boxes per image."""
    outs = [[], [], [], []]
    for b in range(batch):
        rs = np.random.RandomState(b)
        nb = rs.randint(1, max_boxes + 1)
        boxes, labels = [], []
        for _ in range(nb):
            x0, y0 = rs.uniform(0, 0.7 * imsize, 2)
            bw, bh = rs.uniform(0.05, 0.3, 2) * imsize
            boxes.append([x0, y0, min(x0 + bw, imsize - 1), min(y0 + bh, imsize - 1)])
            labels.append(int(rs.randint(0, num_cls)))
        for lst, arr in zip(outs, encode_boxes(boxes, labels, (imsize, imsize), scale_factor, num_cls)):
            lst.append(arr)
    return tuple(np.stack(o) for o in outs)
