"""ORACLE (test infrastructure): CenterNet decode + NMS restated in numpy (fp32 arithmetic, integer indices).

Follows
  /root/reference/transform.py:73-110  `hm2box`   3x3 max-pool peak test by equality (plateaus all pass), joint
        top-k over the flattened (C,H,W) peak map, index -> (cls, y, x), gather offset / size, box assembly
        ((x + xoff) -+ w/2) * scale in that operation order, threshold `score >= conf_th` applied after top-k;
  /root/reference/evaluate.py:126-182  `Prediction.forward` / `nonmaximum_supression`: per image and stack the
        head split [C,2,2] + sigmoid(heatmap) (+ sigmoid(offset/wh) with normalized_coord), candidates of all
        stacks concatenated, then CLASS-AGNOSTIC hard NMS with torchvision.ops.nms semantics.

torchvision.ops.nms (torchvision 0.26 csrc/ops/cpu/nms_kernel.cpp / cuda/nms_kernel.cu; pinned upstream in prose as
v0.7.0, README.md:13; not vendored): boxes sorted by descending score; a box is suppressed by an earlier kept box
when inter / (area_a + area_b - inter) > threshold, areas without "+1"; the result is returned in score order.

Tie-breaking, which torch leaves implementation-defined, is FIXED here and in the CUDA kernels: equal scores are
ordered by ascending flat index (top-k) / ascending candidate position (NMS sort).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def sigmoid_f32(x: np.ndarray) -> np.ndarray:
    x = x.astype(f32)
    return (f32(1.0) / (f32(1.0) + np.exp(-x, dtype=f32))).astype(f32)


def peak_map(heat: np.ndarray) -> np.ndarray:
    """heat (C,H,W) fp32 -> heat * (maxpool3x3(heat) == heat)."""
    C, H, W = heat.shape
    padded = np.full((C, H + 2, W + 2), -np.inf, f32)
    padded[:, 1:-1, 1:-1] = heat
    m = heat.copy()
    for dy in range(3):
        for dx in range(3):
            m = np.maximum(m, padded[:, dy:dy + H, dx:dx + W])
    return np.where(m == heat, heat, f32(0)).astype(f32)


def hm2box(heat, offset, wh, scale_factor=4, topk=10, conf_th=0.3, normalized=False):
    heat = np.asarray(heat, f32)
    offset = np.asarray(offset, f32)
    wh = np.asarray(wh, f32)
    C, H, W = heat.shape
    if topk > C * H * W:
        raise RuntimeError("selected index k out of range")
    flat = peak_map(heat).reshape(-1)
    order = np.argsort(-flat, kind="stable")[:topk]        # descending score, ties by ascending index
    scores = flat[order]
    cls = order // (H * W)
    rem = order % (H * W)
    ys, xs = rem // W, rem % W
    xo, yo = offset[0, ys, xs], offset[1, ys, xs]
    xs_, ys_ = wh[0, ys, xs], wh[1, ys, xs]
    if normalized:
        xo, yo = xo * f32(scale_factor), yo * f32(scale_factor)
        xs_, ys_ = xs_ * f32(W), ys_ * f32(H)
    xf, yf = xs.astype(f32), ys.astype(f32)
    sf = f32(scale_factor)
    half_w, half_h = xs_ / f32(2), ys_ / f32(2)
    boxes = np.stack([((xf + xo) - half_w) * sf, ((yf + yo) - half_h) * sf,
                      ((xf + xo) + half_w) * sf, ((yf + yo) + half_h) * sf], axis=1).astype(f32)
    keep = scores >= f32(conf_th)
    return boxes[keep], cls[keep].astype(np.int64), scores[keep]


def nms(boxes: np.ndarray, scores: np.ndarray, thr: float) -> np.ndarray:
    boxes = np.asarray(boxes, f32).reshape(-1, 4)
    scores = np.asarray(scores, f32)
    n = boxes.shape[0]
    order = np.argsort(-scores, kind="stable")
    x1, y1, x2, y2 = (boxes[:, i] for i in range(4))
    areas = ((x2 - x1) * (y2 - y1)).astype(f32)
    suppressed = np.zeros(n, bool)
    keep = []
    t = f32(thr)
    for a in range(n):
        i = order[a]
        if suppressed[i]:
            continue
        keep.append(i)
        for b in range(a + 1, n):
            j = order[b]
            if suppressed[j]:
                continue
            w = max(f32(0), min(x2[i], x2[j]) - max(x1[i], x1[j]))
            h = max(f32(0), min(y2[i], y2[j]) - max(y1[i], y1[j]))
            inter = f32(w * h)
            ovr = f32(inter / f32(f32(areas[i] + areas[j]) - inter))
            if ovr > t:
                suppressed[j] = True
    return np.asarray(keep, np.int64)


def predict(logits: np.ndarray, topk=100, scale_factor=4, conf_th=0.2, nms_th=0.2, normalized_coord=False,
            num_cls=None, do_nms=True):
    """logits: (B, S, C+4, H, W) raw network output -> three lists (len B): boxes (n,4), cls (n,), scores (n,)."""
    logits = np.asarray(logits, f32)
    B, S, O, H, W = logits.shape
    C = O - 4 if num_cls is None else num_cls
    out_b, out_c, out_s = [], [], []
    for b in range(B):
        bs, cs, ss = [], [], []
        for s in range(S):
            heat = sigmoid_f32(logits[b, s, :C])
            off, wh = logits[b, s, C:C + 2], logits[b, s, C + 2:C + 4]
            if normalized_coord:
                off, wh = sigmoid_f32(off), sigmoid_f32(wh)
            bx, cl, sc = hm2box(heat, off, wh, scale_factor, topk, conf_th, normalized_coord)
            bs.append(bx), cs.append(cl), ss.append(sc)
        bx, cl, sc = np.concatenate(bs), np.concatenate(cs), np.concatenate(ss)
        if do_nms:
            k = nms(bx, sc, nms_th)
            bx, cl, sc = bx[k], cl[k], sc[k]
        out_b.append(bx), out_c.append(cl), out_s.append(sc)
    return out_b, out_c, out_s


def synthetic_head(S=1, H=128, W=128, num_cls=2, seed=0, blobs=60):
    """Config-5 decode input: -6 logit floor + Gaussian blobs, offsets U(0,1), sizes U(4,24) cells."""
    rs = np.random.RandomState(seed)
    out = np.zeros((1, S, num_cls + 4, H, W), f32)
    yy, xx = np.mgrid[0:H, 0:W].astype(f32)
    for s in range(S):
        heat = (-6 + 0.5 * rs.randn(num_cls, H, W)).astype(f32)
        for _ in range(blobs):
            c = rs.randint(0, num_cls)
            cy, cx = rs.uniform(0, H - 1), rs.uniform(0, W - 1)
            amp, sig = rs.uniform(4, 10), rs.uniform(1, 4)
            blob = (-6 + (amp + 6) * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * sig * sig))).astype(f32)
            heat[c] = np.maximum(heat[c], blob)
        out[0, s, :num_cls] = heat
        out[0, s, num_cls:num_cls + 2] = rs.uniform(0, 1, (2, H, W)).astype(f32)
        out[0, s, num_cls + 2:] = rs.uniform(4, 24, (2, H, W)).astype(f32)
    return out
