"""Device-side tail of the reference's collate_fn (data.py:98-123): ground-truth encoding and image normalisation.

The reference builds, per image and on the host, the CenterNet targets with `box2hm` (transform.py:4-45, NumPy) and the
normalised float image with `TF.to_tensor` + `Normalize` (data.py:118, utils.py:55-68), then uploads five float tensors
(and re-uploads the four GT tensors once per stack, train.py:115-118). Here the host only stages what the augmentation
produced - uint8 HWC images and a padded box list - in pinned memory; ONE small H2D copy later two kernels
(`hd_normalize_u8`, `hd_encode_targets`, csrc/encode.cu) produce the same five tensors on the device
(SURVEY.md 8(f)-2). 4x fewer bytes cross PCIe and no NumPy loop runs per image.

There is no CPU fallback: the host-side `transform.box2hm` remains for callers that want NumPy arrays.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream

BOX_CHUNK = 128       # kEncMaxBoxes in csrc/encode.cu: boxes are processed in chunks of this size, any count works

_NORMALIZERS = {      # utils.py:55-62 `get_normalizer`
    "imagenet": ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225)),
    "scratch": ((0.5, 0.5, 0.5), (0.5, 0.5, 0.5)),
}


def normalizer_constants(pretrained: str):
    key = pretrained.lower()
    if key not in _NORMALIZERS:
        raise NotImplementedError("Not expected dataset pretrained parameter: %s" % pretrained)   # utils.py:63
    return _NORMALIZERS[key]


def normalize_images(images_u8: torch.Tensor, pretrained: str = "imagenet", out: Optional[torch.Tensor] = None):
    """uint8 (B,H,W,3) CUDA tensor -> normalised fp32 (B,3,H,W) (`TF.to_tensor` + `Normalize`, data.py:118)."""
    _lib.require_cuda(images_u8, "images")
    if images_u8.dtype != torch.uint8 or images_u8.dim() != 4 or images_u8.shape[-1] != 3:
        raise RuntimeError(f"expected a uint8 (B,H,W,3) tensor, got {images_u8.dtype} {tuple(images_u8.shape)}")
    images_u8 = images_u8.contiguous()
    B, H, W, _ = images_u8.shape
    mean, std = normalizer_constants(pretrained)
    if out is None:
        out = torch.empty(B, 3, H, W, device=images_u8.device, dtype=torch.float32)
    m = (ctypes.c_float * 3)(*mean)
    s = (ctypes.c_float * 3)(*std)
    with torch.cuda.device(images_u8.device):
        check(_lib.lib().hd_normalize_u8(ptr(images_u8), ptr(out), B, H, W, m, s, stream(images_u8.device)),
              "hd_normalize_u8")
    return out


def encode_targets(boxes: torch.Tensor, labels: torch.Tensor, imsize, scale_factor: int = 4, num_cls: int = 2,
                   normalized: bool = False, return_errors: bool = False):
    """Batched `box2hm` (transform.py:4-45) on the device.

    boxes (B,Nmax,4) fp32 CUDA (xmin,ymin,xmax,ymax in input pixels), labels (B,Nmax) int32 CUDA with -1 marking
    empty slots; imsize = (width, height) of the network input as in the reference. Returns
    (heat (B,C,h,w), offset (B,2,h,w), size (B,2,h,w), mask (B,1,h,w)) fp32; with return_errors also a device int32
    tensor counting boxes skipped because their centre fell outside the map (an IndexError in the reference)."""
    _lib.require_cuda(boxes, "boxes")
    _lib.require_cuda(labels, "labels")
    if boxes.dim() != 3 or boxes.shape[-1] != 4 or labels.shape != boxes.shape[:2]:
        raise RuntimeError(f"expected boxes (B,Nmax,4) and labels (B,Nmax), got {tuple(boxes.shape)} / {tuple(labels.shape)}")
    if boxes.dtype != torch.float32 or labels.dtype != torch.int32:
        raise RuntimeError("boxes must be float32 and labels int32")
    boxes, labels = boxes.contiguous(), labels.contiguous()
    B, nmax = labels.shape
    w, h = int(imsize[0]) // scale_factor, int(imsize[1]) // scale_factor      # transform.py:5
    dev = boxes.device
    heat = torch.empty(B, num_cls, h, w, device=dev, dtype=torch.float32)
    off = torch.empty(B, 2, h, w, device=dev, dtype=torch.float32)
    size = torch.empty(B, 2, h, w, device=dev, dtype=torch.float32)
    mask = torch.empty(B, 1, h, w, device=dev, dtype=torch.float32)
    err = torch.zeros(1, device=dev, dtype=torch.int32) if return_errors else None
    with torch.cuda.device(dev):
        check(_lib.lib().hd_encode_targets(ptr(boxes), ptr(labels), B, nmax, h, w, num_cls, int(scale_factor),
                                           int(bool(normalized)), ptr(heat), ptr(off), ptr(size), ptr(mask), ptr(err),
                                           stream(dev)), "hd_encode_targets")
    return (heat, off, size, mask, err) if return_errors else (heat, off, size, mask)


def pad_boxes(batch_bbs_lst: Sequence[Sequence], batch_id_lst: Sequence[Sequence], nmax: Optional[int] = None):
    """The nested lists data.py:99-106 builds -> padded (B,Nmax,4) float32 / (B,Nmax) int32 NumPy arrays (-1 = empty).
    `None` boxes are skipped like transform.py:15-16."""
    B = len(batch_bbs_lst)
    longest = max([len(b) for b in batch_bbs_lst] + [1])
    nmax = longest if nmax is None else max(nmax, longest)     # like box2hm, no cap on the boxes per image
    boxes = np.zeros((B, nmax, 4), np.float32)
    labels = np.full((B, nmax), -1, np.int32)
    for b, (bbs, ids) in enumerate(zip(batch_bbs_lst, batch_id_lst)):
        for j, (bb, lab) in enumerate(zip(bbs, ids)):
            if bb is None:
                continue
            boxes[b, j] = np.asarray(bb, np.float32).reshape(4)
            labels[b, j] = int(lab)
    return boxes, labels


class DeviceCollate:
    """`collate_fn` tail on the device: (uint8 HWC images, box lists) -> the five training tensors on `device`.

    Two pinned staging slots alternate, so batch i+1 can be staged while batch i's copy is still in flight. `max_boxes`
    is only the initial size of the pinned box slots: an image with more boxes grows them (the reference's
    collate_fn / box2hm accept any number of boxes, and so does the encoder kernel)."""

    def __init__(self, device, num_cls: int = 2, normalized_coord: bool = False, pretrained: str = "imagenet",
                 scale_factor: int = 4, max_boxes: int = 32):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DeviceCollate needs a CUDA device: this path has no CPU implementation")
        normalizer_constants(pretrained)
        self.num_cls, self.normalized, self.pretrained = num_cls, normalized_coord, pretrained
        self.scale_factor, self.max_boxes = scale_factor, max_boxes
        self._slots = [None, None]
        self._turn = 0

    def _slot(self, B, H, W, nbox):
        if nbox > self.max_boxes:            # a crowded image: grow the box slots (rounded up, so this happens rarely)
            self.max_boxes = (nbox + 31) // 32 * 32
        s = self._slots[self._turn]
        if s is None or s["img"].shape != (B, H, W, 3) or s["box"].shape[1] != self.max_boxes:
            s = {"img": torch.empty(B, H, W, 3, dtype=torch.uint8).pin_memory(),
                 "box": torch.empty(B, self.max_boxes, 4, dtype=torch.float32).pin_memory(),
                 "lab": torch.empty(B, self.max_boxes, dtype=torch.int32).pin_memory(),
                 "done": torch.cuda.Event()}
            self._slots[self._turn] = s
        else:
            s["done"].synchronize()          # the copy that last used this slot has finished
        self._turn ^= 1
        return s

    def __call__(self, img_np_lst, batch_bbs_lst, batch_id_lst):
        B = len(img_np_lst)
        H, W = img_np_lst[0].shape[:2]
        s = self._slot(B, H, W, max([len(b) for b in batch_bbs_lst] + [1]))
        img_host = s["img"].numpy()
        for b, im in enumerate(img_np_lst):
            img_host[b] = im
        boxes, labels = pad_boxes(batch_bbs_lst, batch_id_lst, self.max_boxes)
        s["box"].numpy()[...] = boxes
        s["lab"].numpy()[...] = labels
        img_d = s["img"].to(self.device, non_blocking=True)
        box_d = s["box"].to(self.device, non_blocking=True)
        lab_d = s["lab"].to(self.device, non_blocking=True)
        s["done"].record()
        image = normalize_images(img_d, self.pretrained)
        heat, off, size, mask = encode_targets(box_d, lab_d, (W, H), self.scale_factor, self.num_cls, self.normalized)
        return image, heat, off, size, mask

    @property
    def h2d_bytes(self):
        s = self._slots[0] or self._slots[1]
        return 0 if s is None else sum(s[k].numel() * s[k].element_size() for k in ("img", "box", "lab"))
