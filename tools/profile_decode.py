"""Config-5 decode (+ a small loss fwd/bwd) for ncu: python tools/profile_decode.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from real_time_helmet_detection_b200.synthetic import synthetic_head
from real_time_helmet_detection_b200.transform import _decode_call, _decode_buffers

dev = torch.device("cuda:0")
head = torch.from_numpy(synthetic_head(S=1)).to(dev)
B, S, O, H, W = head.shape
C, hw = O - 4, H * W
strides = ((S * O * hw, O * hw),) * 3
bufs = _decode_buffers(dev, B, S, C, H, W, 100)
for i in range(12):
    if i == 10:
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push("measured")
    _decode_call(head, head[:, :, C:], head[:, :, C + 2:], strides, B, S, C, H, W, 100, 4, 0.2, 0.2, False, True, True, bufs=bufs)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("boxes", int(bufs[3][0]))
