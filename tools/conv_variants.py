"""Time the conv variants (1 generic, 2 halo) on the dominant shapes, plus the halo kernel's stage isolation
(debug 1 = no epilogue, 2 = no MMA issue)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from real_time_helmet_detection_b200 import ops, _lib
dev = torch.device("cuda:0")
for (B, H, W, C) in ((32, 128, 128, 128), (32, 64, 64, 128), (32, 32, 32, 128)):
    xs = [torch.randn(B, H, W, C, device=dev).bfloat16() for _ in range(2 if H == 256 else 3)]
    w = ops.pack_weight(torch.randn(C, C, 3, 3, device=dev) * 0.03)
    out = torch.empty(B, H, W, C, device=dev, dtype=torch.bfloat16)
    stats = torch.zeros(2, C, device=dev)
    flops = 2.0 * B * H * W * C * C * 9
    for v, st in ((1, stats), (2, stats), (2, None)):
        _lib.lib().hd_set_conv_variant(v)
        for i in range(3):
            ops.conv2d_igemm(xs[i % len(xs)], w, C, 3, stats=st, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        n = 20
        for i in range(n):
            ops.conv2d_igemm(xs[i % len(xs)], w, C, 3, stats=st, out=out)
        e1.record(); torch.cuda.synchronize()
        dt = e0.elapsed_time(e1) * 1e-3 / n
        print(f"B={B} {H}x{W} variant {v} stats={st is not None}: {dt*1e6:8.1f} us  {flops/dt/1e12:7.1f} TFLOP/s")
    _lib.lib().hd_set_conv_variant(2)
    for dbg in (1, 2):
        _lib.lib().hd_set_conv_debug(dbg)
        for st in (stats, None):
            for i in range(3):
                ops.conv2d_igemm(xs[i % len(xs)], w, C, 3, stats=st, out=out)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for i in range(20):
                ops.conv2d_igemm(xs[i % len(xs)], w, C, 3, stats=st, out=out)
            e1.record(); torch.cuda.synchronize()
            print(f"   halo dbg={dbg} stats={st is not None}: {e0.elapsed_time(e1)/20*1e3:8.1f} us")
    _lib.lib().hd_set_conv_debug(0)
    _lib.lib().hd_set_conv_variant(0)
