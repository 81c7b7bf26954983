"""Do the tensor-bound wgrad kernel and the HBM-bound BN-backward kernels overlap when issued on two streams?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from real_time_helmet_detection_b200 import ops
dev = torch.device("cuda:0")
B, H, W, C = 32, 128, 128, 128
x = torch.randn(B, H, W, C, device=dev).bfloat16()
dy = torch.randn(B, H, W, C, device=dev).bfloat16()
out = torch.randn(B, H, W, C, device=dev).bfloat16()
y = torch.randn(B, H, W, C, device=dev).bfloat16()
gamma = torch.ones(C, device=dev)
stats = torch.stack([y.float().sum((0, 1, 2)), (y.float() ** 2).sum((0, 1, 2))])
bnp = ops.bn_finalize(stats, B * H * W, gamma, torch.zeros(C, device=dev))
grad = torch.zeros(C, C, 3, 3, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def wg():
    ops.conv2d_wgrad(x, dy, C, 3, grad=grad)

def ew():
    ops.bn_bwd(dy, out, y, bnp, gamma)

def timeit(fa, fb, n=10):
    for _ in range(2):
        if fa: fa()
        if fb: fb()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    for _ in range(n):
        if fa:
            with torch.cuda.stream(s1): fa()
        if fb:
            with torch.cuda.stream(s2): fb()
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

print("wgrad alone      %.1f us" % timeit(wg, None))
print("bn_bwd alone     %.1f us" % timeit(None, ew))
print("both, 2 streams  %.1f us" % timeit(wg, ew))
w = ops.pack_weight(torch.randn(C, C, 3, 3, device=dev) * 0.03)
def cv():
    ops.conv2d_igemm(x, w, C, 3)
print("conv alone       %.1f us" % timeit(cv, None))
print("conv + bn_bwd    %.1f us" % timeit(cv, ew))
print("conv + wgrad     %.1f us" % timeit(cv, wg))
