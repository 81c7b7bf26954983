"""One warm train step of config 2 (or the given batch) for ncu: `python tools/profile_step.py [B] [S] [steps]`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from real_time_helmet_detection_b200.hourglass import StackedHourglass
from real_time_helmet_detection_b200.loss import LossCalculator
from real_time_helmet_detection_b200.synthetic import synthetic_targets
from real_time_helmet_detection_b200.train import train_step

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda:0")
torch.autograd.set_multithreading_enabled(False)    # backward on this thread, so that the NVTX range below covers it
torch.manual_seed(777)
net = StackedHourglass(S, 128, 6).to(dev).train()
crit = LossCalculator(1.0, 1.0, 0.1, 2.0, 4.0)
x = torch.randn(B, 3, 512, 512, device=dev)
gts = [torch.from_numpy(a).to(dev) for a in synthetic_targets(B, imsize=512)]
for i in range(steps):
    last = i == steps - 1
    if last:                      # `ncu --nvtx --nvtx-include "measured/"` profiles exactly one warm step
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push("measured")
    for p in net.parameters():
        p.grad = None
    train_step(net, crit, x, *gts)
    if last:
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
torch.cuda.synchronize()
print("done")
