"""HD_TRACE=1 python tools/trace_step.py [B] [S] > trace.txt : per-launch (stream, start, end) rows of one warm train step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from real_time_helmet_detection_b200 import _lib
from real_time_helmet_detection_b200.hourglass import StackedHourglass
from real_time_helmet_detection_b200.loss import LossCalculator
from real_time_helmet_detection_b200.synthetic import synthetic_targets
from real_time_helmet_detection_b200.train import train_step
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda:0")
torch.manual_seed(777)
net = StackedHourglass(S, 128, 6).to(dev).train()
crit = LossCalculator(1.0, 1.0, 0.1, 2.0, 4.0)
x = torch.randn(B, 3, 512, 512, device=dev)
gts = [torch.from_numpy(a).to(dev) for a in synthetic_targets(B, imsize=512)]
for i in range(4):
    for p in net.parameters():
        p.grad = None
    if i == 3:
        torch.cuda.synchronize()
        _lib.lib().hd_trace_dump()          # drop the warm-up rows
        print("==== traced step", file=sys.stderr)
    train_step(net, crit, x, *gts)
torch.cuda.synchronize()
_lib.lib().hd_trace_dump()
