"""Run-to-run variation of the flat gradient of one train step (same inputs, same weights)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from real_time_helmet_detection_b200.hourglass import StackedHourglass
from real_time_helmet_detection_b200.loss import LossCalculator
from real_time_helmet_detection_b200.synthetic import synthetic_targets
from real_time_helmet_detection_b200.train import train_step
dev = torch.device("cuda:0")
for (B, size) in ((2, 128), (8, 256), (16, 512)):
    torch.manual_seed(777)
    net = StackedHourglass(1, 128, 6).to(dev).train()
    crit = LossCalculator(1.0, 1.0, 0.1, 2.0, 4.0)
    x = torch.randn(B, 3, size, size, device=dev)
    gts = [torch.from_numpy(a).to(dev) for a in synthetic_targets(B, imsize=size)]
    runs, outs = [], []
    for _ in range(3):
        for p in net.parameters():
            p.grad = None
        with torch.no_grad():
            outs.append(net(x).clone())
        train_step(net, crit, x, *gts)
        runs.append(torch.cat([p.grad.flatten() for p in net.parameters()]).clone())
    r = lambda a, b: ((a - b).norm() / b.norm()).item()
    print(f"B={B} size={size} serial={os.environ.get('HD_SERIAL_WGRAD')}: logits run-to-run {r(outs[1], outs[0]):.3e} {r(outs[2], outs[0]):.3e} | grads {r(runs[1], runs[0]):.3e} {r(runs[2], runs[0]):.3e}")
