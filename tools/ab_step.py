#!/usr/bin/env python
"""Quick A/B timing of the config-2 train step (device-resident inputs): python tools/ab_step.py [steps] [B] [S].
Knobs are environment variables read by the library (HD_SM_RESERVE, HD_NO_DUAL_DGRAD, HD_TRAIN_PDL, ...), so one gpurun call
can compare several settings on the same box:  for r in 0 8 16 24; do HD_SM_RESERVE=$r python tools/ab_step.py; done"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from real_time_helmet_detection_b200.hourglass import StackedHourglass  # noqa: E402
from real_time_helmet_detection_b200.loss import LossCalculator  # noqa: E402
from real_time_helmet_detection_b200.synthetic import synthetic_targets  # noqa: E402
from real_time_helmet_detection_b200.train import train_step  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
S = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda:0")
torch.manual_seed(777)
net = StackedHourglass(S, 128, 6).to(dev).train()
crit = LossCalculator(1.0, 1.0, 0.1, 2.0, 4.0)
x = torch.randn(B, 3, 512, 512, device=dev)
gts = [torch.from_numpy(a).to(dev) for a in synthetic_targets(B, imsize=512)]


def step():
    for p in net.parameters():
        p.grad = None
    return train_step(net, crit, x, *gts)


if os.environ.get("AB_GRAPH"):                       # the same step replayed from a CUDA graph
    from real_time_helmet_detection_b200.train import GraphedTrainStep
    graphed = GraphedTrainStep(net, crit, x, *gts)

    def step():                                       # noqa: F811
        return graphed(x, *gts, log=False)


for _ in range(6):
    step()
torch.cuda.synchronize()
best = []
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    best.append(e0.elapsed_time(e1) / steps)
knobs = {k: v for k, v in os.environ.items() if k.startswith(("HD_", "AB_"))}
print(f"ab_step B={B} S={S} {knobs}: ms/step {min(best):.3f} (reps {', '.join('%.3f' % b for b in best)})  "
      f"img/s {B / min(best) * 1e3:.1f}  loss {float(loss):.4f}", flush=True)
