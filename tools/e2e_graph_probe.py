import os, sys, time
sys.path.insert(0, "/root/repo")
import torch
from real_time_helmet_detection_b200.hourglass import StackedHourglass
from real_time_helmet_detection_b200.loss import LossCalculator
from real_time_helmet_detection_b200.synthetic import synthetic_targets
from real_time_helmet_detection_b200.train import GraphedTrainStep
dev = torch.device("cuda:0")
torch.manual_seed(777)
net = StackedHourglass(1, 128, 6).to(dev).train()
crit = LossCalculator(1.0, 1.0, 0.1, 2.0, 4.0)
B = 32
x = torch.randn(B, 3, 512, 512)
gts = [torch.from_numpy(a) for a in synthetic_targets(B, imsize=512)]
xp, gp = x.pin_memory(), [g.pin_memory() for g in gts]
xd, gd = x.to(dev), [g.to(dev) for g in gts]
g = GraphedTrainStep(net, crit, xd, *gd, buffers=2)
def timeit(fn, n=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("replay graph0 only        %.3f" % timeit(lambda: g._replay(0, False)))
st = {"k": 0}
def alt():
    g._replay(st["k"], False); st["k"] ^= 1
print("alternate graphs, no copy %.3f" % timeit(alt))
g.stage((xd, *gd))
def dev_stage():
    g.run(log=False); g.stage((xd, *gd))
print("stage from device         %.3f" % timeit(dev_stage))
def host_stage():
    g.run(log=False); g.stage((xp, *gp))
print("stage from pinned host    %.3f" % timeit(host_stage))
# bulk H2D alone
def h2d():
    for d_, s_ in zip(g.sets[0], (xp, *gp)): d_.copy_(s_, non_blocking=True)
print("H2D alone (115 MB)        %.3f" % timeit(h2d))
