"""Where does the end-to-end step spend host time? `python tools/e2e_probe.py [S] [B]`"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from real_time_helmet_detection_b200.hourglass import StackedHourglass
from real_time_helmet_detection_b200.loss import LossCalculator
from real_time_helmet_detection_b200.synthetic import synthetic_targets
from real_time_helmet_detection_b200.train import train_step, DevicePrefetcher
S = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = torch.device("cuda:0")
torch.manual_seed(777)
net = StackedHourglass(S, 128, 6).to(dev).train()
crit = LossCalculator(1.0, 1.0, 0.1, 2.0, 4.0)
img = torch.randn(B, 3, 512, 512)
gts = [torch.from_numpy(a) for a in synthetic_targets(B, imsize=512)]
img_d, gts_d = img.to(dev), [g.to(dev) for g in gts]
img_p, gts_p = img.pin_memory(), [g.pin_memory() for g in gts]

def gen():
    while True:
        yield (img_p, *gts_p)
loader = DevicePrefetcher(gen(), dev)
lh = [torch.zeros(1).pin_memory() for _ in range(2)]
le = [torch.cuda.Event() for _ in range(2)]
st = {"i": 0}

def zero():
    for p in net.parameters():
        p.grad = None

def dev_step():
    zero(); train_step(net, crit, img_d, *gts_d)

def e2e_full():
    i = st["i"]; zero()
    loss = train_step(net, crit, *next(loader))
    lh[i & 1].copy_(loss.reshape(1), non_blocking=True); le[i & 1].record()
    if i > 0:
        le[(i - 1) & 1].synchronize(); float(lh[(i - 1) & 1])
    st["i"] = i + 1

def e2e_noread():
    zero(); train_step(net, crit, *next(loader))

def e2e_direct():
    zero(); train_step(net, crit, img_p, *gts_p)

def e2e_imgonly():
    zero(); train_step(net, crit, img_p.to(dev, non_blocking=True), *gts_d)

for name, fn in (("device", dev_step), ("e2e full", e2e_full), ("e2e no loss readback", e2e_noread),
                 ("direct .to on the compute stream", e2e_direct), ("image only H2D", e2e_imgonly), ("device", dev_step)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(10):
        fn()
    t_host = time.perf_counter() - t0
    e1.record(); torch.cuda.synchronize()
    print(f"{name:36s}: device {e0.elapsed_time(e1)/10:7.2f} ms/step   host enqueue {t_host*100:7.2f} ms/step")
