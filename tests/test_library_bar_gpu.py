"""The "library bar" of SURVEY.md §8(d): the reference's architecture restated with stock torch.nn modules
(nn.Conv2d -> cuDNN, nn.BatchNorm2d -> cuDNN/ATen fused BN, MaxPool2d, Upsample, autograd - exactly the library calls
hourglass.py:94-237 makes) run on the B200 under bf16 autocast, timed beside our hand-written path on the same
workload; the oracle port (hand-expanded BN, more elementwise launches) is timed too. It is a measurement, not a parity test: it runs only when
HD_LIBRARY_BAR=1 (it autotunes cuDNN and allocates ~20 GB) and writes gpurun_out/library_bar.json.

    HD_LIBRARY_BAR=1 python -m pytest tests/test_library_bar_gpu.py -m gpu -q -s
"""
import json
import os

import pytest

pytestmark = pytest.mark.gpu


def _library_net(torch, S=1, ch=128, out_ch=6):
    nn = torch.nn

    class Conv(nn.Module):                       # hourglass.py:94-108
        def __init__(self, i, o, k, stride=1, bias=False, bn=True, relu=True):
            super().__init__()
            self.c = nn.Conv2d(i, o, k, stride, (k - 1) // 2, bias=bias)
            self.b = nn.BatchNorm2d(o) if bn else nn.Identity()
            self.a = nn.ReLU() if relu else nn.Identity()

        def forward(self, x):
            return self.a(self.b(self.c(x)))

    class Res(nn.Module):                        # hourglass.py:111-127
        def __init__(self, i, o):
            super().__init__()
            self.c1, self.c2 = Conv(i, o, 3), Conv(o, o, 3, relu=False)
            self.skip = Conv(i, o, 1, relu=False) if i != o else nn.Identity()

        def forward(self, x):
            return torch.relu(self.c2(self.c1(x)) + self.skip(x))

    class HG(nn.Module):                         # hourglass.py:130-156
        def __init__(self, depth):
            super().__init__()
            self.up1, self.low1, self.low3 = Res(ch, ch), Res(ch, ch), Res(ch, ch)
            self.low2 = HG(depth - 1) if depth > 1 else Res(ch, ch)
            self.pool, self.up = nn.MaxPool2d(2, 2), nn.Upsample(scale_factor=2, mode="nearest")

        def forward(self, x):
            return self.up1(x) + self.up(self.low3(self.low2(self.low1(self.pool(x)))))

    class Net(nn.Module):                        # hourglass.py:159-237
        def __init__(self):
            super().__init__()
            self.pre = nn.Sequential(Conv(3, 64, 7, 2, bias=True), Res(64, ch), nn.MaxPool2d(2, 2), Res(ch, ch), Res(ch, ch))
            self.hg = nn.ModuleList(HG(4) for _ in range(S))
            self.neck = nn.ModuleList(nn.Sequential(Conv(ch, ch, 1, bias=True), Res(ch, ch)) for _ in range(S))
            self.head = nn.ModuleList(Conv(ch, out_ch, 1, bias=True, bn=False, relu=False) for _ in range(S))
            self.mf = nn.ModuleList(Conv(ch, ch, 1, bias=True, bn=False, relu=False) for _ in range(S - 1))
            self.mp = nn.ModuleList(Conv(out_ch, ch, 1, bias=True, bn=False, relu=False) for _ in range(S - 1))

        def forward(self, x):
            x = self.pre(x)
            outs = []
            for i in range(S):
                f = self.neck[i](self.hg[i](x))
                p = self.head[i](f)
                outs.append(p)
                if i < S - 1:
                    x = x + self.mf[i](f) + self.mp[i](p)
            return torch.stack(outs, 1)

    return Net()


@pytest.mark.skipif(os.environ.get("HD_LIBRARY_BAR") != "1", reason="measurement only: set HD_LIBRARY_BAR=1")
def test_library_bar():
    import torch
    from oracle import hourglass_ref, loss_ref
    from real_time_helmet_detection_b200.hourglass import StackedHourglass
    from real_time_helmet_detection_b200.loss import LossCalculator
    from real_time_helmet_detection_b200.synthetic import synthetic_targets
    from real_time_helmet_detection_b200.train import train_step

    dev = torch.device("cuda:0")
    B, S, size = int(os.environ.get("HD_LIBRARY_BAR_BATCH", "32")), 1, 512
    steps, warmup = 10, 4
    torch.manual_seed(777)
    net = StackedHourglass(S, 128, 6).to(dev).train()
    sd = {k: (v.detach().clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k
              else v.detach().clone()) for k, v in net.state_dict().items()}
    x = torch.randn(B, 3, size, size, device=dev)
    gts = [torch.from_numpy(a).to(dev) for a in synthetic_targets(B, imsize=size)]
    torch.backends.cudnn.benchmark = True

    def lib_step(inp):
        for v in sd.values():
            if v.requires_grad:
                v.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = hourglass_ref.stacked_hourglass_forward(sd, inp, training=True)
        tot = sum(loss_ref.losses_from_logits(out[:, s].float(), *gts)[3] for s in range(S))
        tot.backward()

    def timed(fn):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    libnet = _library_net(torch, S).to(dev).train()
    assert sum(p.numel() for p in libnet.parameters()) == sum(p.numel() for p in net.parameters())

    def nn_step(inp):
        for p in libnet.parameters():
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = libnet(inp)
        tot = sum(loss_ref.losses_from_logits(out[:, s].float(), *gts)[3] for s in range(S))
        tot.backward()

    res = {"workload": f"{S}-stack hourglass, {size}x{size}, batch {B}, train fwd + loss + bwd, bf16 autocast",
           "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}
    xcl = x.contiguous(memory_format=torch.channels_last)
    for name, fn in (("nn_nchw", lambda: nn_step(x)), ("nn_channels_last", None), ("oracle_port_nchw", lambda: lib_step(x))):
        try:
            if fn is None:
                libnet.to(memory_format=torch.channels_last)
                fn = lambda: nn_step(xcl)  # noqa: E731
            ms = timed(fn)
            res[f"library_{name}_ms"] = ms
            res[f"library_{name}_img_s"] = B / ms * 1e3
        except RuntimeError as e:  # e.g. out of memory on a shared box
            res[f"library_{name}_error"] = str(e)[:200]
        torch.cuda.empty_cache()
    crit = LossCalculator(1.0, 1.0, 0.1, 2.0, 4.0)

    def ours():
        for p in net.parameters():
            p.grad = None
        train_step(net, crit, x, *gts)

    ms = timed(ours)
    res["ours_ms"], res["ours_img_s"] = ms, B / ms * 1e3
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/library_bar.json", "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))
    assert res["ours_img_s"] > 0
