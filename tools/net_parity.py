"""Prints parity metrics of the CUDA network path vs the golden reference outputs and the oracle (run on a GPU box)."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from real_time_helmet_detection_b200.hourglass import StackedHourglass
from real_time_helmet_detection_b200.loss import LossCalculator
from oracle import hourglass_ref, loss_ref
from oracle.encode_ref import encode_boxes
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden import FIXED_BOXES

def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()

def gt_for(size, batch):
    outs = [[], [], [], []]
    for b in range(batch):
        boxes, labels = FIXED_BOXES[b % len(FIXED_BOXES)]
        sc = size / 256.0
        boxes = [[v * sc for v in bx] for bx in boxes]
        for lst, arr in zip(outs, encode_boxes(boxes, labels, (size, size))):
            lst.append(arr)
    return [torch.from_numpy(np.stack(o)) for o in outs]

dev = torch.device("cuda:0")
for S in (1, 2):
    for size in (128, 192):
        gold = np.load(os.path.join(ROOT, "tests", "golden", f"hourglass_s{S}_{size}.npz"))
        torch.manual_seed(777)
        net = StackedHourglass(S, 128, 6)
        sd0 = {k: v.clone() for k, v in net.state_dict().items()}
        x = torch.randn(2, 3, size, size, generator=torch.Generator().manual_seed(1))
        gts = gt_for(size, 2)
        # oracle fp32 + bf16-emulated, with autograd
        res = {}
        for mode in ("fp32", "bf16"):
            sd = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone()) for k, v in sd0.items()}
            out = hourglass_ref.stacked_hourglass_forward(sd, x, training=True, emulate_bf16=(mode == "bf16"))
            tot = 0
            for s in range(S):
                tot = tot + loss_ref.losses_from_logits(out[:, s], *gts)[3]
            tot.backward()
            res[mode] = (out.detach(), tot.item(), {k: v.grad for k, v in sd.items() if v.requires_grad})
        print(f"S={S} size={size}: oracle fp32 vs golden out_train rel {rel(res['fp32'][0], torch.from_numpy(gold['out_train'])):.2e}  loss {res['fp32'][1]:.6f} vs {float(gold['loss_total']):.6f}")
        net = net.to(dev).train()
        crit = LossCalculator(1.0, 1.0, 0.1, 2.0, 4.0)
        t0 = time.time()
        out = net(x.to(dev))
        tot = 0
        for s in range(S):
            tot = tot + crit.forward_logits(out[:, s], *[g.to(dev) for g in gts])
        tot.backward()
        torch.cuda.synchronize()
        o = out.detach().cpu()
        print(f"   cuda vs golden(ref fp32): rel {rel(o, torch.from_numpy(gold['out_train'])):.3e} maxabs {(o - torch.from_numpy(gold['out_train'])).abs().max():.3e} | vs oracle-bf16: rel {rel(o, res['bf16'][0]):.3e} | oracle-bf16 vs fp32 rel {rel(res['bf16'][0], res['fp32'][0]):.3e}")
        print(f"   loss cuda {tot.item():.5f} oracle-bf16 {res['bf16'][1]:.5f} ref {float(gold['loss_total']):.5f}")
        worst = []
        for n, p in net.named_parameters():
            g = p.grad.detach().cpu()
            gb, gf = res['bf16'][2][n], res['fp32'][2][n]
            worst.append((rel(g, gb), rel(g, gf), rel(gb, gf), gf.norm().item(), n))
        worst.sort(reverse=True)
        for w in worst[:6]:
            print("   grad rel(cuda,bf16)=%.3e rel(cuda,fp32)=%.3e rel(bf16,fp32)=%.3e |g|=%.3e %s" % w)
        med = sorted(w[0] for w in worst)[len(worst)//2]
        med2 = sorted(w[1] for w in worst)[len(worst)//2]
        print(f"   grad median rel vs bf16-oracle {med:.3e}, vs fp32 {med2:.3e}")
        # running stats
        sd1 = net.state_dict()
        names = list(gold["rstat_names"]); vals = gold["rstat_values"]; off = 0; mx = 0
        for k in names:
            v = sd1[k].cpu().float().numpy().ravel(); r = vals[off:off + v.size]; off += v.size
            mx = max(mx, np.abs(v - r).max() / (np.abs(r).max() + 1e-6))
        print(f"   running stats max rel-to-max err {mx:.3e}")
        net.eval()
        with torch.no_grad():
            oe = net(x.to(dev)).cpu()
        # eval uses running stats updated by OUR train step; the golden eval used the reference's updated stats
        print(f"   eval cuda vs golden out_eval rel {rel(oe, torch.from_numpy(gold['out_eval'])):.3e}")
