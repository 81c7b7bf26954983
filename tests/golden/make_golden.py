"""Generates the golden fixtures in this directory by running the UNMODIFIED reference modules.

Run inside the build container (where the reference is mounted read-only at /root/reference):

    python tests/golden/make_golden.py

It imports the reference's own hourglass.py / loss.py / transform.py / evaluate.py (imgaug and torchsummary,
which are not installed and only used by the dataloader / summary printing, are stubbed in sys.modules), feeds
them seeded inputs and stores inputs-by-seed + outputs as small .npz files. The fixtures are what pins the
oracle (oracle/*.py) and, through it, the CUDA path; /root/reference is never needed at test time.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("HD_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def import_reference():
    ia = _stub("imgaug")
    ia.augmenters = _stub("imgaug.augmenters")
    ia.augmentables = _stub("imgaug.augmentables")
    ia.augmentables.bbs = _stub("imgaug.augmentables.bbs", BoundingBox=object, BoundingBoxesOnImage=object)
    _stub("torchsummary", summary=lambda *a, **k: None)
    sys.path.insert(0, REF)
    import hourglass as ref_hg          # noqa
    import loss as ref_loss              # noqa
    import transform as ref_tf           # noqa
    import evaluate as ref_ev            # noqa
    sys.path.remove(REF)
    return ref_hg, ref_loss, ref_tf, ref_ev


FIXED_BOXES = [
    ([[10, 20, 100, 200], [30, 40, 60, 90]], [1, 0]),
    ([[5, 5, 50, 40]], [0]),
]


def gt_for(ref_tf, size, batch):
    outs = [[], [], [], []]
    for b in range(batch):
        boxes, labels = FIXED_BOXES[b % len(FIXED_BOXES)]
        sc = size / 256.0
        boxes = [[v * sc for v in bx] for bx in boxes]
        for lst, arr in zip(outs, ref_tf.box2hm(boxes, labels, (size, size), scale_factor=4, num_cls=2)):
            lst.append(arr)
    return [torch.from_numpy(np.stack(o)) for o in outs]


def golden_hourglass(ref_hg, ref_loss, ref_tf):
    for S in (1, 2):
        for size in (128, 192):
            torch.manual_seed(777)
            net = ref_hg.StackedHourglass(num_stack=S, in_ch=128, out_ch=6)
            x = torch.randn(2, 3, size, size, generator=torch.Generator().manual_seed(1))
            ghm, goff, gsize, gmask = gt_for(ref_tf, size, 2)
            crit = ref_loss.LossCalculator(1.0, 1.0, 0.1, 2.0, 4.0)
            net.train()
            out_train = net(x)
            total = 0
            comps = []
            for s in range(S):                       # train.py:105-120 with the out-of-place squeeze
                o = out_train[:, s]
                phm, poff, psz = o.split([2, 2, 2], dim=1)
                total = total + crit(torch.sigmoid(phm), poff, psz, ghm, goff, gsize, gmask)
                comps.append([crit.log[k][-1] for k in ("hm", "offset", "size", "total")])
            total.backward()
            names = [n for n, _ in net.named_parameters()]
            grads = {n: p.grad.detach().clone() for n, p in net.named_parameters()}
            sd = net.state_dict()
            rstats = {k: v.numpy().copy() for k, v in sd.items() if "running_" in k}
            net.eval()
            with torch.no_grad():
                out_eval = net(x)
            keep_full = ["head_lst.0.layer.convolution.weight", "head_lst.0.layer.convolution.bias",
                         "pre_layer.layers.0.bn.weight", "pre_layer.layers.0.bn.bias",
                         "pre_layer.layers.0.convolution.weight",
                         "hourglass_lst.0.low2.low2.low2.low2.conv1.bn.weight",
                         "neck_lst.0.layers.1.convolution.bias"]
            np.savez_compressed(
                os.path.join(HERE, f"hourglass_s{S}_{size}.npz"),
                out_train=out_train.detach().numpy(), out_eval=out_eval.numpy(),
                loss_components=np.asarray(comps, np.float64), loss_total=float(total),
                param_names=np.asarray(names), grad_norms=np.asarray([grads[n].norm().item() for n in names]),
                param_sums=np.asarray([sd[n].double().sum().item() for n in names]),
                rstat_names=np.asarray(sorted(rstats)),
                rstat_values=np.concatenate([rstats[k].ravel() for k in sorted(rstats)]),
                **{"grad::" + n: grads[n].numpy() for n in keep_full},
                grad_conv_slice=grads["hourglass_lst.0.up1.conv1.convolution.weight"][:8, :8].numpy(),
            )
            print("hourglass", S, size, "loss", float(total), "params", sum(p.numel() for p in net.parameters()))


def golden_loss(ref_loss, ref_tf):
    g = torch.Generator().manual_seed(42)
    B, h = 3, 32
    logits = torch.randn(B, 6, h, h, generator=g) * 2.0
    cases = {}
    gts = {"boxes": gt_for(ref_tf, 128, B)}
    gts["nopos"] = [torch.zeros(B, 2, h, h), torch.zeros(B, 2, h, h), torch.zeros(B, 2, h, h), torch.zeros(B, 1, h, h)]
    for name, (ghm, goff, gsize, gmask) in gts.items():
        for norm in (False, True):
            lg = logits.clone().requires_grad_(True)
            crit = ref_loss.LossCalculator(1.0, 1.0, 0.1, 2.0, 4.0)
            phm, poff, psz = lg.split([2, 2, 2], dim=1)
            phm = torch.sigmoid(phm)
            if norm:
                poff, psz = torch.sigmoid(poff), torch.sigmoid(psz)
            total = crit(phm, poff, psz, ghm, goff, gsize, gmask)
            total.backward()
            key = f"{name}_{'norm' if norm else 'lin'}"
            cases[key + "_values"] = np.asarray([crit.log[k][-1] for k in ("hm", "offset", "size", "total")])
            cases[key + "_dlogits"] = lg.grad.numpy()
    crit = ref_loss.LossCalculator(1.0, 1.0, 0.1, 2.0, 4.0)
    crit.log = {k: [0.5 * i for i in range(150)] for k in ("hm", "offset", "size", "total")}
    np.savez_compressed(os.path.join(HERE, "loss.npz"), logits=logits.numpy(),
                        get_log=np.asarray(crit.get_log()), **cases)
    print("loss", {k: v for k, v in cases.items() if k.endswith("values")})


class _FixedNet(torch.nn.Module):
    def __init__(self, out):
        super().__init__()
        self.out = out

    def forward(self, x):
        return self.out.clone()


def golden_decode(ref_tf, ref_ev):
    from oracle.decode_ref import synthetic_head
    res = {}
    for S in (1, 2):
        head = torch.from_numpy(synthetic_head(S=S, seed=S - 1))
        for norm in (False, True):
            pred = ref_ev.Prediction(_FixedNet(head), topk=100, scale_factor=4, conf_th=0.2, nms="nms", nms_th=0.2,
                                     normalized_coord=norm)
            with torch.no_grad():
                b, c, s = pred(torch.zeros(1, 3, 512, 512))
            tag = f"pred_s{S}_{'norm' if norm else 'lin'}"
            res[tag + "_boxes"], res[tag + "_cls"], res[tag + "_scores"] = b[0].numpy(), c[0].numpy(), s[0].numpy()
            print(tag, b[0].shape)
        # decode without NMS (hm2box on the activated maps of stack 0)
        hm = torch.sigmoid(head[0, 0, :2])
        b, c, s = ref_tf.hm2box(hm, head[0, 0, 2:4], head[0, 0, 4:6], scale_factor=4, topk=100, conf_th=0.2)
        res[f"hm2box_s{S}_boxes"], res[f"hm2box_s{S}_cls"], res[f"hm2box_s{S}_scores"] = b.numpy(), c.numpy(), s.numpy()
    # plateau: two equal neighbouring maxima are both peaks (transform.py:78)
    hm = torch.zeros(2, 8, 8)
    hm[0, 3, 3] = hm[0, 3, 4] = 0.9
    hm[1, 6, 1] = 0.7
    off, wh = torch.full((2, 8, 8), 0.25), torch.full((2, 8, 8), 2.0)
    b, c, s = ref_tf.hm2box(hm, off, wh, scale_factor=4, topk=5, conf_th=0.3)
    res["plateau_boxes"], res["plateau_cls"], res["plateau_scores"] = b.numpy(), c.numpy(), s.numpy()
    # class-agnostic suppression: overlapping boxes of different classes (evaluate.py:174)
    boxes = torch.tensor([[0., 0., 10., 10.], [1., 1., 11., 11.], [20., 20., 30., 30.]])
    pred = ref_ev.Prediction(None, 10, 4, 0.0, "nms", 0.2)
    kb, kc, ks = pred.nonmaximum_supression(boxes, torch.tensor([0, 1, 1]), torch.tensor([0.9, 0.8, 0.7]))
    res["agnostic_boxes"], res["agnostic_cls"], res["agnostic_scores"] = kb.numpy(), kc.numpy(), ks.numpy()
    np.savez_compressed(os.path.join(HERE, "decode.npz"), **res)


def golden_encode(ref_tf):
    # known-answer self-test of the reference (transform.py:112-131)
    hm, off, wh, mask = ref_tf.box2hm([[10, 20, 100, 200]], [1], (512, 512), normalized=True)
    b, c, s = ref_tf.hm2box(torch.from_numpy(hm), torch.from_numpy(off), torch.from_numpy(wh), normalized=True)
    res = dict(kat_heat=hm[:, 27, 13], kat_off=off[:, 27, 13], kat_wh=wh[:, 27, 13], kat_box=b.numpy(),
               kat_cls=c.numpy(), kat_scores=s.numpy())
    from oracle.encode_ref import synthetic_targets
    import numpy.random as npr
    # the same seeded boxes synthetic_targets draws, pushed through the reference encoder
    outs = [[], [], [], []]
    for bidx in range(4):
        rs = npr.RandomState(bidx)
        nb = rs.randint(1, 6)
        boxes, labels = [], []
        for _ in range(nb):
            x0, y0 = rs.uniform(0, 0.7 * 512, 2)
            bw, bh = rs.uniform(0.05, 0.3, 2) * 512
            boxes.append([x0, y0, min(x0 + bw, 511), min(y0 + bh, 511)])
            labels.append(int(rs.randint(0, 2)))
        for lst, arr in zip(outs, ref_tf.box2hm(boxes, labels, (512, 512))):
            lst.append(arr)
    ghm, goff, gsz, gmask = (np.stack(o) for o in outs)
    mine = synthetic_targets(4)
    assert all(np.array_equal(a, b) for a, b in zip(mine, (ghm, goff, gsz, gmask))), "encode_ref != reference"
    res.update(enc_heat_sum=ghm.sum(axis=(2, 3)), enc_mask_idx=np.argwhere(gmask > 0),
               enc_off=goff[gmask.repeat(2, 1) > 0], enc_size=gsz[gmask.repeat(2, 1) > 0])
    np.savez_compressed(os.path.join(HERE, "encode.npz"), **res)
    print("encode KAT", res["kat_heat"], res["kat_off"], res["kat_wh"], res["kat_box"])


def encode_f32_cases():
    """Box lists with float32-representable coordinates (what the device encoder takes) covering: ordinary boxes,
    boxes clipped by every border, two boxes sharing a centre cell (the later one owns offset/size), concentric
    same-class boxes (running maximum), an empty image, a `None` slot, both classes. 256x256 input -> 64x64 maps."""
    rs = np.random.RandomState(20260921)
    imgs = []
    for b in range(5):
        boxes, labels = [], []
        for _ in range(rs.randint(1, 7)):
            x0, y0 = rs.uniform(0, 0.8 * 256, 2)
            bw, bh = rs.uniform(0.04, 0.5, 2) * 256
            boxes.append([float(np.float32(v)) for v in (x0, y0, min(x0 + bw, 255.0), min(y0 + bh, 255.0))])
            labels.append(int(rs.randint(0, 2)))
        imgs.append((boxes, labels))
    imgs.append(([[0.0, 0.0, 37.5, 21.25], [200.5, 0.0, 255.0, 90.0], [0.0, 180.0, 60.0, 255.75], [190.0, 170.0, 255.0, 255.0]],
                 [0, 1, 0, 1]))                                                     # clipped by each border / corner
    imgs.append(([[100.0, 100.0, 140.0, 140.0], [110.0, 110.0, 130.5, 131.0], [90.0, 80.0, 150.0, 160.0]], [1, 1, 1]))
    imgs.append(([[64.0, 64.0, 96.0, 96.0], None, [66.0, 62.0, 94.5, 98.0]], [0, 1, 1]))   # same centre cell, None slot
    imgs.append(([], []))                                                           # no boxes
    return imgs


def golden_encode_f32(ref_tf):
    res = {}
    for normalized in (False, True):
        outs = [[], [], [], []]
        for boxes, labels in encode_f32_cases():
            for lst, arr in zip(outs, ref_tf.box2hm(boxes, labels, (256, 256), normalized=normalized)):
                lst.append(arr)
        tag = "norm" if normalized else "raw"
        for name, o in zip(("heat", "off", "size", "mask"), outs):
            res[f"{tag}_{name}"] = np.stack(o)
    np.savez_compressed(os.path.join(HERE, "encode_f32.npz"), **res)
    print("encode_f32", {k: v.shape for k, v in res.items()})


if __name__ == "__main__":
    torch.set_num_threads(8)
    ref_hg, ref_loss, ref_tf, ref_ev = import_reference()
    golden_encode(ref_tf)
    golden_encode_f32(ref_tf)
    if os.environ.get("HD_GOLDEN_ONLY") == "encode":
        sys.exit(0)
    golden_decode(ref_tf, ref_ev)
    golden_loss(ref_loss, ref_tf)
    golden_hourglass(ref_hg, ref_loss, ref_tf)
