"""CPU tests: the oracle (oracle/*.py) against the golden outputs of the UNMODIFIED reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py in the build container). This is what pins the oracle.
Also pins that our nn.Module reproduces the reference's initial weights (same RNG stream) and state_dict layout."""
import os
import sys
import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLD)


def _gt(size, batch):
    from make_golden import FIXED_BOXES
    from oracle.encode_ref import encode_boxes
    outs = [[], [], [], []]
    for b in range(batch):
        boxes, labels = FIXED_BOXES[b % len(FIXED_BOXES)]
        boxes = [[v * size / 256.0 for v in bx] for bx in boxes]
        for lst, arr in zip(outs, encode_boxes(boxes, labels, (size, size))):
            lst.append(arr)
    return [torch.from_numpy(np.stack(o)) for o in outs]


@pytest.mark.parametrize("S,size", [(1, 128), (2, 128), (1, 192)])
def test_hourglass_oracle_vs_reference(S, size):
    from oracle import hourglass_ref, loss_ref
    from real_time_helmet_detection_b200.hourglass import StackedHourglass
    gold = np.load(os.path.join(GOLD, f"hourglass_s{S}_{size}.npz"))
    torch.manual_seed(777)
    net = StackedHourglass(S, 128, 6)
    names = [n for n, _ in net.named_parameters()]
    assert names == list(gold["param_names"])                       # state_dict / parameter order of the reference
    assert len(net.state_dict()) == (226 if S == 1 else 407)
    sums = np.asarray([p.double().sum().item() for p in net.parameters()])
    assert np.allclose(sums, gold["param_sums"], rtol=0, atol=1e-9)   # identical initial weights
    sd = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone())
          for k, v in net.state_dict().items()}
    x = torch.randn(2, 3, size, size, generator=torch.Generator().manual_seed(1))
    new_stats = {}
    out = hourglass_ref.stacked_hourglass_forward(sd, x, training=True, new_stats=new_stats)
    ref = torch.from_numpy(gold["out_train"])
    assert (out.detach() - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()
    gts = _gt(size, 2)
    total, comps = 0, []
    for s in range(S):
        hm, off, sz, tot = loss_ref.losses_from_logits(out[:, s], *gts)
        comps.append([hm.item(), off.item(), sz.item(), tot.item()])
        total = total + tot
    assert np.allclose(np.asarray(comps), gold["loss_components"], rtol=1e-4)
    assert abs(total.item() - float(gold["loss_total"])) <= 1e-4 * abs(float(gold["loss_total"]))
    total.backward()
    norms = np.asarray([sd[n].grad.norm().item() for n in names])
    ref_norms = gold["grad_norms"]
    big = ref_norms > 1e-3 * ref_norms.max()                            # conv biases in front of a BN have ~0 gradient
    assert np.allclose(norms[big], ref_norms[big], rtol=5e-3)
    g = sd["head_lst.0.layer.convolution.weight"].grad.numpy()
    assert np.allclose(g, gold["grad::head_lst.0.layer.convolution.weight"], rtol=1e-3, atol=1e-4 * np.abs(g).max())
    # running statistics after one training step
    off = 0
    for k in gold["rstat_names"]:
        v = new_stats[str(k)].numpy().ravel()
        r = gold["rstat_values"][off:off + v.size]
        off += v.size
        assert np.allclose(v, r, rtol=1e-3, atol=1e-5), k
    # eval mode with the updated running statistics
    sd_eval = {k: v.detach() for k, v in sd.items()}
    sd_eval.update(new_stats)
    oe = hourglass_ref.stacked_hourglass_forward(sd_eval, x, training=False)
    re = torch.from_numpy(gold["out_eval"])
    assert (oe - re).abs().max().item() <= 1e-4 * re.abs().max().item()


def test_loss_oracle_vs_reference():
    from oracle import loss_ref
    gold = np.load(os.path.join(GOLD, "loss.npz"))
    logits = torch.from_numpy(gold["logits"])
    B, h = logits.shape[0], logits.shape[2]
    gts = {"boxes": _gt(128, B),
           "nopos": [torch.zeros(B, 2, h, h), torch.zeros(B, 2, h, h), torch.zeros(B, 2, h, h), torch.zeros(B, 1, h, h)]}
    for name, gt in gts.items():
        for norm in (False, True):
            key = f"{name}_{'norm' if norm else 'lin'}"
            lg = logits.clone().requires_grad_(True)
            vals = loss_ref.losses_from_logits(lg, *gt, normalized_coord=norm)
            vals[3].backward()
            assert np.allclose([v.item() for v in vals], gold[key + "_values"], rtol=1e-5), key
            assert np.allclose(lg.grad.numpy(), gold[key + "_dlogits"], rtol=1e-4, atol=1e-9), key


def test_decode_oracle_vs_reference():
    from oracle import decode_ref
    gold = np.load(os.path.join(GOLD, "decode.npz"))
    for S in (1, 2):
        head = decode_ref.synthetic_head(S=S, seed=S - 1)
        for norm in (False, True):
            b, c, s = decode_ref.predict(head, normalized_coord=norm)
            tag = f"pred_s{S}_{'norm' if norm else 'lin'}"
            assert np.array_equal(c[0], gold[tag + "_cls"])
            assert np.allclose(b[0], gold[tag + "_boxes"], rtol=1e-6, atol=1e-4)
            assert np.allclose(s[0], gold[tag + "_scores"], atol=1e-6, rtol=0)
        hm = decode_ref.sigmoid_f32(head[0, 0, :2])
        b, c, s = decode_ref.hm2box(hm, head[0, 0, 2:4], head[0, 0, 4:6], topk=100, conf_th=0.2)
        assert np.array_equal(c, gold[f"hm2box_s{S}_cls"]) and np.allclose(b, gold[f"hm2box_s{S}_boxes"], atol=1e-4)
    # class-agnostic NMS known answer (evaluate.py:174)
    keep = decode_ref.nms(np.array([[0, 0, 10, 10], [1, 1, 11, 11], [20, 20, 30, 30]], np.float32),
                          np.array([0.9, 0.8, 0.7], np.float32), 0.2)
    assert keep.tolist() == [0, 2] and gold["agnostic_cls"].tolist() == [0, 1]
    # plateau: both equal neighbours are peaks (transform.py:78)
    hm = np.zeros((2, 8, 8), np.float32)
    hm[0, 3, 3] = hm[0, 3, 4] = 0.9
    hm[1, 6, 1] = 0.7
    b, c, s = decode_ref.hm2box(hm, np.full((2, 8, 8), 0.25, np.float32), np.full((2, 8, 8), 2.0, np.float32), topk=5,
                                conf_th=0.3)
    assert np.array_equal(s, gold["plateau_scores"]) and sorted(map(tuple, b.tolist())) == sorted(map(tuple, gold["plateau_boxes"].tolist()))


def test_encode_oracle_and_product_box2hm_vs_reference():
    from oracle.encode_ref import encode_boxes, synthetic_targets
    from real_time_helmet_detection_b200.transform import box2hm
    gold = np.load(os.path.join(GOLD, "encode.npz"))
    for fn in (encode_boxes, box2hm):
        heat, off, wh, mask = fn([[10, 20, 100, 200]], [1], (512, 512), normalized=True)
        assert np.array_equal(heat[:, 27, 13], gold["kat_heat"]) and np.array_equal(off[:, 27, 13], gold["kat_off"])
        assert np.array_equal(wh[:, 27, 13], gold["kat_wh"]) and mask[0, 27, 13] == 1.0
    ghm, goff, gsz, gmask = synthetic_targets(4)
    assert np.allclose(ghm.sum(axis=(2, 3)), gold["enc_heat_sum"], rtol=1e-6)
    assert np.array_equal(np.argwhere(gmask > 0), gold["enc_mask_idx"])
    assert np.array_equal(goff[gmask.repeat(2, 1) > 0], gold["enc_off"])
