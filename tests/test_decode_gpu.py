"""GPU parity of the fused decode + NMS kernel (csrc/decode.cu) against the golden outputs of the unmodified
reference (tests/golden/decode.npz) and the oracle (oracle/decode_ref.py).
Bar (north_star): identical top-k indices / classes / box coordinates (bit-exact fp32); scores within 1e-6
(sigmoid is evaluated with the device expf, the CPU reference with Sleef)."""
import os
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _pred(**kw):
    from real_time_helmet_detection_b200.evaluate import Prediction
    args = dict(network=None, topk=100, scale_factor=4, conf_th=0.2, nms="nms", nms_th=0.2)
    args.update(kw)
    return Prediction(**args)


@pytest.mark.parametrize("S", [1, 2])
@pytest.mark.parametrize("norm", [False, True])
def test_prediction_vs_reference_golden(cuda_device, S, norm):
    from oracle.decode_ref import synthetic_head
    gold = np.load(os.path.join(GOLD, "decode.npz"))
    head = torch.from_numpy(synthetic_head(S=S, seed=S - 1)).to(cuda_device)
    b, c, s = _pred(normalized_coord=norm).decode(head)
    tag = f"pred_s{S}_{'norm' if norm else 'lin'}"
    assert c[0].dtype == torch.int64 and b[0].dtype == torch.float32
    assert np.array_equal(c[0].cpu().numpy(), gold[tag + "_cls"])
    if norm:   # sigmoid(offset/size) differs by <= 1 ulp between the device expf and the CPU reference
        assert np.allclose(b[0].cpu().numpy(), gold[tag + "_boxes"], rtol=1e-5, atol=1e-4)
    else:
        assert np.array_equal(b[0].cpu().numpy(), gold[tag + "_boxes"])
    assert np.allclose(s[0].cpu().numpy(), gold[tag + "_scores"], rtol=0, atol=1e-6)


def test_hm2box_vs_reference_golden(cuda_device):
    from oracle.decode_ref import synthetic_head
    from real_time_helmet_detection_b200.transform import hm2box
    gold = np.load(os.path.join(GOLD, "decode.npz"))
    for S in (1, 2):
        head = torch.from_numpy(synthetic_head(S=S, seed=S - 1))
        hm = torch.sigmoid(head[0, 0, :2])          # CPU sigmoid: bit-identical scores to the golden run
        b, c, s = hm2box(hm.to(cuda_device), head[0, 0, 2:4].to(cuda_device), head[0, 0, 4:6].to(cuda_device),
                         scale_factor=4, topk=100, conf_th=0.2)
        assert np.array_equal(c.cpu().numpy(), gold[f"hm2box_s{S}_cls"])
        assert np.array_equal(b.cpu().numpy(), gold[f"hm2box_s{S}_boxes"])
        assert np.array_equal(s.cpu().numpy(), gold[f"hm2box_s{S}_scores"])


def test_plateau_and_kat(cuda_device):
    from real_time_helmet_detection_b200.transform import hm2box, box2hm
    gold = np.load(os.path.join(GOLD, "decode.npz"))
    hm = torch.zeros(2, 8, 8)
    hm[0, 3, 3] = hm[0, 3, 4] = 0.9
    hm[1, 6, 1] = 0.7
    off, wh = torch.full((2, 8, 8), 0.25), torch.full((2, 8, 8), 2.0)
    b, c, s = hm2box(hm.to(cuda_device), off.to(cuda_device), wh.to(cuda_device), scale_factor=4, topk=5, conf_th=0.3)
    assert np.array_equal(s.cpu().numpy(), gold["plateau_scores"])
    # the two plateau cells tie: the reference's order is implementation-defined, compare as sets
    ref = {tuple(r) for r in np.concatenate([gold["plateau_boxes"], gold["plateau_cls"][:, None]], 1).tolist()}
    got = {tuple(r) for r in np.concatenate([b.cpu().numpy(), c.cpu().numpy()[:, None]], 1).tolist()}
    assert ref == got
    # known-answer self-test of the reference (transform.py:112-131)
    enc = np.load(os.path.join(GOLD, "encode.npz"))
    heat, o, w, m = box2hm([[10, 20, 100, 200]], [1], (512, 512), normalized=True)
    assert np.array_equal(heat[:, 27, 13], enc["kat_heat"]) and np.array_equal(o[:, 27, 13], enc["kat_off"])
    assert np.array_equal(w[:, 27, 13], enc["kat_wh"])
    bb, cc, ss = hm2box(*(torch.from_numpy(t).to(cuda_device) for t in (heat, o, w)), normalized=True)
    assert np.array_equal(bb.cpu().numpy(), enc["kat_box"]) and np.array_equal(cc.cpu().numpy(), enc["kat_cls"])


def test_edge_cases(cuda_device):
    from oracle import decode_ref
    from real_time_helmet_detection_b200.transform import hm2box
    g = torch.Generator().manual_seed(11)
    # conf_th = 0 keeps zero-score fillers: exactly k rows, positive part identical to the oracle
    hm = torch.rand(2, 16, 16, generator=g)
    off, wh = torch.rand(2, 16, 16, generator=g), torch.rand(2, 16, 16, generator=g) * 8
    b, c, s = hm2box(hm.to(cuda_device), off.to(cuda_device), wh.to(cuda_device), topk=200, conf_th=0.0)
    rb, rc, rs = decode_ref.hm2box(hm.numpy(), off.numpy(), wh.numpy(), topk=200, conf_th=0.0)
    assert b.shape == (200, 4) and np.array_equal(s.cpu().numpy(), rs)
    assert np.array_equal(b.cpu().numpy(), rb) and np.array_equal(c.cpu().numpy(), rc)   # same tie-break as the oracle
    # k larger than the map raises like torch.topk
    with pytest.raises(RuntimeError):
        hm2box(hm.to(cuda_device), off.to(cuda_device), wh.to(cuda_device), topk=2 * 16 * 16 + 1)
    # odd, non power-of-two map and a map too large for the shared-memory score buffer
    for (C, H, W, k) in ((3, 40, 24, 50), (2, 160, 168, 100)):
        hm = torch.rand(C, H, W, generator=g)
        off, wh = torch.rand(2, H, W, generator=g), torch.rand(2, H, W, generator=g) * 8
        b, c, s = hm2box(hm.to(cuda_device), off.to(cuda_device), wh.to(cuda_device), topk=k, conf_th=0.3)
        rb, rc, rs = decode_ref.hm2box(hm.numpy(), off.numpy(), wh.numpy(), topk=k, conf_th=0.3)
        assert np.array_equal(b.cpu().numpy(), rb) and np.array_equal(c.cpu().numpy(), rc)
        assert np.array_equal(s.cpu().numpy(), rs)


def test_batched_prediction_vs_oracle(cuda_device):
    from oracle import decode_ref
    heads = np.concatenate([decode_ref.synthetic_head(S=2, seed=s) for s in (3, 4, 5)], 0)
    b, c, s = _pred().decode(torch.from_numpy(heads).to(cuda_device))
    rb, rc, rs = decode_ref.predict(heads)
    for i in range(3):
        assert np.array_equal(c[i].cpu().numpy(), rc[i])
        assert np.array_equal(b[i].cpu().numpy(), rb[i])
        assert np.allclose(s[i].cpu().numpy(), rs[i], atol=1e-6, rtol=0)
    # class-agnostic suppression (evaluate.py:174): a class-1 box overlapping a better class-0 box disappears
    gold = np.load(os.path.join(GOLD, "decode.npz"))
    assert gold["agnostic_boxes"].shape[0] == 2
    with pytest.raises(NotImplementedError):
        _pred(nms="soft-nms").decode(torch.from_numpy(heads).to(cuda_device))
    with pytest.raises(NotImplementedError):
        _pred(nms="bogus").decode(torch.from_numpy(heads).to(cuda_device))


def test_evaluate_step_and_writers(cuda_device, tmp_path):
    """evaluate_step + save_predictions reproduce the reference's eval artefacts (evaluate.py:40-99): dict of (n,6)
    float64 arrays in original-image pixels, the pickle, and the "%d %f %d %d %d %d" txt lines."""
    import pickle
    import types
    from real_time_helmet_detection_b200.evaluate import (Prediction, evaluate_step, save_predictions,
                                                          resize_box_to_original_scale)
    from real_time_helmet_detection_b200.synthetic import synthetic_head

    class FakeNet(torch.nn.Module):           # a "network" that returns the config-5 blob head for every image
        def forward(self, x):
            return torch.from_numpy(synthetic_head(S=1)).to(x.device).repeat(x.shape[0], 1, 1, 1, 1)

    pred = Prediction(FakeNet(), 100, 4, 0.2, "nms", 0.2)
    sizes = [(640, 480), (500, 375), (512, 512)]
    batches = [(torch.zeros(2, 3, 512, 512), None, None, None, None,
                [{"annotation": {"filename": f"img{i}.jpg", "size": {"width": str(sizes[i][0]), "height": str(sizes[i][1])}}}
                 for i in range(2)]),
               (torch.zeros(1, 3, 512, 512), None, None, None, None,
                [{"annotation": {"filename": "img2.jpg", "size": {"width": "512", "height": "512"}}}])]
    args = types.SimpleNamespace(imsize=512)
    res = evaluate_step(batches, pred, cuda_device, args)
    boxes, clss, scores = pred(torch.zeros(1, 3, 512, 512, device=cuda_device))
    assert sorted(res) == ["img0.jpg", "img1.jpg", "img2.jpg"]
    for i in range(3):
        r = res[f"img{i}.jpg"]
        assert r.dtype == np.float64 and r.shape == (44, 6)
        assert np.array_equal(r[:, 0], clss[0].cpu().numpy()) and np.allclose(r[:, 1], scores[0].cpu().numpy())
        want = resize_box_to_original_scale(boxes[0].cpu().numpy(), sizes[i], (512, 512))
        assert np.allclose(r[:, 2:], want, rtol=1e-6, atol=1e-4)
    save_predictions(res, str(tmp_path))
    back = pickle.load(open(tmp_path / "prediction_results.pickle", "rb"))
    assert all(np.array_equal(back[k], res[k]) for k in res)
    lines = open(tmp_path / "txt" / "img1.txt").read().splitlines()
    assert len(lines) == 44
    f0 = lines[0].split()
    r = res["img1.jpg"][0]
    assert int(f0[0]) == int(r[0]) and abs(float(f0[1]) - r[1]) < 1e-6 and [int(v) for v in f0[2:]] == [int(v) for v in r[2:]]
