"""Device GT encoder + image normalisation (csrc/encode.cu) against the reference-generated golden
fixture tests/golden/encode_f32.npz and the oracle (oracle/encode_ref.py), through the C ABI."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _cases():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.encode_f32_cases()


def test_oracle_and_host_box2hm_vs_reference_f32_cases():
    from oracle.encode_ref import encode_boxes
    from real_time_helmet_detection_b200.transform import box2hm
    gold = np.load(os.path.join(GOLD, "encode_f32.npz"))
    for normalized, tag in ((False, "raw"), (True, "norm")):
        for i, (boxes, labels) in enumerate(_cases()):
            for fn in (encode_boxes, box2hm):
                got = fn(boxes, labels, (256, 256), normalized=normalized)
                for name, g in zip(("heat", "off", "size", "mask"), got):
                    assert np.array_equal(g, gold[f"{tag}_{name}"][i]), (fn.__name__, tag, name, i)


def _to_device(cases, nmax=None):
    import torch
    from real_time_helmet_detection_b200.data import pad_boxes
    boxes, labels = pad_boxes([c[0] for c in cases], [c[1] for c in cases], nmax)
    return torch.from_numpy(boxes).cuda(), torch.from_numpy(labels).cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("normalized", [False, True])
def test_encode_targets_vs_reference_golden(normalized):
    import torch
    from real_time_helmet_detection_b200.data import encode_targets
    gold = np.load(os.path.join(GOLD, "encode_f32.npz"))
    tag = "norm" if normalized else "raw"
    boxes, labels = _to_device(_cases())
    heat, off, size, mask, err = encode_targets(boxes, labels, (256, 256), normalized=normalized, return_errors=True)
    assert int(err.item()) == 0
    # index / mask work and the fp64 box arithmetic: bit-exact
    assert np.array_equal(mask.cpu().numpy(), gold[f"{tag}_mask"])
    assert np.array_equal(off.cpu().numpy(), gold[f"{tag}_off"])
    assert np.array_equal(size.cpu().numpy(), gold[f"{tag}_size"])
    # Gaussian: exp() evaluated in fp64 on the device vs numpy's, then rounded to fp32 - tolerance one fp32 ulp
    h, g = heat.cpu().numpy(), gold[f"{tag}_heat"]
    assert np.array_equal(h > 0, g > 0)
    assert np.abs(h - g).max() <= 6e-8, np.abs(h - g).max()
    assert (h == g).mean() > 0.9999


@pytest.mark.gpu
def test_encode_targets_full_size_vs_oracle_and_edge_cases():
    import torch
    from oracle.encode_ref import encode_boxes
    from real_time_helmet_detection_b200.data import encode_targets
    rs = np.random.RandomState(5)
    cases = []
    for b in range(32):
        bx, lb = [], []
        for _ in range(rs.randint(0, 24)):
            x0, y0 = rs.uniform(0, 480, 2)
            bw, bh = rs.uniform(2, 300, 2)
            bx.append([float(np.float32(v)) for v in (x0, y0, min(x0 + bw, 511.0), min(y0 + bh, 511.0))])
            lb.append(int(rs.randint(0, 3)))
        cases.append((bx, lb))
    boxes, labels = _to_device(cases, 32)
    got = encode_targets(boxes, labels, (512, 512), num_cls=3)
    for i, (bx, lb) in enumerate(cases):
        want = encode_boxes(bx, lb, (512, 512), num_cls=3)
        for name, g, w in zip(("heat", "off", "size", "mask"), got, want):
            g = g[i].cpu().numpy()
            if name == "heat":
                assert np.abs(g - w).max() <= 6e-8, (i, np.abs(g - w).max())
            else:
                assert np.array_equal(g, w), (i, name)
    # a centre outside the map (IndexError in the reference) and a label >= num_cls are skipped and counted;
    # a degenerate zero-size box puts NaN at its centre exactly like numpy's 0/0
    bad = [([[600.0, 10.0, 700.0, 50.0], [10.0, 10.0, 50.0, 50.0], [8.0, 8.0, 8.0, 8.0]], [0, 5, 1])]
    b2, l2 = _to_device(bad)
    heat, off, size, mask, err = encode_targets(b2, l2, (512, 512), return_errors=True)
    assert int(err.item()) == 2
    want = encode_boxes([[8.0, 8.0, 8.0, 8.0]], [1], (512, 512))
    assert np.isnan(want[0][1, 2, 2]) and torch.isnan(heat[0, 1, 2, 2])
    assert np.array_equal(np.isnan(want[0]), torch.isnan(heat[0]).cpu().numpy())
    assert np.array_equal(mask[0].cpu().numpy(), want[3])
    # empty slots only
    z = encode_targets(torch.zeros(2, 4, 4, device="cuda"), torch.full((2, 4), -1, device="cuda", dtype=torch.int32), (128, 128))
    assert all(float(t.abs().sum()) == 0.0 for t in z)
    # a crowded image: more boxes than one shared-memory chunk (128) - like the reference's box2hm, no cap; list order
    # decides which box owns a shared centre cell, across chunk boundaries too
    rs = np.random.RandomState(7)
    nb = 300
    x0, y0 = rs.uniform(0, 440, nb), rs.uniform(0, 440, nb)
    crowd = np.stack([x0, y0, x0 + rs.uniform(6, 60, nb), y0 + rs.uniform(6, 60, nb)], 1).astype(np.float32)
    crowd[129] = crowd[3]                                  # same centre cell in chunk 0 and chunk 1: the later one wins
    crowd[129, 2:] += 1.0
    labs = rs.randint(0, 2, nb)
    got = encode_targets(torch.from_numpy(crowd)[None].cuda(), torch.from_numpy(labs.astype(np.int32))[None].cuda(), (512, 512))
    want = encode_boxes(crowd.tolist(), labs.tolist(), (512, 512))
    for name, g_, w_ in zip(("heat", "off", "size", "mask"), got, want):
        g_ = g_[0].cpu().numpy()
        if name == "heat":
            assert np.abs(g_ - w_).max() <= 6e-8, np.abs(g_ - w_).max()
        else:
            assert np.array_equal(g_, w_), name


@pytest.mark.gpu
@pytest.mark.parametrize("pretrained", ["imagenet", "scratch"])
def test_normalize_u8_bit_exact(pretrained):
    import torch
    from real_time_helmet_detection_b200.data import normalize_images, normalizer_constants
    g = torch.Generator().manual_seed(3)
    img = torch.randint(0, 256, (3, 96, 160, 3), generator=g, dtype=torch.uint8)
    img[0, 0, :256 // 3 + 1].view(-1)[:256] = torch.arange(256, dtype=torch.uint8)      # every byte value
    mean, std = normalizer_constants(pretrained)
    # TF.to_tensor (HWC uint8 -> CHW float / 255) followed by Normalize (sub mean, div std), in fp32 like torchvision
    want = img.permute(0, 3, 1, 2).to(torch.float32).div(255)
    want = (want - torch.tensor(mean).view(1, 3, 1, 1)) / torch.tensor(std).view(1, 3, 1, 1)
    got = normalize_images(img.cuda(), pretrained).cpu()
    assert got.shape == want.shape and torch.equal(got, want)
    with pytest.raises(NotImplementedError, match="Not expected dataset pretrained"):
        normalize_images(img.cuda(), "coco")


@pytest.mark.gpu
def test_device_collate_matches_host_collate():
    import torch
    from oracle.encode_ref import encode_boxes
    from real_time_helmet_detection_b200.data import DeviceCollate
    rs = np.random.RandomState(0)
    collate = DeviceCollate("cuda:0", num_cls=2, max_boxes=8)
    for it in range(3):   # three rounds: both staging slots are reused
        imgs = [rs.randint(0, 256, (128, 128, 3)).astype(np.uint8) for _ in range(4)]
        bbs = [[[float(np.float32(v)) for v in (10 + 5 * j + it, 12.5, 60.25 + 7 * j, 90.0)] for j in range(b + 1)] for b in range(4)]
        ids = [[j % 2 for j in range(b + 1)] for b in range(4)]
        image, heat, off, size, mask = collate(imgs, bbs, ids)
        torch.cuda.synchronize()
        for b in range(4):
            want = encode_boxes(bbs[b], ids[b], (128, 128))
            assert np.abs(heat[b].cpu().numpy() - want[0]).max() <= 6e-8
            assert np.array_equal(off[b].cpu().numpy(), want[1]) and np.array_equal(mask[b].cpu().numpy(), want[3])
            t = torch.from_numpy(imgs[b]).permute(2, 0, 1).float().div(255)
            t = (t - torch.tensor([0.485, 0.456, 0.406]).view(3, 1, 1)) / torch.tensor([0.229, 0.224, 0.225]).view(3, 1, 1)
            assert torch.equal(image[b].cpu(), t)
    assert collate.h2d_bytes == 4 * 128 * 128 * 3 + 4 * 8 * 4 * 4 + 4 * 8 * 4
