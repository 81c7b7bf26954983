"""BASELINE.json's full sizes (config 2: 512x512, batch 32; config 5 decode) through size-independent properties - the
oracle is too slow there, so exact algebraic identities stand in for it: linearity under power-of-two scaling (exact
in floating point), batch-permutation equivariance, a checksum of checksums for the fused BN statistics, sortedness /
NMS invariants / idempotence of the decode, and the encode -> decode round trip of the reference's own self-test
(transform.py:112-131) for hundreds of boxes."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("B,H,W", [(32, 128, 128), (8, 256, 256)])
def test_conv_fullsize_linearity_and_stat_checksum(dev, B, H, W):
    from real_time_helmet_detection_b200 import ops
    g = torch.Generator(dev).manual_seed(1)
    x = (torch.randn(B, H, W, 128, device=dev, generator=g)).to(torch.bfloat16)
    w = torch.randn(128, 128, 3, 3, device=dev, generator=g) * 0.03
    wp, wpd = ops.pack_weight(w), ops.pack_weight(w, mode=1)
    stats = torch.zeros(2, 128, device=dev)
    y1 = ops.conv2d_igemm(x, wp, 128, 3, stats=stats)
    y2 = ops.conv2d_igemm(x * 2, wp, 128, 3)
    assert torch.equal(y2.float(), y1.float() * 2)                   # conv(2x) == 2 conv(x), bit for bit
    # checksum of checksums: the epilogue's per-channel sums (taken from the fp32 accumulators) against the stored bf16
    yf = y1.float()
    assert torch.allclose(stats[0], yf.sum(dim=(0, 1, 2)), rtol=2e-3, atol=2e-3 * yf.abs().sum(dim=(0, 1, 2)).max().item())
    assert torch.allclose(stats[1], (yf * yf).sum(dim=(0, 1, 2)), rtol=2e-3)
    # zero padding really is zero: an all-ones input and weight give 9*128 inside, 4*128 at the corners, 6*128 on edges
    ones = torch.ones(1, H, W, 128, device=dev, dtype=torch.bfloat16)
    yo = ops.conv2d_igemm(ones, ops.pack_weight(torch.ones(128, 128, 3, 3, device=dev)), 128, 3).float()
    assert yo[0, 0, 0, 0] == 4 * 128 and yo[0, 0, 5, 7] == 6 * 128 and yo[0, 9, 9, 127] == 9 * 128 and yo[0, H - 1, W - 1, 3] == 4 * 128
    # dgrad and wgrad obey the same identity
    d1 = ops.conv2d_igemm(y1, wpd, 128, 3)
    d2 = ops.conv2d_igemm(y1 * 2, wpd, 128, 3)
    assert torch.equal(d2.float(), d1.float() * 2)
    g1 = ops.conv2d_wgrad(x, y1, 128, 3)
    g2 = ops.conv2d_wgrad(x, y1 * 2, 128, 3)
    assert torch.equal(g2, g1 * 2)
    # <dy, conv(x)> == <wgrad(x, dy), w> (adjointness of the weight gradient), fp32 sums of ~5e8 products
    lhs = (y1.double() * y1.double()).sum().item()
    rhs = (g1.double() * w.to(torch.bfloat16).double()).sum().item()
    assert abs(lhs - rhs) <= 2e-2 * abs(lhs), (lhs, rhs)


def test_config2_network_fullsize(dev):
    from real_time_helmet_detection_b200.hourglass import StackedHourglass
    from real_time_helmet_detection_b200.loss import LossCalculator
    from real_time_helmet_detection_b200.synthetic import synthetic_targets
    from real_time_helmet_detection_b200.train import train_step
    torch.manual_seed(777)
    net = StackedHourglass(1, 128, 6).to(dev)
    B = 32
    x = torch.randn(B, 3, 512, 512, device=dev, generator=torch.Generator(dev).manual_seed(3))
    # eval: a batch permutation permutes the logits, bit for bit (no batch statistics, no atomics on this path)
    net.eval()
    perm = torch.randperm(B, device=dev, generator=torch.Generator(dev).manual_seed(4))
    with torch.no_grad():
        a = net(x)
        b = net(x[perm])
    assert a.shape == (B, 1, 6, 128, 128) and torch.isfinite(a).all()
    assert torch.equal(a[perm], b)
    # train: finite loss and gradients for all 226 tensors, running statistics updated exactly once
    net.train()
    nbt0 = net.state_dict()["pre_layer.layers.0.bn.num_batches_tracked"].clone()
    gts = [torch.from_numpy(t).to(dev) for t in synthetic_targets(B, imsize=512)]
    crit = LossCalculator(1.0, 1.0, 0.1, 2.0, 4.0)
    loss = train_step(net, crit, x, *gts)
    assert torch.isfinite(loss) and 0.1 < float(loss) < 1e3
    params = list(net.parameters())
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in params)
    assert sum(p.numel() for p in params) == 4_984_070
    assert net.state_dict()["pre_layer.layers.0.bn.num_batches_tracked"] == nbt0 + 1
    # stem / neck conv biases are followed by a train-mode BN: their true gradient is zero
    gb = dict(net.named_parameters())["pre_layer.layers.0.convolution.bias"].grad
    gw = dict(net.named_parameters())["pre_layer.layers.0.convolution.weight"].grad
    assert gb.abs().max() <= 5e-2 * gw.abs().max()      # a sum of 2M bf16-rounded terms that cancel analytically


def _iou(a, b):
    x0, y0 = torch.maximum(a[:, None, 0], b[None, :, 0]), torch.maximum(a[:, None, 1], b[None, :, 1])
    x1, y1 = torch.minimum(a[:, None, 2], b[None, :, 2]), torch.minimum(a[:, None, 3], b[None, :, 3])
    inter = (x1 - x0).clamp(min=0) * (y1 - y0).clamp(min=0)
    area = lambda t: (t[:, 2] - t[:, 0]) * (t[:, 3] - t[:, 1])
    return inter / (area(a)[:, None] + area(b)[None, :] - inter)


def test_decode_fullsize_invariants(dev):
    from real_time_helmet_detection_b200.evaluate import Prediction
    B, S, C, H, W, k = 32, 2, 2, 128, 128, 100
    g = torch.Generator(dev).manual_seed(9)
    out = torch.randn(B, S, C + 4, H, W, device=dev, generator=g)
    out[:, :, :C] = out[:, :, :C] * 2 - 3                    # logits: a few thousand peaks above the threshold
    out[:, :, C:C + 2] = torch.rand(B, S, 2, H, W, device=dev, generator=g)
    out[:, :, C + 2:] = torch.rand(B, S, 2, H, W, device=dev, generator=g) * 20 + 4
    pred = Prediction(None, k, 4, 0.2, "nms", 0.2)
    boxes, clss, scores = pred.decode(out)
    again = pred.decode(out)
    assert len(boxes) == B
    for b in range(B):
        n = scores[b].numel()
        assert 0 < n <= S * k and boxes[b].shape == (n, 4) and clss[b].dtype == torch.int64
        assert (scores[b][:-1] >= scores[b][1:]).all() and (scores[b] >= 0.2).all()        # sorted, thresholded
        assert ((clss[b] >= 0) & (clss[b] < C)).all() and torch.isfinite(boxes[b]).all()
        iou = _iou(boxes[b], boxes[b])
        iou.fill_diagonal_(0)
        assert iou.max() <= 0.2 + 1e-6                                                      # NMS invariant
        assert torch.equal(boxes[b], again[0][b]) and torch.equal(scores[b], again[2][b])   # deterministic
    # NMS is idempotent: re-running it on its own output keeps every box (checked through the IoU bound above) and the
    # un-suppressed decode (do_nms off) is a superset whose first element is the same top score
    raw = Prediction(None, k, 4, 0.2, "nms", 1.0).decode(out)
    for b in range(B):
        assert raw[2][b].numel() >= scores[b].numel() and raw[2][b][0] == scores[b][0]


def test_encode_decode_round_trip_fullsize(dev):
    """box -> box2hm (device encoder) -> hm2box gives the box back: the reference's self-test (transform.py:112-131) for
    32 images x 24 boxes whose centre cells are distinct."""
    from real_time_helmet_detection_b200.data import encode_targets
    from real_time_helmet_detection_b200.transform import hm2box
    rs = np.random.RandomState(0)
    B, n = 32, 24
    boxes = np.zeros((B, n, 4), np.float32)
    labels = np.zeros((B, n), np.int32)
    for b in range(B):
        cells = rs.choice(np.arange(4, 124)[::3].repeat(1), size=(n, 2), replace=True)
        cells = np.unique(cells, axis=0)
        while len(cells) < n:
            cells = np.unique(np.concatenate([cells, rs.choice(np.arange(4, 124)[::3], size=(n, 2))]), axis=0)
        cells = cells[rs.permutation(len(cells))[:n]]
        cx, cy = (cells[:, 0] + rs.uniform(0.05, 0.95, n)) * 4, (cells[:, 1] + rs.uniform(0.05, 0.95, n)) * 4
        bw, bh = rs.uniform(8, 40, n), rs.uniform(8, 40, n)
        boxes[b] = np.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1)
        labels[b] = rs.randint(0, 2, n)
    for normalized in (False, True):
        heat, off, size, mask = encode_targets(torch.from_numpy(boxes).to(dev), torch.from_numpy(labels).to(dev), (512, 512),
                                               normalized=normalized)
        assert int(mask.sum()) == B * n
        for b in range(0, B, 5):
            bx, cl, sc = hm2box(heat[b], off[b], size[b], scale_factor=4, topk=n, conf_th=0.999, normalized=normalized)
            assert sc.numel() == n and (sc == 1.0).all()
            want = boxes[b][np.lexsort((boxes[b][:, 1], boxes[b][:, 0]))]
            got = bx.cpu().numpy()
            got = got[np.lexsort((got[:, 1], got[:, 0]))]
            assert np.abs(got - want).max() < 2e-3, np.abs(got - want).max()
