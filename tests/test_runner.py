"""Deployment path: the HDW1 weight export and the native runner `runner/hd_infer` (plain C++ over
the C ABI, no LibTorch) against the Python `Prediction` on the same image."""
import os
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUNNER = os.path.join(ROOT, "runner", "hd_infer")


def _net(S=1, seed=5):
    from real_time_helmet_detection_b200.hourglass import StackedHourglass
    torch.manual_seed(seed)
    net = StackedHourglass(S, 128, 6)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():          # non-trivial running statistics and affine parameters, as after training
        for name, buf in net.named_buffers():
            if name.endswith("running_mean"):
                buf.copy_(torch.randn(buf.shape, generator=g) * 0.1)
            elif name.endswith("running_var"):
                buf.copy_(torch.rand(buf.shape, generator=g) + 0.5)
        for name, p in net.named_parameters():
            if name.endswith("bn.weight"):
                p.copy_(torch.rand(p.shape, generator=g) + 0.5)
            elif name.endswith("bn.bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
    return net


def test_weight_export_round_trip(tmp_path):
    from real_time_helmet_detection_b200.export import export_weights, load_weights
    a, b = _net(2, 5), _net(2, 99)
    path = export_weights(a, str(tmp_path / "m.hdw"))
    n_float = sum(v.numel() for k, v in a.state_dict().items() if v.dtype.is_floating_point)
    n_units = len(a.units())
    assert os.path.getsize(path) == 4 + 16 + 20 * n_units + 4 * n_float
    load_weights(b, path)
    sa, sb = a.state_dict(), b.state_dict()
    assert all(torch.equal(sa[k], sb[k]) for k in sa if sa[k].dtype.is_floating_point)
    with pytest.raises(RuntimeError, match="mismatch"):
        load_weights(_net(1), path)


def test_runner_is_built_and_rejects_bad_input(tmp_path):
    assert os.path.exists(RUNNER), "runner/hd_infer missing: run __graft_entry__.build()"
    r = subprocess.run([RUNNER], capture_output=True, text=True)
    assert r.returncode == 1 and "usage: hd_infer" in r.stderr
    bad = tmp_path / "bad.hdw"
    bad.write_bytes(b"nope")
    r = subprocess.run([RUNNER, "-m", str(bad), "--random", "128"], capture_output=True, text=True)
    assert r.returncode == 1 and "not an HDW1 weight file" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("S", [1, 2])
def test_runner_matches_python_prediction(tmp_path, S):
    from real_time_helmet_detection_b200.data import normalize_images
    from real_time_helmet_detection_b200.evaluate import Prediction
    from real_time_helmet_detection_b200.export import export_weights
    dev = torch.device("cuda:0")
    net = _net(S).to(dev).eval()
    model = export_weights(net, str(tmp_path / "m.hdw"))
    rs = np.random.RandomState(3)
    img = rs.randint(0, 256, (192, 256, 3)).astype(np.uint8)
    ppm = tmp_path / "img.ppm"
    with open(ppm, "wb") as f:
        f.write(b"P6\n# test image\n256 192\n255\n")
        f.write(img.tobytes())
    csv = tmp_path / "out.csv"
    r = subprocess.run([RUNNER, "-m", model, "-i", str(ppm), "--topk", "50", "--conf-th", "0.3", "--nms-th", "0.3",
                        "--iters", "5", "--csv", str(csv)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "model load!" in r.stdout and "Inference Time:" in r.stdout and "CUDA graph replay:" in r.stdout
    x = normalize_images(torch.from_numpy(img).to(dev).unsqueeze(0))
    boxes, clss, scores = Prediction(net, 50, 4, 0.3, "nms", 0.3)(x)
    got = np.loadtxt(csv, delimiter=",", ndmin=2)
    assert got.shape[0] == scores[0].numel() > 0
    assert np.array_equal(got[:, 4].astype(np.int64), clss[0].cpu().numpy())
    assert np.allclose(got[:, :4], boxes[0].cpu().numpy(), rtol=0, atol=1e-4)
    assert np.allclose(got[:, 5], scores[0].cpu().numpy(), rtol=0, atol=1e-6)
    assert f"index: {got.shape[0] - 1}," in r.stdout          # the reference demo's per-detection print
