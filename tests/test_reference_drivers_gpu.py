"""The drop-in claim, exercised with the REFERENCE'S OWN drivers instead of a restatement.

baseline/_ref/ holds byte-for-byte copies of the reference's modules (tools/stage_reference.py; git-ignored, travels to the
GPU box). baseline.refload.load_reference_drivers() imports the reference's `train.py` / `evaluate.py` / `optim.py` /
`config.py` with the three shim modules of INTEGRATION.md section 1 ahead of them on sys.path - so `from hourglass import
StackedHourglass`, `from loss import LossCalculator`, `from transform import hm2box` inside the reference resolve to the
B200 package - and with the one-token `squeeze_(1)` fix torch >= 2 needs (baseline/_ref/patched/squeeze_patch.diff).

What runs here is therefore train.py:164-201 `load_network` (4-tuple, Adam + MultiStepLR from optim.py:3-12),
train.py:86-162 `train_step` (fp16 autocast + GradScaler with --amp, --sub-divisions 2 accumulation, the
`loss_calculator.get_log()` print branch incl. utils.blend_heatmap), the checkpoint dict of train.py:76-82 -> resume through
`--model-load` (train.py:190-199), and evaluate.py:58-99 `evaluate_step` with the reference's own `Prediction`
(evaluate.py:114-182: per-image / per-stack loops over OUR hm2box + torchvision NMS) beside our fused `Prediction`.
"""
import collections
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def drivers():
    from baseline import refload
    if not refload.available():
        pytest.skip("baseline/_ref not staged (python tools/stage_reference.py in the build container)")
    return refload.load_reference_drivers()


def _args(drivers, tmp, extra=()):
    """The reference's own argparse surface (config.py:11-136), parsed from a CLI list."""
    argv = sys.argv
    sys.argv = ["main.py", "--train-flag", "--gpu-no", "0", "--batch-size", "2", "--save-path", str(tmp), *extra]
    try:
        return drivers["config"].build_parser()
    finally:
        sys.argv = argv


def _batches(n, size, B=2, seed=0):
    """What data.py's collate_fn yields: (image, gt_heatmap, gt_offset, gt_size, gt_mask, gt_dict) on the host."""
    from real_time_helmet_detection_b200.synthetic import synthetic_targets
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(n):
        image = torch.randn(B, 3, size, size, generator=g)
        gts = [torch.from_numpy(a) for a in synthetic_targets(B, imsize=size)]
        info = [{"annotation": {"filename": f"img_{i}_{b}.jpg", "size": {"width": str(640 + 32 * b), "height": "480"}}}
                for b in range(B)]
        out.append((image, *gts, info))
    return out


def test_reference_train_loop_amp_accumulation_checkpoint_resume(drivers, cuda_device, tmp_path, capsys):
    train = drivers["train"]
    from real_time_helmet_detection_b200.hourglass import StackedHourglass
    from real_time_helmet_detection_b200.loss import LossCalculator
    assert train.StackedHourglass is StackedHourglass and train.LossCalculator is LossCalculator   # the shims took
    os.makedirs(tmp_path / "training_log", exist_ok=True)
    args = _args(drivers, tmp_path, ["--amp", "--sub-divisions", "2", "--print-interval", "2", "--lr", "1e-3",
                                     "--lr-milestone", "1", "2"])
    torch.manual_seed(args.random_seed)
    network, optimizer, scheduler, loss_calculator = train.load_network(args, cuda_device)      # train.py:164-201
    assert isinstance(network, StackedHourglass) and isinstance(optimizer, torch.optim.Adam)
    assert isinstance(scheduler, torch.optim.lr_scheduler.MultiStepLR) and isinstance(loss_calculator, LossCalculator)
    p0 = [p.detach().clone() for p in network.parameters()]
    scaler = torch.cuda.amp.GradScaler()                                                         # train.py:63
    data = _batches(4, 128)
    train.train_step(data, network, loss_calculator, optimizer, scheduler, scaler, 0, 0, args)   # train.py:86-162
    log = capsys.readouterr().out
    assert "Iteration [   2/   4]" in log and "Loss [hm:" in log and "offset:" in log          # get_log branch ran
    assert os.path.exists(tmp_path / "training_log" / "training_pred.png")                      # blend_heatmap branch
    assert len(loss_calculator.log["total"]) == 4 and all(np.isfinite(loss_calculator.log["total"]))
    # --sub-divisions 2: two optimizer steps for four iterations, and the weights moved
    assert int(optimizer.state[next(iter(network.parameters()))]["step"]) == 2
    assert any(not torch.equal(a, b.detach()) for a, b in zip(p0, network.parameters()))
    scheduler.step()
    # checkpoint exactly as train.py:76-82 writes it, then resume through --model-load (train.py:190-199)
    ckpt = str(tmp_path / "check_point_1.pth")
    torch.save({"epoch": 1, "state_dict": network.state_dict(), "optimizer": optimizer.state_dict(),
                "scheduler": scheduler.state_dict(), "scaler": scaler.state_dict(), "loss_log": loss_calculator.log}, ckpt)
    args2 = _args(drivers, tmp_path, ["--amp", "--sub-divisions", "2", "--print-interval", "100", "--lr", "1e-3",
                                      "--lr-milestone", "1", "2", "--model-load", ckpt])
    # torch >= 2.6 unpickles with weights_only=True by default; the scheduler state holds a collections.Counter
    torch.serialization.add_safe_globals([collections.Counter])
    net2, opt2, sch2, crit2 = train.load_network(args2, cuda_device)
    for (k, a), (_, b) in zip(network.state_dict().items(), net2.state_dict().items()):
        assert torch.equal(a, b), k
    assert crit2.log["total"] == loss_calculator.log["total"] and sch2.last_epoch == 1
    assert abs(opt2.param_groups[0]["lr"] - 1e-4) < 1e-12                                       # milestone 1 applied
    before = len(crit2.log["total"])
    train.train_step(_batches(2, 128, seed=5), net2, crit2, opt2, sch2, torch.cuda.amp.GradScaler(), 1, 0, args2)
    assert len(crit2.log["total"]) == before + 2 and np.isfinite(crit2.log["total"][-1])
    assert int(opt2.state[next(iter(net2.parameters()))]["step"]) == 3
    # the same batches, same weights: the loss right after resume is what the first run would give in eval of the loop
    assert crit2.log["total"][-1] < 1e3


def test_reference_train_loop_matches_native_train_step(drivers, cuda_device, tmp_path):
    """train.py's loop body (per-stack split -> sigmoid -> LossCalculator.__call__) and our fused `train_step`
    (forward_logits) produce the same losses and the same updated weights from the same state, no AMP."""
    train = drivers["train"]
    from real_time_helmet_detection_b200.train import train_step
    args = _args(drivers, tmp_path, ["--num-stack", "2", "--print-interval", "1000", "--lr", "1e-3"])
    torch.manual_seed(1)
    net_a, opt_a, sch_a, crit_a = train.load_network(args, cuda_device)
    torch.manual_seed(1)
    net_b, opt_b, sch_b, crit_b = train.load_network(args, cuda_device)
    data = _batches(2, 128, seed=3)
    train.train_step(data, net_a, crit_a, opt_a, sch_a, None, 0, 0, args)
    net_b.train()
    for image, ghm, goff, gsz, gmask, _ in data:
        train_step(net_b, crit_b, image, ghm, goff, gsz, gmask, optimizer=opt_b)
    la, lb = crit_a.log["total"], crit_b.log["total"]
    assert len(la) == len(lb) == 4                              # 2 stacks x 2 iterations, one log entry per stack
    # first iteration: same weights, same batch -> only the atomics-order noise of the BN statistics (at 128x128 the deepest
    # hourglass level normalises over 2 x 2 x 2 = 8 samples, measured ~1e-3 between two runs of the same binary); second
    # iteration: that noise has gone through one Adam step (every gradient normalised to ~ +-lr) - measured ~2e-2
    np.testing.assert_allclose(la[:2], lb[:2], rtol=5e-3)
    np.testing.assert_allclose(la[2:], lb[2:], rtol=6e-2)
    # the bulk of the 9 M weights agree closely after the two steps (a sign flip of a near-zero gradient moves a weight by 2*lr)
    close = total = 0
    for (k, a), (_, b) in zip(net_a.state_dict().items(), net_b.state_dict().items()):
        if a.dtype.is_floating_point and "running" not in k:
            close += int(((a - b).abs() <= 5e-4).sum())
            total += a.numel()
    assert close >= 0.6 * total, (close, total)


def test_reference_evaluate_step_and_prediction(drivers, cuda_device, tmp_path):
    """evaluate.py:58-99 `evaluate_step` driving (a) the reference's own `Prediction` class over our network + hm2box and
    (b) our fused `Prediction`: identical result dictionaries (class, score, rescaled box per image)."""
    evaluate = drivers["evaluate"]
    from real_time_helmet_detection_b200.evaluate import Prediction, evaluate_step, save_predictions
    from real_time_helmet_detection_b200.hourglass import StackedHourglass
    from real_time_helmet_detection_b200.synthetic import synthetic_head
    args = _args(drivers, tmp_path, ["--imsize", "512", "--topk", "100", "--conf-th", "0.2", "--nms-th", "0.2"])

    class Blobs(torch.nn.Module):
        """A "network" whose output is the config-5 blob tensor (every decode stage has work), shifted per image."""
        def __init__(self, S):
            super().__init__()
            self.head = torch.from_numpy(synthetic_head(S=S)).to(cuda_device)

        def forward(self, x):
            return torch.cat([torch.roll(self.head, shifts=7 * b, dims=-1) for b in range(x.shape[0])], dim=0)

    data = _batches(2, 512, B=2, seed=1)
    for S in (1, 2):
        net = Blobs(S)
        ref_pred = evaluate.Prediction(net, args.topk, args.scale_factor, args.conf_th, args.nms, args.nms_th)
        ours = Prediction(net, args.topk, args.scale_factor, args.conf_th, args.nms, args.nms_th)
        with torch.no_grad():
            res_ref = evaluate.evaluate_step(data, ref_pred, cuda_device, args)       # the reference's loop + Prediction
        res_ours = evaluate_step(data, ours, cuda_device, args)
        assert sorted(res_ref) == sorted(res_ours) and len(res_ref) == 4
        for k in res_ref:
            a, b = np.asarray(res_ref[k], np.float64), res_ours[k]
            assert a.shape == b.shape and a.shape[0] > 10, (k, a.shape, b.shape)
            assert np.array_equal(a[:, 0], b[:, 0])                                   # classes
            np.testing.assert_allclose(a[:, 1], b[:, 1], rtol=0, atol=1e-6)           # scores (sigmoid on device)
            np.testing.assert_allclose(a[:, 2:], b[:, 2:], rtol=1e-6, atol=1e-4)      # boxes in original-image pixels
    save_predictions(res_ours, str(tmp_path / "out"))
    assert len(os.listdir(tmp_path / "out" / "txt")) == 4
    # and a real network in eval mode through the reference's Prediction (grad mode on, as evaluate.py runs it)
    torch.manual_seed(0)
    net = StackedHourglass(1, 128, 6).to(cuda_device).eval()
    x = torch.randn(2, 3, 128, 128, device=cuda_device)
    ref_pred = evaluate.Prediction(net, 20, 4, 0.0, "nms", 0.5).eval()
    ours = Prediction(net, 20, 4, 0.0, "nms", 0.5).eval()
    with torch.no_grad():
        rb, rc, rs = ref_pred(x)
    ob, oc, os_ = ours(x)
    for b in range(2):
        # a randomly initialised head has scores within 1e-3 of each other, so the two sigmoid implementations
        # (torch vs device expf, <= 1 ulp apart) may order near-ties differently: compare counts and the sorted scores
        assert rs[b].numel() == os_[b].numel() > 0 and rb[b].shape == ob[b].shape
        assert torch.allclose(rs[b], os_[b], atol=1e-6)
