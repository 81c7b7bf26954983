"""GPU parity of the fused loss kernels (csrc/loss.cu) against the oracle (oracle/loss_ref.py, fp32 CPU autograd)
and against the golden outputs of the unmodified reference (tests/golden/loss.npz).
Tolerance (north_star): 1e-3 relative on loss values; gradients: relative L2 <= 1e-4 (fp32 math on both sides)."""
import os
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _targets(B, size):
    from oracle.encode_ref import synthetic_targets
    return [torch.from_numpy(a) for a in synthetic_targets(B, imsize=size)]


def _rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


@pytest.mark.parametrize("normalized", [False, True])
@pytest.mark.parametrize("from_logits", [False, True])
def test_loss_vs_oracle(cuda_device, normalized, from_logits):
    from oracle import loss_ref
    from real_time_helmet_detection_b200.loss import LossCalculator
    B, size = 4, 256
    ghm, goff, gsize, gmask = _targets(B, size)
    logits = torch.randn(B, 6, size // 4, size // 4, generator=torch.Generator().manual_seed(5)) * 2
    lo = logits.clone().requires_grad_(True)
    hm, off, sz, tot = loss_ref.losses_from_logits(lo, ghm, goff, gsize, gmask, normalized_coord=normalized)
    (tot * 3.0).backward()

    crit = LossCalculator(1.0, 1.0, 0.1, 2.0, 4.0).to(cuda_device)
    ld = logits.to(cuda_device).requires_grad_(True)
    g = [t.to(cuda_device) for t in (ghm, goff, gsize, gmask)]
    if from_logits:
        total = crit.forward_logits(ld, *g, normalized_coord=normalized)
    else:
        phm, poff, psz = ld.split([2, 2, 2], dim=1)
        phm = torch.sigmoid(phm)
        if normalized:
            poff, psz = torch.sigmoid(poff), torch.sigmoid(psz)
        total = crit(phm, poff, psz, *g)
    (total * 3.0).backward()
    vals = [crit.log[k][-1] for k in ("hm", "offset", "size", "total")]
    for v, r in zip(vals, (hm, off, sz, tot)):
        assert abs(v - r.item()) <= 1e-3 * abs(r.item()) + 1e-6
    assert abs(total.item() - tot.item()) <= 1e-4 * abs(tot.item())
    assert _rel(ld.grad.cpu(), lo.grad) <= 1e-4


def test_loss_vs_reference_golden(cuda_device):
    from real_time_helmet_detection_b200.loss import LossCalculator
    gold = np.load(os.path.join(GOLD, "loss.npz"))
    logits = torch.from_numpy(gold["logits"])
    import sys
    sys.path.insert(0, GOLD)
    from make_golden import FIXED_BOXES
    from oracle.encode_ref import encode_boxes
    B, h = logits.shape[0], logits.shape[2]
    outs = [[], [], [], []]
    for b in range(B):
        boxes, labels = FIXED_BOXES[b % len(FIXED_BOXES)]
        boxes = [[v * 0.5 for v in bx] for bx in boxes]
        for lst, arr in zip(outs, encode_boxes(boxes, labels, (128, 128))):
            lst.append(arr)
    gts = {"boxes": [torch.from_numpy(np.stack(o)) for o in outs],
           "nopos": [torch.zeros(B, 2, h, h), torch.zeros(B, 2, h, h), torch.zeros(B, 2, h, h), torch.zeros(B, 1, h, h)]}
    for name, gt in gts.items():
        for norm in (False, True):
            key = f"{name}_{'norm' if norm else 'lin'}"
            crit = LossCalculator(1.0, 1.0, 0.1, 2.0, 4.0)
            ld = logits.to(cuda_device).requires_grad_(True)
            total = crit.forward_logits(ld, *[t.to(cuda_device) for t in gt], normalized_coord=norm)
            total.backward()
            vals = np.asarray([crit.log[k][-1] for k in ("hm", "offset", "size", "total")])
            ref = gold[key + "_values"]
            assert np.all(np.abs(vals - ref) <= 1e-3 * np.abs(ref) + 1e-6), (key, vals, ref)
            assert _rel(ld.grad.cpu(), torch.from_numpy(gold[key + "_dlogits"])) <= 1e-4, key


def test_get_log_and_checkpoint_roundtrip(cuda_device):
    from real_time_helmet_detection_b200.loss import LossCalculator
    gold = np.load(os.path.join(GOLD, "loss.npz"))
    crit = LossCalculator(1.0, 1.0, 0.1, 2.0, 4.0)
    crit.log = {k: [0.5 * i for i in range(150)] for k in ("hm", "offset", "size", "total")}
    assert crit.get_log() == str(gold["get_log"])
    assert set(crit.log) == {"hm", "offset", "size", "total"} and isinstance(crit.log["hm"][0], float)


def test_loss_amp_gradscaler_inputs(cuda_device):
    """Half / bf16 predictions (ambient autocast) and a 65536x upstream gradient must work."""
    from real_time_helmet_detection_b200.loss import LossCalculator
    ghm, goff, gsize, gmask = [t.to(cuda_device) for t in _targets(2, 128)]
    logits = (torch.randn(2, 6, 32, 32, device=cuda_device)).requires_grad_(True)
    crit = LossCalculator(1.0, 1.0, 0.1, 2.0, 4.0)
    total = crit.forward_logits(logits.half(), ghm, goff, gsize, gmask)
    (total * 65536.0).backward()
    assert torch.isfinite(logits.grad).all() and logits.grad.abs().sum() > 0
