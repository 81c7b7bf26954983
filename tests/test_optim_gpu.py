"""GPU parity of the fused multi-tensor Adam (csrc/optim.cu) against torch.optim.Adam (the optimizer the reference
builds, optim.py:4) on identical parameters / gradients, including the GradScaler protocol and checkpoint round trip."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _params(dev, seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [(128, 128, 3, 3), (128,), (6, 128, 1, 1), (64, 3, 7, 7), (1,), (1500,)]
    return [torch.nn.Parameter(torch.randn(s, generator=g).to(dev)) for s in shapes]


def test_fused_adam_matches_torch_adam(cuda_device):
    from real_time_helmet_detection_b200.optim import FusedAdam
    pa, pb = _params(cuda_device, 0), _params(cuda_device, 0)
    ref = torch.optim.Adam(pa, lr=5e-4)
    mine = FusedAdam(pb, lr=5e-4)
    g = torch.Generator().manual_seed(1)
    for it in range(5):
        for a, b in zip(pa, pb):
            gr = torch.randn(a.shape, generator=g).to(cuda_device)
            a.grad, b.grad = gr.clone(), gr.clone()
        ref.step()
        mine.step()
    for a, b in zip(pa, pb):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-7), (a - b).abs().max()
    # checkpoint round trip in both directions (train.py:76-82, :195)
    sd = mine.state_dict()
    assert float(sd["state"][0]["step"]) == 5.0
    ref2 = torch.optim.Adam(_params(cuda_device, 0), lr=5e-4)
    ref2.load_state_dict(sd)
    mine2 = FusedAdam(_params(cuda_device, 0), lr=5e-4)
    mine2.load_state_dict(ref.state_dict())
    for p, q in zip(mine2.param_groups[0]["params"], pa):
        p.data.copy_(q.data)
        p.grad = torch.ones_like(p)
        q.grad = torch.ones_like(q)
    mine2.step()
    ref.step()
    for p, q in zip(mine2.param_groups[0]["params"], pa):
        assert torch.allclose(p, q, rtol=1e-5, atol=1e-7)


def test_fused_adam_with_gradscaler(cuda_device):
    from real_time_helmet_detection_b200.optim import FusedAdam
    p = torch.nn.Parameter(torch.ones(1000, device=cuda_device))
    q = torch.nn.Parameter(torch.ones(1000, device=cuda_device))
    mine, ref = FusedAdam([p], lr=1e-2), torch.optim.Adam([q], lr=1e-2)
    s1, s2 = torch.amp.GradScaler("cuda", init_scale=1024.0), torch.amp.GradScaler("cuda", init_scale=1024.0)
    for it in range(3):
        for par, opt, sc in ((p, mine, s1), (q, ref, s2)):
            opt.zero_grad()
            loss = (par * par).sum() * (float("inf") if it == 1 else 1.0)    # second step overflows -> skipped
            sc.scale(loss).backward()
            sc.step(opt)
            sc.update()
    assert torch.allclose(p, q, rtol=1e-5, atol=1e-7)
    assert float(mine.state_dict()["state"][0]["step"]) == 2.0               # the inf step did not count


def test_get_optimizer_factory(cuda_device):
    from real_time_helmet_detection_b200.optim import get_optimizer, FusedAdam
    net = torch.nn.Linear(4, 4).to(cuda_device)
    opt, sched = get_optimizer(net, 5e-4, [50, 90], 0.1)
    assert isinstance(opt, FusedAdam) and isinstance(sched, torch.optim.lr_scheduler.MultiStepLR)
    assert get_optimizer(net, 5e-4, None, 0.1)[1] is None
