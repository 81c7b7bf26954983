"""Oracle parity at BASELINE.json's resolution (512x512; configs 2 and 3 at a batch the CPU oracle finishes in seconds) and
the statement of what "1e-3 on heat-maps" means for a bf16 network.

Three implementations of the same train step (train.py:99-134: forward, per-stack loss, backward) on the same seeded
weights / inputs / targets:

  fp32   the oracle (oracle/hourglass_ref.py + loss_ref.py, CPU fp32) - pinned bit-for-bit-class to the unmodified reference
         by tests/test_oracle_golden.py; this is the golden result;
  lib16  the UNMODIFIED reference modules (baseline/_ref/hourglass.py + loss.py; the torch.nn.functional port of
         baseline/torch_eager.py when the reference is not staged) on the SAME GPU under bf16 autocast - what BASELINE
         configs 2-4 ("bf16") mean for the reference: cuDNN bf16 convolutions, bf16 activations and BN I/O;
  ours   the B200 path (tcgen05 bf16 operands, fp32 accumulation, fp32 BN statistics, bf16 stored activations).

A chain of ~100 bf16 rounding points cannot reproduce fp32 logits to 1e-3 in ANY implementation (lib16 is ~3e-2 from fp32);
the contract checked here, per tensor, is therefore: the CUDA path is at least as close to the fp32 golden as the reference
itself is when it runs in bf16 on this GPU - for the logits, the loss, each of the 115 / 209 parameter gradients and
every BN running statistic - with the aggregate (loss) also held to 1e-3 where bf16 allows. Every 256x256-level kernel variant
(64-channel convs, two-BN residual tail fused with the pool, index pool backward) is on this path: the gradients of
`pre_layer.layers.0/1.*` are asserted by name. Measured values go to gpurun_out/parity512_S{S}_B{B}.json (copied into
profiles/ for the record).
"""
import json
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _oracle_fp32(sd0, x, gts, S):
    from oracle import hourglass_ref, loss_ref
    sd = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone())
          for k, v in sd0.items()}
    stats = {}
    out = hourglass_ref.stacked_hourglass_forward(sd, x, training=True, new_stats=stats)
    tot = sum(loss_ref.losses_from_logits(out[:, s], *gts)[3] for s in range(S))
    tot.backward()
    return out.detach(), float(tot), {k: v.grad for k, v in sd.items() if v.requires_grad}, stats


@pytest.mark.parametrize("S,B,knobs", [(1, 2, {}), (2, 2, {}), (1, 4, {}), (1, 4, {"HD_DGRAD_BNSTAT": "1"})])
def test_train_step_512_vs_oracle_and_bf16_reference(cuda_device, monkeypatch, S, B, knobs):
    # knobs: opt-in executor variants that must meet the same contract (HD_DGRAD_BNSTAT: conv2's dgrad reduces the
    # statistics of conv1's BN backward in its epilogue - at B = 4 the 256x256 and 128x128 levels take that path)
    for k_, v_ in knobs.items():
        monkeypatch.setenv(k_, v_)
    from baseline import torch_eager
    from real_time_helmet_detection_b200.hourglass import StackedHourglass
    from real_time_helmet_detection_b200.loss import LossCalculator
    from real_time_helmet_detection_b200.synthetic import synthetic_targets
    size = 512
    torch.manual_seed(777)
    net = StackedHourglass(S, 128, 6)
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    x = torch.randn(B, 3, size, size, generator=torch.Generator().manual_seed(S * 10 + B))
    gts = [torch.from_numpy(a) for a in synthetic_targets(B, imsize=size)]
    names = [n for n, _ in net.named_parameters()]
    assert len(sd0) == (226 if S == 1 else 407)

    # ---- golden: fp32 oracle on the host
    o32, l32, g32, st32 = _oracle_fp32(sd0, x, gts, S)

    # ---- the reference itself in bf16 on this GPU
    lib, kind = torch_eager.reference_network(S, sd0, cuda_device)
    lib.train()
    crit_ref = torch_eager.reference_loss(cuda_device)
    gts_d = [g.to(cuda_device) for g in gts]
    out_l, tot_l = torch_eager.train_loop_body(lib, crit_ref, x.to(cuda_device), gts_d, autocast_dtype=torch.bfloat16)
    lib_params = dict((lib.net if kind == "port" else lib).named_parameters())
    g16 = {n: lib_params[n].grad.detach().float().cpu() for n in names}
    st16 = {k: v.detach().cpu() for k, v in (lib.net if kind == "port" else lib).state_dict().items() if "running_" in k}
    o16, l16 = out_l.detach().float().cpu(), float(tot_l)
    del lib, out_l, tot_l
    torch.cuda.empty_cache()

    # ---- ours
    net = net.to(cuda_device).train()
    crit = LossCalculator(1.0, 1.0, 0.1, 2.0, 4.0)
    out = net(x.to(cuda_device))
    total = sum(crit.forward_logits(out[:, s], *gts_d) for s in range(S))
    total.backward()
    o, l = out.detach().cpu(), float(total)
    g = {n: p.grad.detach().cpu() for n, p in net.named_parameters()}
    st = {k: v.detach().cpu() for k, v in net.state_dict().items() if "running_" in k}

    # ---- measurements first (kept even when an assertion below fails)
    gmax = max(v.norm().item() for v in g32.values())
    # "live" = every tensor whose fp32 gradient is not analytically zero (the zero ones - conv biases in front of a
    # train-mode BN and the like - sit at <= 1e-7 of the largest gradient norm in fp32; see the end of the test)
    live = [n for n in names if g32[n].norm().item() >= 1e-5 * gmax]
    e_ours = {n: rel(g[n], g32[n]) for n in live}
    e_lib = {n: rel(g16[n], g32[n]) for n in live}
    ratios = np.array([e_ours[n] / max(e_lib[n], 1e-12) for n in live])
    flat = lambda d: torch.cat([d[n].flatten().double() for n in names])     # noqa: E731
    f32, fo, fl = flat(g32), flat(g), flat(g16)
    cos_o = float(torch.dot(fo, f32) / (fo.norm() * f32.norm()))
    cos_l = float(torch.dot(fl, f32) / (fl.norm() * f32.norm()))
    stat_o = {k: float((st[k] - st32[k]).abs().max() / (st32[k].abs().max() + 1e-3)) for k in st32}
    stat_l = {k: float((st16[k] - st32[k]).abs().max() / (st32[k].abs().max() + 1e-3)) for k in st32}
    rec = {"S": S, "B": B, "size": size, "comparator": kind,
           "logits_rel_l2": {"ours": rel(o, o32), "lib_bf16": rel(o16, o32)},
           "heatmap_logits_rel_l2": {"ours": rel(o[:, :, :2], o32[:, :, :2]), "lib_bf16": rel(o16[:, :, :2], o32[:, :, :2])},
           "loss": {"fp32": l32, "ours": l, "lib_bf16": l16, "ours_rel": abs(l - l32) / abs(l32),
                    "lib_rel": abs(l16 - l32) / abs(l32)},
           "grad_flat_rel_l2": {"ours": rel(fo, f32), "lib_bf16": rel(fl, f32)}, "grad_cosine": {"ours": cos_o, "lib_bf16": cos_l},
           "grad_tensors_checked": len(live), "grad_ratio_ours_over_lib": {
               "median": float(np.median(ratios)), "p90": float(np.quantile(ratios, 0.9)), "max": float(ratios.max()),
               "worst": live[int(ratios.argmax())]},
           "grad_rel_l2_median": {"ours": float(np.median(list(e_ours.values()))), "lib_bf16": float(np.median(list(e_lib.values())))},
           "level256_grads": {n: {"ours": e_ours[n], "lib_bf16": e_lib[n]} for n in live if n.startswith(("pre_layer.layers.0", "pre_layer.layers.1"))},
           "running_stats_max_err_over_range": {"ours": max(stat_o.values()), "lib_bf16": max(stat_l.values())}}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    tag = "".join("_" + k_.lower() for k_ in knobs)
    with open(os.path.join(ROOT, "gpurun_out", f"parity512_S{S}_B{B}{tag}.json"), "w") as f:
        json.dump(rec, f, indent=1)
    print(json.dumps(rec))

    # ---- the contract
    assert o.shape == o32.shape == (B, S, 6, 128, 128) and torch.isfinite(o).all()
    # logits / heat-maps: at least as close to fp32 as the reference in bf16 (10 % slack for run-to-run atomics order)
    assert rec["logits_rel_l2"]["ours"] <= 1.1 * rec["logits_rel_l2"]["lib_bf16"] + 1e-3, rec["logits_rel_l2"]
    assert rec["heatmap_logits_rel_l2"]["ours"] <= 1.1 * rec["heatmap_logits_rel_l2"]["lib_bf16"] + 1e-3
    # loss: 1e-3 of the fp32 value, or as close as the bf16 reference gets
    assert abs(l - l32) <= max(1e-3 * abs(l32), 1.5 * abs(l16 - l32)), rec["loss"]
    # every live parameter gradient (incl. the 256x256 level by name), per tensor and in aggregate
    assert len(live) >= len(names) - 4 * S and len(names) == (115 if S == 1 else 209)
    for n in live:
        assert torch.isfinite(g[n]).all() and g[n].shape == g32[n].shape, n
        assert e_ours[n] <= 1.5 * e_lib[n] + 2e-2, (n, e_ours[n], e_lib[n])
    for n in ("pre_layer.layers.0.convolution.weight", "pre_layer.layers.1.conv1.convolution.weight",
              "pre_layer.layers.1.conv2.convolution.weight", "pre_layer.layers.1.skip.convolution.weight",
              "pre_layer.layers.1.skip.bn.weight", "pre_layer.layers.1.conv2.bn.bias"):
        assert n in live, n
    assert np.median(ratios) <= 1.05, float(np.median(ratios))
    assert rec["grad_flat_rel_l2"]["ours"] <= 1.1 * rec["grad_flat_rel_l2"]["lib_bf16"] + 1e-3
    assert cos_o >= cos_l - 1e-3
    # the zero-gradient biases (stem / neck conv bias in front of a train-mode BN, quirk Q12): the fp32 reference returns
    # ~1e-7 of the weight gradient; ours returns exactly 0, the bf16 reference returns rounding noise
    for n in names:
        if n not in live:
            assert g[n].norm().item() <= 1e-3 * gmax, (n, g[n].norm().item(), gmax)
    for n in ("pre_layer.layers.0.convolution.bias", "neck_lst.0.layers.1.convolution.bias"):
        assert n not in live and g[n].abs().max().item() == 0.0 and g32[n].norm().item() <= 1e-6 * gmax, n
    # BN running statistics after one step
    assert max(stat_o.values()) <= max(stat_l.values()), (max(stat_o.values()), max(stat_l.values()))
    for k in st32:
        assert stat_o[k] <= max(2.5 * stat_l[k], 5e-3), (k, stat_o[k], stat_l[k])
