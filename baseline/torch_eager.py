"""Stock PyTorch / cuDNN execution of the stacked hourglass - the "library bar" (SURVEY.md 2.1, 8d last row).

Measurement / test infrastructure, not product code (the package never imports baseline/). Two providers of the same
thing, "the reference's network run by the libraries the reference itself calls" (nn.Conv2d -> cuDNN, nn.BatchNorm2d ->
cuDNN/ATen, MaxPool2d, Upsample, autograd; hourglass.py:94-237):

  * `reference_network(...)`: the UNMODIFIED reference class from baseline/_ref/hourglass.py when it is staged
    (tools/stage_reference.py) - kind "reference";
  * `EagerHourglass`: when it is not (e.g. a checkout without the staging step), the same computation expressed with
    torch.nn.functional calls over the parameter containers of real_time_helmet_detection_b200.hourglass (whose module
    tree, state_dict and initial values are the reference's, tests/test_oracle_golden.py) - kind "port".

Both take / return the reference's tensors: x (B,3,H,W) -> (B, S, num_cls+4, H/4, W/4) raw logits, NCHW or channels_last.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class EagerHourglass(nn.Module):
    """Runs a `real_time_helmet_detection_b200.hourglass.StackedHourglass` (a tree of parameter containers) with stock
    PyTorch ops. The wrapped module's parameters ARE this module's parameters."""

    def __init__(self, containers):
        super().__init__()
        self.net = containers

    @staticmethod
    def _conv(u, x, relu):
        c = u.convolution
        y = F.conv2d(x, c.weight, c.bias, stride=c.stride, padding=c.padding)
        if isinstance(u.bn, nn.BatchNorm2d):
            y = u.bn(y)                       # a real nn.BatchNorm2d: train / eval mode and running stats as usual
        return F.relu(y) if relu else y

    def _res(self, r, x):
        y = self._conv(r.conv2, self._conv(r.conv1, x, True), False)
        s = self._conv(r.skip, x, False) if hasattr(r.skip, "convolution") else x
        return F.relu(y + s)

    def _hg(self, h, x):
        up1 = self._res(h.up1, x)
        low = self._res(h.low1, F.max_pool2d(x, 2, 2))
        low = self._hg(h.low2, low) if hasattr(h.low2, "up1") else self._res(h.low2, low)
        low = self._res(h.low3, low)
        return up1 + F.interpolate(low, scale_factor=2, mode="nearest")

    def forward(self, x):
        n = self.net
        pl = n.pre_layer.layers
        x = self._conv(pl[0], x, True)
        x = self._res(pl[1], x)
        x = F.max_pool2d(x, 2, 2)
        x = self._res(pl[4], self._res(pl[3], x))
        outs = []
        for i in range(n.num_stack):
            f = self._hg(n.hourglass_lst[i], x)
            f = self._res(n.neck_lst[i].layers[2], self._conv(n.neck_lst[i].layers[1], f, True))
            p = self._conv(n.head_lst[i].layer, f, False)
            outs.append(p)
            if i < n.num_stack - 1:
                x = x + self._conv(n.merge_feature[i], f, False) + self._conv(n.merge_prediction[i], p, False)
        return torch.stack(outs, dim=1)


def reference_network(num_stack, state_dict=None, device="cpu", seed=777):
    """(module, kind): the staged reference class if available, else EagerHourglass over our containers; same weights
    either way (seeded default init, or `state_dict`)."""
    from . import refload
    if refload.available():
        ref = refload.load_reference(("hourglass",))["hourglass"]
        torch.manual_seed(seed)
        net = ref.StackedHourglass(num_stack=num_stack, in_ch=128, out_ch=6)
        kind = "reference"
    else:
        from real_time_helmet_detection_b200.hourglass import StackedHourglass
        torch.manual_seed(seed)
        net = EagerHourglass(StackedHourglass(num_stack, 128, 6))
        kind = "port"
    if state_dict is not None:
        (net.net if kind == "port" else net).load_state_dict(state_dict)
    return net.to(device), kind


def reference_loss(device="cpu"):
    """The reference's LossCalculator (loss.py:6-40) if staged, else None (callers fall back to the oracle port)."""
    from . import refload
    if not refload.available():
        return None
    return refload.load_reference(("loss",))["loss"].LossCalculator(1.0, 1.0, 0.1, 2.0, 4.0).to(device)


def train_loop_body(net, crit, image, gts, autocast_dtype=None, num_cls=2):
    """train.py:97-134 with the out-of-place squeeze: forward, per-stack split + sigmoid + loss, backward. Returns
    (outputs, total_loss). `crit` None -> the oracle port of the loss (only when the reference is not staged)."""
    device_type = image.device.type
    with torch.autocast(device_type, dtype=autocast_dtype, enabled=autocast_dtype is not None):
        outputs = net(image)
        total = 0
        for output in outputs.split(1, dim=1):
            output = output.squeeze(1)
            if crit is not None:
                phm, poff, psz = output.split([num_cls, 2, 2], dim=1)
                total = total + crit(torch.sigmoid(phm), poff, psz, *gts)
            else:
                from oracle import loss_ref
                total = total + loss_ref.losses_from_logits(output.float(), *gts)[3]
    total.backward()
    return outputs, total
