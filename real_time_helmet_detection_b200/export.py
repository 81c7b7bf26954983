"""Deployment export — the role of the reference's export.py (TorchScript trace for PytorchToCpp/main.cpp).

The custom sm_100a kernels are not `torch.jit.trace`-able, so deployment does not go through TorchScript: the native
runner `runner/hd_infer.cpp` links `libhd_b200.so` directly and needs only the parameters, in the
executor's unit order (`StackedHourglass.units()`), as one flat file:

    "HDW1" | int32 num_stack, in_ch, out_ch, n_units |
    per unit: int32 cout, cin, k, has_bias, has_bn | fp32 weight OIHW | [fp32 bias] | [fp32 gamma, beta, mean, var]

`export_weights(network, path)` writes it (from a live module or after `load_state_dict` of a reference checkpoint),
`load_weights(network, path)` reads it back (round-trip check).
"""
from __future__ import annotations

import struct

import numpy as np
import torch
import torch.nn as nn

MAGIC = b"HDW1"


def _arrays(unit):
    conv, bn = unit.convolution, unit.bn
    arrs = [conv.weight]
    if conv.bias is not None:
        arrs.append(conv.bias)
    if isinstance(bn, nn.BatchNorm2d):
        arrs += [bn.weight, bn.bias, bn.running_mean, bn.running_var]
    return arrs


def export_weights(network, path):
    units = network.units()
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<4i", network.num_stack, network.in_ch, network.out_ch, len(units)))
        for u in units:
            w = u.convolution.weight
            f.write(struct.pack("<5i", w.shape[0], w.shape[1], w.shape[2], int(u.convolution.bias is not None),
                                int(isinstance(u.bn, nn.BatchNorm2d))))
            for a in _arrays(u):
                f.write(a.detach().to("cpu", torch.float32).contiguous().numpy().tobytes())
    return path


def load_weights(network, path):
    units = network.units()
    with open(path, "rb") as f:
        if f.read(4) != MAGIC:
            raise RuntimeError(f"{path}: not an HDW1 weight file")
        S, in_ch, out_ch, n = struct.unpack("<4i", f.read(16))
        if (S, in_ch, out_ch, n) != (network.num_stack, network.in_ch, network.out_ch, len(units)):
            raise RuntimeError(f"{path}: architecture mismatch (file: {S} stacks, {in_ch}/{out_ch} channels, {n} units)")
        for u in units:
            cout, cin, k, has_b, has_bn = struct.unpack("<5i", f.read(20))
            w = u.convolution.weight
            if tuple(w.shape) != (cout, cin, k, k) or bool(has_b) != (u.convolution.bias is not None) or \
                    bool(has_bn) != isinstance(u.bn, nn.BatchNorm2d):
                raise RuntimeError(f"{path}: unit shape mismatch")
            with torch.no_grad():
                for a in _arrays(u):
                    a.copy_(torch.from_numpy(np.frombuffer(f.read(a.numel() * 4), dtype=np.float32).copy()).view_as(a))
    return network
