"""Optimizer factory — drop-in for the reference's optim.py (`get_optimizer`, optim.py:3-12) with a fused Adam.

`FusedAdam` subclasses torch.optim.Adam (same constructor defaults, same `state_dict()` layout, so checkpoints written
by train.py:76-82 load in either direction) but `step()` is ONE kernel launch over all parameter tensors
(csrc/optim.cu) instead of a foreach sequence per tensor list; it implements GradScaler's
`_step_supports_amp_scaling` protocol (unscale + found_inf skip inside the kernel, train.py:128-132).
"""
from __future__ import annotations

import ctypes
from ctypes import c_float, c_int, c_longlong, c_void_p

import torch
import torch.optim as optim

from . import _lib
from ._lib import check, ptr, stream

_lib.register("hd_adam_step", c_int, [c_void_p, c_int, c_void_p, c_longlong, c_float, c_float, c_float, c_float,
                                      c_void_p, c_void_p, c_void_p, c_void_p])


class _Job(ctypes.Structure):
    _fields_ = [("p", c_void_p), ("g", c_void_p), ("m", c_void_p), ("v", c_void_p), ("n", c_longlong),
                ("chunk_start", c_longlong)]


class FusedAdam(optim.Adam):
    _step_supports_amp_scaling = True

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False)
        self._flat = {}        # group index -> (exp_avg flat, exp_avg_sq flat, step_dev, jobs_dev)

    def _init_group_state(self, gi, group):
        params = [p for p in group["params"]]
        dev = params[0].device
        total = sum(p.numel() for p in params)
        m = torch.zeros(total, dtype=torch.float32, device=dev)
        v = torch.zeros(total, dtype=torch.float32, device=dev)
        step_dev = torch.zeros((), dtype=torch.float32, device=dev)
        off = 0
        for p in params:
            st = self.state[p]
            n = p.numel()
            if "exp_avg" in st:                                   # restored by load_state_dict
                m[off:off + n].copy_(st["exp_avg"].reshape(-1))
                v[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                step_dev.fill_(float(st["step"]))
            st["exp_avg"] = m[off:off + n].view_as(p)
            st["exp_avg_sq"] = v[off:off + n].view_as(p)
            st["step"] = torch.tensor(float(step_dev))
            off += n
        jobs_dev = torch.empty(len(params) * ctypes.sizeof(_Job), dtype=torch.uint8, device=dev)
        self._flat[gi] = (m, v, step_dev, jobs_dev)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        grad_scale = getattr(self, "grad_scale", None)
        found_inf = getattr(self, "found_inf", None)
        for gi, group in enumerate(self.param_groups):
            if group.get("weight_decay", 0) != 0 or group.get("amsgrad", False) or group.get("maximize", False):
                raise NotImplementedError("FusedAdam: weight_decay / amsgrad / maximize are not on the B200 path "
                                          "(the reference uses plain Adam, optim.py:4)")
            if gi not in self._flat:
                self._init_group_state(gi, group)
            m, v, step_dev, jobs_dev = self._flat[gi]
            params = group["params"]
            jobs = (_Job * len(params))()
            off, chunk, used = 0, 0, 0
            for p in params:
                n = p.numel()
                if p.grad is not None:
                    _lib.require_cuda(p, "parameter")
                    g = p.grad
                    if g.dtype != torch.float32 or not g.is_contiguous():
                        g = g.float().contiguous()
                        p.grad = g
                    j = jobs[used]
                    j.p, j.g = p.data_ptr(), g.data_ptr()
                    j.m, j.v = m.data_ptr() + 4 * off, v.data_ptr() + 4 * off
                    j.n, j.chunk_start = n, chunk
                    chunk += (n + 1023) // 1024
                    used += 1
                off += n
            if used == 0:
                continue
            beta1, beta2 = group["betas"]
            with torch.cuda.device(m.device):
                check(_lib.lib().hd_adam_step(ctypes.cast(jobs, c_void_p), used, ptr(jobs_dev), chunk,
                                              float(group["lr"]), float(beta1), float(beta2), float(group["eps"]),
                                              ptr(step_dev), ptr(grad_scale) if grad_scale is not None else None,
                                              ptr(found_inf) if found_inf is not None else None, stream(m.device)),
                      "adam_step")
        return loss

    def state_dict(self):
        for gi, (m, v, step_dev, _) in self._flat.items():      # one D2H read, only when a checkpoint is written
            s = float(step_dev)
            for p in self.param_groups[gi]["params"]:
                if p in self.state:
                    self.state[p]["step"] = torch.tensor(s)
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._flat = {}                                          # re-flattened from the loaded tensors on next step


def get_optimizer(network, lr, lr_milestone, lr_gamma):
    optimizer = FusedAdam(network.parameters(), lr=lr)
    scheduler = None
    if lr_milestone is not None:
        scheduler = optim.lr_scheduler.MultiStepLR(optimizer=optimizer, milestones=lr_milestone, gamma=lr_gamma)
    return optimizer, scheduler
