"""Stacked-hourglass network — drop-in for the reference's hourglass.py (`StackedHourglass`, hourglass.py:198-237).

Same constructor signature, same `forward(x) -> (B, num_stack, num_cls+4, H/4, W/4)` raw-logit contract, and the SAME
module tree / parameter construction order, so that
  * `state_dict()` has the reference's 226 (1 stack) / 407 (2 stacks) keys, shapes and dtypes and released checkpoints
    load with `load_state_dict`;
  * `torch.manual_seed(s); StackedHourglass(...)` yields bit-identical initial weights to the reference
    (the parameter containers are the same torch.nn classes, created in the same order);
  * `optim.Adam(network.parameters())`, `DistributedDataParallel(network)`, `.train()/.eval()`, `.to(device)` work.

What differs is execution: the sub-modules are parameter containers only. `StackedHourglass.forward` hands the whole
network to the native executor (csrc/net.cu) which runs hand-written sm_100a kernels (tcgen05 implicit-GEMM
convolutions, fused BN / ReLU / residual / pool / upsample kernels) over one pre-planned HBM arena; the backward pass
is a single autograd node that runs the hand-written backward schedule and returns every parameter gradient as a
view of one flat fp32 buffer (which is also what the flat NCCL all-reduce of parallel.py operates on).
Internal activation dtype is bf16 (fp32 accumulation, fp32 BN statistics) regardless of the ambient autocast dtype;
the returned logits are fp32. There is no CPU path: forward() on a CPU tensor raises.

Only the default architecture flags of the reference are implemented (activation/neck_activation 'ReLU', pool 'Max',
neck_pool 'None', increase_ch 0, in_ch 128) — the configurations of every BASELINE workload. Other valid flag values
raise NotImplementedError, unknown strings raise the reference's "Not expected ..." NotImplementedError.
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, Structure, byref, c_float, c_longlong, c_void_p, c_int, c_size_t

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, ptr, stream

_ACTIVATIONS = ('ReLU', 'LReLU', 'PReLU', 'Linear', 'Mish', 'Sigmoid', 'CELU')
_POOLS = ('Max', 'Avg', 'Conv', 'SPP', 'None')


class UnitPtrs(Structure):
    """ctypes mirror of `hd_unit_ptrs` (include/hd_b200.h)."""
    _fields_ = [("w", c_void_p), ("b", c_void_p), ("gamma", c_void_p), ("beta", c_void_p),
                ("running_mean", c_void_p), ("running_var", c_void_p), ("num_batches_tracked", c_void_p),
                ("dw", c_void_p), ("db", c_void_p), ("dgamma", c_void_p), ("dbeta", c_void_p)]


_lib.register("hd_net_create", c_int, [c_int, c_int, c_int, POINTER(c_void_p)])
_lib.register("hd_net_destroy", None, [c_void_p])
_lib.register("hd_net_num_units", c_int, [c_void_p])
_lib.register("hd_net_set_static_weights", None, [c_void_p, c_int])
_lib.register("hd_net_workspace_bytes", c_size_t, [c_void_p, c_int, c_int, c_int, c_int])
_lib.register("hd_net_forward", c_int, [c_void_p, POINTER(UnitPtrs), c_int, c_void_p, c_void_p, c_void_p, c_size_t,
                                         c_int, c_int, c_int, c_int, c_void_p])
_lib.register("hd_net_backward", c_int, [c_void_p, POINTER(UnitPtrs), c_int, c_void_p, c_void_p, c_size_t, c_void_p])
_lib.register("hd_net_backward_stage", c_int, [c_void_p, POINTER(UnitPtrs), c_int, c_void_p, c_void_p, c_size_t, c_void_p,
                                               c_int, c_void_p])


def _container_forward(self, *args, **kwargs):
    raise RuntimeError(f"{type(self).__name__} is a parameter container: the B200 path executes the whole "
                       "StackedHourglass through the fused network executor (call the StackedHourglass module)")


class Activation(nn.Module):
    def __init__(self, activation: str):
        super().__init__()
        if activation not in _ACTIVATIONS:
            raise NotImplementedError("Not expected activation: %s" % activation)
        if activation not in ('ReLU', 'Linear'):
            raise NotImplementedError("activation '%s' has no sm_100a kernel: the B200 hot path covers the reference's "
                                      "default 'ReLU' (and the internal 'Linear')" % activation)
        self.kind = activation
        self.activation = nn.ReLU() if activation == 'ReLU' else nn.Identity()

    forward = _container_forward


class Pool(nn.Module):
    def __init__(self, channel: int, pool: str):
        super().__init__()
        if pool not in _POOLS:
            raise NotImplementedError("Not expected pool: %s" % pool)
        if pool not in ('Max', 'None'):
            raise NotImplementedError("pool '%s' has no sm_100a kernel: the B200 hot path covers the reference's "
                                      "default 'Max' (hourglass) and 'None' (neck)" % pool)
        self.kind = pool
        self.pool = nn.MaxPool2d(2, 2) if pool == 'Max' else nn.Identity()

    forward = _container_forward


class Convolution(nn.Module):
    """Parameter container of hourglass.py:94-108: `convolution` (nn.Conv2d) and `bn` (nn.BatchNorm2d | Identity)."""

    def __init__(self, in_ch, out_ch, kernel_size=3, stride=1, bias=True, bn=False, activation='ReLU'):
        super().__init__()
        self.activation = Activation(activation)
        self.convolution = nn.Conv2d(in_ch, out_ch, kernel_size, stride, padding=(kernel_size - 1) // 2, bias=bias)
        self.bn = nn.BatchNorm2d(out_ch, affine=True, track_running_stats=True) if bn else nn.Identity()

    forward = _container_forward


class Residual(nn.Module):
    def __init__(self, in_ch, out_ch, kernel_size=3, stride=1, activation='ReLU'):
        super().__init__()
        self.activation = Activation(activation)
        self.conv1 = Convolution(in_ch, out_ch, kernel_size, stride, bias=False, bn=True, activation=activation)
        self.conv2 = Convolution(out_ch, out_ch, kernel_size, stride, bias=False, bn=True, activation='Linear')
        if in_ch != out_ch:
            self.skip = Convolution(in_ch, out_ch, kernel_size=1, stride=stride, bias=False, bn=True,
                                    activation='Linear')
        else:
            self.skip = nn.Identity()

    forward = _container_forward

    def units(self):
        return [self.conv1, self.conv2] + ([self.skip] if isinstance(self.skip, Convolution) else [])


class Hourglass(nn.Module):
    def __init__(self, num_layer, in_ch, increase_ch=0, activation='ReLU', pool='Max'):
        super().__init__()
        mid_ch = in_ch + increase_ch
        self.up1 = Residual(in_ch, in_ch, activation=activation)
        self.pool1 = Pool(in_ch, pool=pool)
        self.low1 = Residual(in_ch, mid_ch, activation=activation)
        if num_layer > 1:
            self.low2 = Hourglass(num_layer - 1, mid_ch, increase_ch, activation=activation, pool=pool)
        else:
            self.low2 = Residual(mid_ch, mid_ch, activation=activation)
        self.low3 = Residual(mid_ch, in_ch, activation=activation)
        self.up2 = nn.Upsample(scale_factor=2, mode='nearest')

    forward = _container_forward

    def units(self):
        return self.up1.units() + self.low1.units() + self.low2.units() + self.low3.units()


class PreLayer(nn.Module):
    def __init__(self, in_ch=3, mid_ch=128, out_ch=5, activation='ReLU', pool='Max'):
        super().__init__()
        self.layers = nn.Sequential(
            Convolution(in_ch=in_ch, out_ch=64, kernel_size=7, stride=2, bias=True, bn=True, activation=activation),
            Residual(in_ch=64, out_ch=mid_ch),
            Pool(channel=mid_ch, pool=pool),
            Residual(in_ch=mid_ch, out_ch=mid_ch),
            Residual(in_ch=mid_ch, out_ch=out_ch))

    forward = _container_forward

    def units(self):
        return [self.layers[0]] + self.layers[1].units() + self.layers[3].units() + self.layers[4].units()


class Neck(nn.Module):
    def __init__(self, ch=128, activation='ReLU', pool='None'):
        super().__init__()
        if pool not in _POOLS:
            raise NotImplementedError("Not expected pool: %s" % pool)
        if pool != 'None':
            raise NotImplementedError("neck_pool '%s' is not on the B200 hot path (reference default: 'None')" % pool)
        self.layers = nn.Sequential(
            Pool(ch, pool),
            Convolution(in_ch=ch, out_ch=ch, kernel_size=1, bn=True, activation=activation),
            Residual(ch, ch))

    forward = _container_forward

    def units(self):
        return [self.layers[1]] + self.layers[2].units()


class Head(nn.Module):
    def __init__(self, in_ch, out_ch, kernel_size=1, stride=1, bias=True, bn=False, activation='Linear'):
        super().__init__()
        self.layer = Convolution(in_ch=in_ch, out_ch=out_ch, kernel_size=kernel_size, stride=stride, bias=bias, bn=bn,
                                 activation=activation)

    forward = _container_forward


class _HourglassFn(torch.autograd.Function):
    """One autograd node for the whole network."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, module, x, *params):
        logits = module._run_forward(x)
        ctx.module = module
        ctx.generation = module._generation
        ctx.bn_train = module.training
        return logits

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dlogits):
        module = ctx.module
        if not ctx.bn_train:
            raise RuntimeError("StackedHourglass: backward through an eval()-mode forward is not supported "
                               "(BatchNorm uses batch statistics on the training path)")
        if ctx.generation != module._generation:
            raise RuntimeError("StackedHourglass: a newer forward pass overwrote the activations of this graph; "
                               "call backward() before running the network again")
        grads = module._run_backward(dlogits)
        return (None, None, *grads)


class StackedHourglass(nn.Module):
    def __init__(self, num_stack: int, in_ch: int, out_ch: int, increase_ch: int = 0, activation: str = 'ReLU',
                 pool: str = 'Max', neck_activation: str = 'ReLU', neck_pool: str = 'None'):
        super().__init__()
        if activation not in _ACTIVATIONS:
            raise NotImplementedError("Not expected activation: %s" % activation)
        if neck_activation not in _ACTIVATIONS:
            raise NotImplementedError("Not expected activation: %s" % neck_activation)
        if activation != 'ReLU' or neck_activation != 'ReLU':
            raise NotImplementedError("only activation='ReLU' / neck_activation='ReLU' are on the B200 hot path")
        if increase_ch != 0:
            raise NotImplementedError("increase_ch != 0 is not on the B200 hot path (reference default: 0)")
        if in_ch != 128:
            raise NotImplementedError("hourglass_inch != 128 is not on the B200 hot path (reference default: 128)")
        # same construction order as the reference => same RNG stream => same initial weights
        self.pre_layer = PreLayer(in_ch=3, mid_ch=128, out_ch=in_ch, activation=activation, pool=pool)
        self.hourglass_lst = nn.ModuleList([Hourglass(num_layer=4, in_ch=in_ch, increase_ch=increase_ch,
                                                      activation=activation, pool=pool) for _ in range(num_stack)])
        self.neck_lst = nn.ModuleList([Neck(in_ch, neck_activation, neck_pool) for _ in range(num_stack)])
        self.head_lst = nn.ModuleList([Head(in_ch=in_ch, out_ch=out_ch, kernel_size=1, stride=1, bias=True, bn=False,
                                            activation='Linear') for _ in range(num_stack)])
        self.merge_feature = nn.ModuleList([Convolution(in_ch=in_ch, out_ch=in_ch, kernel_size=1, stride=1, bias=True,
                                                        bn=False, activation='Linear') for _ in range(num_stack - 1)])
        self.merge_prediction = nn.ModuleList([Convolution(in_ch=out_ch, out_ch=in_ch, kernel_size=1, stride=1,
                                                           bias=True, bn=False, activation='Linear')
                                               for _ in range(num_stack - 1)])
        self.num_stack = num_stack
        self.in_ch = in_ch
        self.out_ch = out_ch
        # native state (not part of state_dict)
        self._handle = None
        self._workspace = None
        self._ws_key = None
        self._table = None
        self._table_key = None
        self._generation = 0
        self._flat_grad = None
        # optional gradient exchange run inside backward (parallel.FlatAllReduce): either a plain callable(flat_grad)
        # invoked after the whole pass, or an object with early(bucket) / late(bucket) for the two-bucket overlap
        self.grad_sync = None

    # ------------------------------------------------------------------ unit table
    def units(self):
        """`Convolution` modules in the executor's unit order (csrc/net.cu header comment)."""
        us = self.pre_layer.units()
        for i in range(self.num_stack):
            us += self.hourglass_lst[i].units() + self.neck_lst[i].units() + [self.head_lst[i].layer]
            if i < self.num_stack - 1:
                us += [self.merge_feature[i], self.merge_prediction[i]]
        return us

    def _param_list(self):
        return list(self.parameters())

    def _native(self):
        if self._handle is None:
            h = c_void_p()
            check(_lib.lib().hd_net_create(self.num_stack, self.in_ch, self.out_ch, byref(h)), "net_create")
            self._handle = h
            n = _lib.lib().hd_net_num_units(h)
            if n != len(self.units()):
                raise RuntimeError(f"unit table mismatch: native {n} vs python {len(self.units())}")
        return self._handle

    def freeze_weights(self, on=True):
        """Inference with frozen parameters: promise that parameters and BN buffers will not change, so eval-mode
        forwards reuse the packed bf16 weights / folded BN constants kept in the workspace instead of rebuilding them on
        every call (two launches, ~40 us of a 0.5 ms batch-1 pass). Call again (or `freeze_weights(False)`) after
        modifying the weights; training-mode forwards always repack."""
        self._frozen = bool(on)
        _lib.lib().hd_net_set_static_weights(self._native(), 1 if on else 0)
        return self

    def __getstate__(self):
        state = self.__dict__.copy()
        for k in ("_handle", "_workspace", "_ws_key", "_table", "_table_key", "_flat_grad", "grad_sync"):
            state[k] = None
        return state

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.lib().hd_net_destroy(self._handle)
        except Exception:
            pass

    def _build_table(self, flat_grad=None):
        units = self.units()
        offsets = None
        if flat_grad is not None:
            offsets, off = {}, 0
            for p in self.parameters():
                offsets[id(p)] = off
                off += p.numel()
        tab = (UnitPtrs * len(units))()

        def gaddr(p):
            if flat_grad is None or p is None:
                return None
            return flat_grad.data_ptr() + 4 * offsets[id(p)]

        for t, u in zip(tab, units):
            conv, bn = u.convolution, u.bn
            t.w, t.b = conv.weight.data_ptr(), (conv.bias.data_ptr() if conv.bias is not None else None)
            t.dw, t.db = gaddr(conv.weight), gaddr(conv.bias)
            if isinstance(bn, nn.BatchNorm2d):
                t.gamma, t.beta = bn.weight.data_ptr(), bn.bias.data_ptr()
                t.running_mean, t.running_var = bn.running_mean.data_ptr(), bn.running_var.data_ptr()
                t.num_batches_tracked = bn.num_batches_tracked.data_ptr()
                t.dgamma, t.dbeta = gaddr(bn.weight), gaddr(bn.bias)
        return tab

    def _check_params(self, device):
        for p in self.parameters():
            if p.device != device or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("StackedHourglass: parameters must be contiguous fp32 tensors on the input's device "
                                   f"(found {p.dtype} on {p.device}, input on {device})")
        for b in self.buffers():
            if b.device != device:
                raise RuntimeError("StackedHourglass: buffers must live on the input's device")

    # ------------------------------------------------------------------ execution
    def _run_forward(self, x):
        L = _lib.lib()
        h = self._native()
        if x.dtype != torch.float32:          # the stem kernel reads fp32 NCHW; never reinterpret another dtype's bytes
            x = x.float()
        x = x.contiguous()
        B, _, H, W = x.shape
        self._check_params(x.device)
        need_bwd = 1 if self.training else 0
        key = (B, H, W, need_bwd, x.device)
        if self._ws_key != key:
            nbytes = L.hd_net_workspace_bytes(h, B, H, W, need_bwd)
            self._workspace = None
            self._workspace = torch.empty((nbytes,), dtype=torch.uint8, device=x.device)
            self._ws_key = key
            if getattr(self, "_frozen", False):          # new workspace: the packed weights must be rebuilt once
                L.hd_net_set_static_weights(h, 1)
        logits = torch.empty((B, self.num_stack, self.out_ch, H // 4, W // 4), dtype=torch.float32, device=x.device)
        tab = self._build_table()
        self._generation += 1
        with torch.cuda.device(x.device):     # kernels, streams and per-device attributes follow the input's device
            check(L.hd_net_forward(h, tab, len(tab), ptr(x), ptr(logits), ptr(self._workspace), self._workspace.numel(),
                                   B, H, W, 1 if self.training else 0, stream(x.device)), "net_forward")
        return logits

    def _run_backward(self, dlogits):
        L = _lib.lib()
        params = self._param_list()
        total = sum(p.numel() for p in params)
        flat = torch.zeros((total,), dtype=torch.float32, device=dlogits.device)
        tab = self._build_table(flat)
        d = dlogits.contiguous().float()
        sync = self.grad_sync
        with torch.cuda.device(d.device):
            if sync is not None and getattr(sync, "overlap", False) and sync.active():
                # two buckets: everything but the 256x256 level (the tail of the flat buffer: parameters() lists the stem
                # and PreLayer's Residual(64,128) first) is exchanged on a communication stream while the 256x256 level's
                # backward (~3 ms) still runs; those 0.2 M gradients follow after the last weight-gradient kernel
                n_pre = sum(p.numel() for m in (self.pre_layer.layers[0], self.pre_layer.layers[1]) for p in m.parameters())
                comm = sync.comm_stream(d.device)
                args = (self._handle, tab, len(tab), ptr(d), ptr(self._workspace), self._workspace.numel(), stream(d.device))
                check(L.hd_net_backward_stage(*args, 1, c_void_p(comm.cuda_stream)), "net_backward_stage(1)")
                sync.early(flat[n_pre:], comm)
                check(L.hd_net_backward_stage(*args, 2, None), "net_backward_stage(2)")
                sync.late(flat[:n_pre])
            else:
                check(L.hd_net_backward(self._handle, tab, len(tab), ptr(d), ptr(self._workspace),
                                        self._workspace.numel(), stream(d.device)), "net_backward")
                if sync is not None:
                    sync(flat)
        self._flat_grad = flat
        grads, off = [], 0
        for p in params:
            grads.append(flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        return grads

    def forward(self, x):
        if not isinstance(x, torch.Tensor) or x.dim() != 4 or x.shape[1] != 3:
            raise RuntimeError(f"StackedHourglass: expected a (B,3,H,W) tensor, got {tuple(getattr(x, 'shape', ()))}")
        if x.shape[2] % 64 or x.shape[3] % 64 or x.shape[0] == 0:
            raise RuntimeError("StackedHourglass: H and W must be positive multiples of 64 "
                               f"(stem /2, pool /2, four hourglass levels), got {tuple(x.shape)}")
        _lib.require_cuda(x, "StackedHourglass input")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return _HourglassFn.apply(self, x, *self._param_list())
        with torch.no_grad():
            return self._run_forward(x.float())
