"""Training-step driver — the body of the reference's hot loop (train.py:92-139) over the B200 kernels.

`train_step` is the call a user of this package makes per iteration: host (or device) batch in, scalar loss out,
parameter gradients accumulated. It mirrors train.py:97-139: forward, per-stack split + head activation + loss
(fused into one kernel per stack here), sum over stacks, optional GradScaler, backward, and - every `sub_divisions`
iterations, the reference's gradient-accumulation flag (train.py:124) - the optimizer step + zero_grad.
`load_network` is the factory of train.py:164-201 with the same 4-tuple return, DDP wrap and checkpoint-resume rules.
"""
from __future__ import annotations

import time

import torch

from .hourglass import StackedHourglass
from .loss import LossCalculator
from .optim import get_optimizer


def load_network(args, device):
    """`args`: the reference's argparse namespace (config.py). Returns (network, optimizer, scheduler, loss_calculator)
    exactly like train.py:164-201: optimizer / scheduler / loss calculator only with --train-flag, DistributedDataParallel
    only for a multi-GPU training run, and `--model-load` restores network (+ optimizer, loss log, scheduler)."""
    net_kwargs = dict(num_stack=args.num_stack, in_ch=args.hourglass_inch, out_ch=args.num_cls + 4,
                      increase_ch=args.increase_ch, activation=args.activation, pool=args.pool,
                      neck_activation=args.neck_activation, neck_pool=args.neck_pool)
    network = StackedHourglass(**net_kwargs).to(device)
    training = bool(getattr(args, "train_flag", False))
    if training and len(getattr(args, "gpu_no", [0])) > 1:
        network = torch.nn.parallel.DistributedDataParallel(network, device_ids=[device])
    optimizer = scheduler = loss_calculator = None
    if training:
        optimizer, scheduler = get_optimizer(network=network, lr=args.lr, lr_milestone=args.lr_milestone,
                                             lr_gamma=args.lr_gamma)
        loss_calculator = LossCalculator(hm_weight=args.hm_weight, offset_weight=args.offset_weight,
                                         size_weight=args.size_weight, focal_alpha=args.focal_alpha,
                                         focal_beta=args.focal_beta).to(device)
    path = getattr(args, "model_load", None)
    if path:
        ckpt = torch.load(path, map_location=device, weights_only=False)
        network.load_state_dict(ckpt["state_dict"])
        print("%s: Weights are loaded from %s" % (time.ctime(), path))
        if training:
            optimizer.load_state_dict(ckpt["optimizer"])
            loss_calculator.log = ckpt["loss_log"]
            if scheduler is not None:
                scheduler.load_state_dict(ckpt["scheduler"])
    return network, optimizer, scheduler, loss_calculator


def train_step(network, loss_calculator, image, gt_heatmap, gt_offset, gt_size, gt_mask, num_cls=2,
               normalized_coord=False, scaler=None, optimizer=None, step=True):
    """One iteration of train.py:92-139. Tensors may live on the host (pinned for async copies) or the device.
    `step=False` accumulates gradients without touching the optimizer (the reference's `--sub-divisions`: pass
    `step=(iteration % sub_divisions == 0)`)."""
    device = next(network.parameters()).device
    image = image.to(device, non_blocking=True)
    gts = [t.to(device, non_blocking=True) for t in (gt_heatmap, gt_offset, gt_size, gt_mask)]   # once, not per stack
    outputs = network(image)                                   # (B, S, num_cls+4, H/4, W/4) fp32 logits
    S = outputs.shape[1]
    total_loss = None
    for s in range(S):
        # one stack: a squeeze is a pure view in both directions; slicing would make autograd allocate a zero tensor of
        # the full output and copy the slice's gradient into it (a fill + a copy on the critical path of every step)
        logits_s = outputs.squeeze(1) if S == 1 else outputs[:, s]
        loss_s = loss_calculator.forward_logits(logits_s, *gts, num_cls=num_cls, normalized_coord=normalized_coord)
        total_loss = loss_s if total_loss is None else total_loss + loss_s
    do_step = step and optimizer is not None
    if scaler is not None:
        scaler.scale(total_loss).backward()
        if do_step:
            scaler.step(optimizer)
            scaler.update()
    else:
        total_loss.backward()
        if do_step:
            optimizer.step()
    if do_step:
        optimizer.zero_grad(set_to_none=True)
    return total_loss.detach()


class GraphedTrainStep:
    """One training iteration (forward + fused loss + backward, train.py:97-134) recorded ONCE into a CUDA graph and
    replayed: the ~260 kernel launches, ~100 cross-stream event edges and the autograd bookkeeping of a step become
    one `cudaGraphLaunch`. Shapes are fixed by the example batch. Parameter gradients land in `p.grad` (static views of
    the flat gradient buffer of the replayed graph), so an optimizer step between calls works as usual; BatchNorm running
    statistics are updated by the replayed kernels exactly as in the eager step.

    Two ways to feed it:
      * `step(image, ghm, goff, gsize, gmask)`  - copies the batch (host or device tensors) into the static inputs on the
        compute stream, replays, returns the loss;
      * `stage(batch)` / `run()`                 - double-buffered: `buffers=2` input sets with one graph each; `stage`
        copies the NEXT batch (e.g. pinned host tensors) into the idle set on a copy stream while `run()` replays the
        graph of the set staged before, so the H2D copy of step i+1 overlaps the compute of step i with no extra
        device-to-device copy.

    Data parallel: the NCCL exchange of `parallel.FlatAllReduce` inside the backward node is recorded with the rest (NCCL
    collectives are graph-capturable); every rank must build and replay its graph in lockstep. The network must be in
    train() mode.
    """

    def __init__(self, network, loss_calculator, image, gt_heatmap, gt_offset, gt_size, gt_mask, num_cls=2,
                 normalized_coord=False, warmup=3, buffers=1):
        sync = getattr(network, "grad_sync", None)
        if sync is not None and not getattr(sync, "capturable", False):
            raise RuntimeError("GraphedTrainStep: the attached gradient exchange hook cannot be recorded into a CUDA graph "
                               "(parallel.FlatAllReduce over NCCL can)")
        if not network.training:
            raise RuntimeError("GraphedTrainStep records the training-mode step: call network.train() first")
        self.network, self.loss_calculator = network, loss_calculator
        self.num_cls, self.normalized_coord = num_cls, normalized_coord
        device = next(network.parameters()).device
        self.device = device
        example = (image, gt_heatmap, gt_offset, gt_size, gt_mask)
        self.sets = [[t.to(device, copy=True) for t in example] for _ in range(max(1, int(buffers)))]
        self.static_in = self.sets[0]
        cur = torch.cuda.current_stream(device)
        side = torch.cuda.Stream(device=device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):                  # eager warm-up: arena, streams, events, kernel attributes
            for _ in range(max(1, warmup)):        # at least one: the capture relies on what an eager pass set up
                self._eager(self.sets[0])
        cur.wait_stream(side)
        torch.cuda.synchronize(device)
        pending = list(loss_calculator._pending)       # the capture must not leave graph-pool tensors in the log queue
        self.graphs, self.losses, self.values, self.grads = [], [], [], []
        params = list(network.parameters())
        from . import _lib
        for inputs in self.sets:
            for p in params:
                p.grad = None
            g = torch.cuda.CUDAGraph()
            l0 = _lib.lib().hd_launch_count()
            with torch.cuda.graph(g):
                loss = self._eager(inputs)
            self.kernels_per_replay = int(_lib.lib().hd_launch_count() - l0)   # library kernels recorded into one graph
            self.graphs.append(g)
            self.losses.append(loss)
            self.values.append(list(loss_calculator._pending[len(pending):]))
            self.grads.append([p.grad for p in params])
            loss_calculator._pending = list(pending)
        self._params = params
        self._cur = 0                                   # set whose graph `run()` replays next
        self._staged = [None] * len(self.sets)          # copy-complete events
        self._used = [None] * len(self.sets)            # replay-complete events
        self._copy_stream = torch.cuda.Stream(device=device)
        self.static_loss, self.static_values = self.losses[0], self.values[0]

    def _eager(self, inputs):
        for p in self.network.parameters():
            p.grad = None
        outputs = self.network(inputs[0])
        S = outputs.shape[1]
        total = None
        for s in range(S):
            logits_s = outputs.squeeze(1) if S == 1 else outputs[:, s]
            loss_s = self.loss_calculator.forward_logits(logits_s, *inputs[1:], num_cls=self.num_cls,
                                                         normalized_coord=self.normalized_coord)
            total = loss_s if total is None else total + loss_s
        total.backward()
        return total.detach()

    def _replay(self, k, log):
        self.graphs[k].replay()
        if len(self.graphs) > 1:                        # p.grad must name the buffer THIS graph wrote
            for p, g in zip(self._params, self.grads[k]):
                p.grad = g
        if log:                                         # LossCalculator.log keeps working: one small copy per stack
            for v in self.values[k]:
                self.loss_calculator._record(v.clone())
        return self.losses[k]

    def __call__(self, image, gt_heatmap, gt_offset, gt_size, gt_mask, log=True):
        for dst, src in zip(self.sets[0], (image, gt_heatmap, gt_offset, gt_size, gt_mask)):
            if src is not dst:
                dst.copy_(src, non_blocking=True)
        return self._replay(0, log).clone()

    step = __call__

    def stage(self, batch):
        """Copy `batch` (5 tensors, e.g. pinned host memory) into the next idle input set on the copy stream."""
        k = (self._cur + sum(e is not None for e in self._staged)) % len(self.sets)
        if self._staged[k] is not None:
            raise RuntimeError("GraphedTrainStep.stage: every input set already holds a staged batch; call run() first")
        if self._used[k] is not None:
            self._copy_stream.wait_event(self._used[k])          # the replay that last read this set has finished
        with torch.cuda.stream(self._copy_stream):
            for dst, src in zip(self.sets[k], batch):
                dst.copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        self._staged[k] = ev

    def run(self, log=True):
        """Replay the graph of the oldest staged input set; returns its loss tensor (static: clone to keep it)."""
        k = self._cur
        if self._staged[k] is None:
            raise RuntimeError("GraphedTrainStep.run: no staged batch")
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self._staged[k])
        self._staged[k] = None
        loss = self._replay(k, log)
        ev = torch.cuda.Event()
        ev.record(cur)
        self._used[k] = ev
        self._cur = (k + 1) % len(self.sets)
        return loss


class DevicePrefetcher:
    """Iterates over host batches (tuples of pinned CPU tensors) and yields them on the device, copying batch i+1 on a
    side stream while batch i is being consumed — the H2D copy of every step still happens, it just overlaps compute.

    Plays the role of `image.to(device)` / `gt.to(device)` in the reference loop (train.py:99,115-118), which copies
    synchronously (and re-copies the targets once per stack).
    """

    def __init__(self, batches, device):
        self.it = iter(batches)
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.next = None
        self._preload()

    def _preload(self):
        # the iterator itself runs under the side stream too: a device-side collate (data.DeviceCollate) then issues its
        # H2D copies and its normalise / GT-encode kernels there, off the compute stream
        with torch.cuda.stream(self.stream):
            try:
                host = next(self.it)
            except StopIteration:
                self.next = None
                return
            self.next = tuple(t.to(self.device, non_blocking=True) if torch.is_tensor(t) else t for t in host)

    def __iter__(self):
        return self

    def __next__(self):
        if self.next is None:
            raise StopIteration
        torch.cuda.current_stream(self.device).wait_stream(self.stream)
        batch = self.next
        for t in batch:
            if torch.is_tensor(t):
                t.record_stream(torch.cuda.current_stream(self.device))
        self._preload()
        return batch
