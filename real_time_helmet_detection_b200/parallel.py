"""Data-parallel training support: one process per GPU, batch sharded across ranks, gradients averaged over NCCL.

The reference wraps the network in torch's DistributedDataParallel (train.py:174-175), whose reducer all-reduces
bucket by bucket while autograd runs. Our whole backward pass is a single autograd node that writes every parameter
gradient into one contiguous fp32 buffer (19.94 MB for 1 stack), so the exchange is at most TWO NCCL all-reduces over
NVLink / NVSwitch on slices of that buffer:

  * bucket "stacks" (hourglass + neck + head [+ merge] and the two 128x128 Residuals of PreLayer: 96 % of the buffer for
    one stack) - enqueued on a communication stream as soon as that part of the backward has been enqueued
    (`hd_net_backward_stage`, csrc/net.cu), i.e. it runs under the ~3 ms of the 256x256 level's backward that follow;
    the persistent convolution grids of that stage leave 8 SMs free (HD_COMM_RESERVE) so the collective's CTAs can run
    beside them instead of between them;
  * bucket "pre_layer" (stem + Residual(64,128): 0.2 M parameters, 0.8 MB) - after the last weight-gradient kernel.

Averaging (DDP semantics: sum / world) is done by NCCL itself (`ReduceOp.AVG`), so no scaling kernel follows.
`overlap=False` restores the round-1 behaviour: one flat all-reduce after the whole backward pass. BatchNorm statistics
stay per-replica, as in the reference (no SyncBN). The drop-in path through torch DDP keeps working as well (the module
exposes ordinary nn.Parameters and returns ordinary gradients).
"""
from __future__ import annotations

import contextlib

import torch
import torch.distributed as dist


class FlatAllReduce:
    """Installed as `StackedHourglass.grad_sync`: averages the flat gradient buffer across ranks."""

    def __init__(self, process_group=None, average: bool = True, overlap: bool = True, timing: bool = False):
        self.group = process_group
        self.average = average
        self.overlap = overlap
        self.timing = timing            # record CUDA events around every collective (tools/allreduce_timing.py)
        self.calls = 0                  # collectives issued
        self.steps = 0                  # backward passes synchronised
        self.elements = 0
        self._comm = {}
        self._pending = None
        self.events = []                # timing: (bucket name, start event, end event)
        backend = dist.get_backend(process_group) if self.active() else None
        self._avg_op = self.active() and average and backend == "nccl"
        # the collectives can be recorded into a CUDA graph (train.GraphedTrainStep) when they run on NCCL and no timing
        # events are requested
        self.capturable = (not self.active()) or (backend == "nccl" and not timing)

    def active(self) -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def comm_stream(self, device) -> torch.cuda.Stream:
        st = self._comm.get(device)
        if st is None:
            st = self._comm[device] = torch.cuda.Stream(device=device)
        return st

    def _reduce(self, buf: torch.Tensor, name: str, async_op: bool):
        ev = None
        if self.timing:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        if self._avg_op:
            work = dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.group, async_op=async_op)
        else:                           # gloo (CPU tests) has no AVG
            work = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
        self.calls += 1
        return work, ev

    def _finish(self, buf, work, ev, name):
        if work is not None:
            work.wait()                 # the current stream waits for the collective; the host does not block (NCCL)
        if self.average and not self._avg_op:
            buf.mul_(1.0 / dist.get_world_size(self.group))
        if ev is not None:
            ev[1].record()
            self.events.append((name, ev[0], ev[1]))

    # ---- two-bucket protocol driven by StackedHourglass._run_backward
    def early(self, bucket: torch.Tensor, comm: torch.cuda.Stream) -> None:
        """`comm` already waits for the producers of `bucket` (hd_net_backward_stage). The collective is issued with
        `comm` current, so NCCL's stream orders itself after it; completion is joined into the compute stream in late()."""
        # comm None: host tensors (the gloo tests of this protocol)
        with (torch.cuda.stream(comm) if comm is not None else contextlib.nullcontext()):
            work, ev = self._reduce(bucket, "stacks", async_op=True)
            if self.timing:
                self._finish(bucket, work, ev, "stacks")
                work = None
        self._pending = (bucket, work, comm)
        self.elements = bucket.numel()

    def late(self, bucket: torch.Tensor) -> None:
        work, ev = self._reduce(bucket, "pre_layer", async_op=False)
        self._finish(bucket, None, ev, "pre_layer")
        pbuf, pwork, comm = self._pending
        self._pending = None
        if pwork is not None:
            self._finish(pbuf, pwork, None, "stacks")
        if comm is not None:
            torch.cuda.current_stream(bucket.device).wait_stream(comm)
        self.elements += bucket.numel()
        self.steps += 1

    # ---- single flat all-reduce after the whole backward pass (overlap=False, or a caller without staging)
    def __call__(self, flat: torch.Tensor) -> None:
        if not self.active():
            return
        work, ev = self._reduce(flat, "flat", async_op=False)
        self._finish(flat, None, ev, "flat")
        self.elements = flat.numel()
        self.steps += 1


def attach_flat_allreduce(network, process_group=None, overlap: bool = True, timing: bool = False) -> FlatAllReduce:
    """Make `network` (StackedHourglass) average its gradients across ranks inside its backward pass."""
    hook = FlatAllReduce(process_group, overlap=overlap, timing=timing)
    network.grad_sync = hook
    return hook


def broadcast_parameters(network, src: int = 0, process_group=None) -> None:
    """Rank `src` -> all, for parameters and buffers (what DDP's constructor does, train.py:175)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(process_group) == 1:
        return
    for t in list(network.parameters()) + list(network.buffers()):
        dist.broadcast(t.data, src=src, group=process_group)


def shard_batch(global_batch: int, rank: int, world: int) -> range:
    """Contiguous per-rank slice of a global batch (the reference: per-GPU batch = batch_size / ngpus, train.py:38)."""
    per = global_batch // world
    return range(rank * per, (rank + 1) * per)
