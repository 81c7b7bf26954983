"""Data-parallel training support: one process per GPU, batch sharded across ranks, ONE flat gradient all-reduce.

The reference wraps the network in torch's DistributedDataParallel (train.py:174-175), whose reducer all-reduces
bucket by bucket while autograd runs. Our whole backward pass is a single autograd node that already writes every
parameter gradient into one contiguous fp32 buffer, so the exchange step is exactly one NCCL all-reduce over NVLink /
NVSwitch on that buffer (19.94 MB for 1 stack), enqueued on the compute stream right after the last wgrad kernel,
followed by an in-place 1/world scale — the semantics of DDP's gradient averaging. BatchNorm statistics stay
per-replica, as in the reference (no SyncBN). The drop-in path through torch DDP keeps working as well (the module
exposes ordinary nn.Parameters and returns ordinary gradients).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class FlatAllReduce:
    """Callable installed as `StackedHourglass.grad_sync`: sum-all-reduce + average of the flat gradient buffer."""

    def __init__(self, process_group=None, average: bool = True):
        self.group = process_group
        self.average = average
        self.calls = 0
        self.elements = 0

    def __call__(self, flat: torch.Tensor) -> None:
        if not dist.is_available() or not dist.is_initialized():
            return
        world = dist.get_world_size(self.group)
        if world == 1:
            return
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        if self.average:
            flat.mul_(1.0 / world)
        self.calls += 1
        self.elements = flat.numel()


def attach_flat_allreduce(network, process_group=None) -> FlatAllReduce:
    """Make `network` (StackedHourglass) average its gradients across ranks with a single flat all-reduce."""
    hook = FlatAllReduce(process_group)
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
