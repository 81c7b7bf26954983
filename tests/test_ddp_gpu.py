"""2-GPU test of the data-parallel path: one process per GPU, batch sharded, the flat gradient buffer averaged over NCCL
inside the backward node - as two overlapped buckets (stacks early on a communication stream, PreLayer at the end;
the default) or as one flat all-reduce (overlap=False). The averaged gradients must equal the mean of the two ranks'
local gradients (per-replica BatchNorm, like the reference's DDP: train.py:174-175)."""
import os
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _grads_for(net, crit, x, gts):
    from real_time_helmet_detection_b200.train import train_step
    for p in net.parameters():
        p.grad = None
    train_step(net, crit, x, *gts)
    return torch.cat([p.grad.flatten() for p in net.parameters()])


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from real_time_helmet_detection_b200.hourglass import StackedHourglass
    from real_time_helmet_detection_b200.loss import LossCalculator
    from real_time_helmet_detection_b200.parallel import attach_flat_allreduce, broadcast_parameters
    from real_time_helmet_detection_b200.synthetic import synthetic_targets
    torch.manual_seed(100 + rank)                       # different init per rank: the broadcast must fix it
    net = StackedHourglass(1, 128, 6).to(dev).train()
    broadcast_parameters(net, src=0)
    crit = LossCalculator(1.0, 1.0, 0.1, 2.0, 4.0)
    g = torch.Generator().manual_seed(5)
    x_all = torch.randn(4, 3, 128, 128, generator=g)
    gts_all = [torch.from_numpy(a) for a in synthetic_targets(4, imsize=128)]
    shard = slice(2 * rank, 2 * rank + 2)
    # reference: each shard alone, no communication
    singles = [_grads_for(net, crit, x_all[s].to(dev), [t[s].to(dev) for t in gts_all])
               for s in (slice(0, 2), slice(2, 4))]
    from real_time_helmet_detection_b200.parallel import FlatAllReduce

    class Probe(FlatAllReduce):                          # remembers this rank's gradient before the exchange
        def __call__(self, flat):
            self.local = flat.clone()
            super().__call__(flat)

        def early(self, bucket, comm):
            with torch.cuda.stream(comm):                # comm is already ordered after the bucket's producers
                self.local_tail = bucket.clone()
            super().early(bucket, comm)

        def late(self, bucket):
            self.local_head = bucket.clone()
            super().late(bucket)

    out = []
    for overlap in (False, True):
        hook = Probe(overlap=overlap)
        net.grad_sync = hook
        synced = _grads_for(net, crit, x_all[shard].to(dev), [t[shard].to(dev) for t in gts_all])
        torch.cuda.synchronize()
        local = torch.cat([hook.local_head, hook.local_tail]) if overlap else hook.local
        locals_ = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(locals_, local)
        mean = (locals_[0] + locals_[1]) / 2
        err = ((synced - mean).norm() / mean.norm()).item()                 # exact up to fp32 rounding
        drift = ((local - singles[rank]).norm() / singles[rank].norm()).item()   # run-to-run noise of one shard
        out.append((err, hook.calls, hook.elements, synced[:1000].cpu(), drift, hook.steps))
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_allreduce_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, out in res:
        for overlap, (err, calls, elements, head, drift, steps) in zip((False, True), out):
            # one flat all-reduce of the 19.94 MB buffer, or its two buckets (stacks + PreLayer)
            assert calls == (2 if overlap else 1) and elements == 4984070 and steps == 1, (overlap, calls, elements)
            assert err < 1e-5, (overlap, err)               # synced == mean of the two ranks' local gradients
            assert drift < 0.5, drift                       # same shard, second run: only reduction-order noise
    for i in range(2):
        assert torch.equal(res[0][1][i][3], res[1][1][i][3])   # both ranks hold identical averaged gradients


def _worker_torch_ddp(rank, world, port, q):
    """The reference's own wrapper (train.py:174-175): torch DistributedDataParallel around our module."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from torch.nn.parallel import DistributedDataParallel
    from real_time_helmet_detection_b200.hourglass import StackedHourglass
    from real_time_helmet_detection_b200.loss import LossCalculator
    from real_time_helmet_detection_b200.synthetic import synthetic_targets
    torch.manual_seed(100 + rank)
    net = StackedHourglass(1, 128, 6).to(dev).train()
    ddp = DistributedDataParallel(net, device_ids=[rank])          # broadcasts rank 0's parameters and buffers
    crit = LossCalculator(1.0, 1.0, 0.1, 2.0, 4.0)
    g = torch.Generator().manual_seed(5)
    x_all = torch.randn(4, 3, 128, 128, generator=g)
    gts_all = [torch.from_numpy(a).to(dev) for a in synthetic_targets(4, imsize=128)]
    shard = slice(2 * rank, 2 * rank + 2)
    out = ddp(x_all[shard].to(dev))
    loss = crit.forward_logits(out[:, 0], *[t[shard] for t in gts_all])
    loss.backward()
    flat = torch.cat([p.grad.flatten() for p in net.parameters()])
    w0 = torch.cat([p.detach().flatten() for p in net.parameters()])[:1000].cpu()
    q.put((rank, bool(torch.isfinite(flat).all()), float(flat.norm()), flat[:1000].cpu(), w0))
    dist.barrier()
    dist.destroy_process_group()


def test_torch_ddp_wrapper_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker_torch_ddp, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(r[1] for r in res) and res[0][2] > 0
    assert torch.equal(res[0][4], res[1][4])                 # parameters were broadcast from rank 0
    assert torch.equal(res[0][3], res[1][3])                 # DDP's reducer averaged the gradients of our backward node
