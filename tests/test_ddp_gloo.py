"""CPU (gloo, world_size 2) test of the multi-GPU host logic: the flat gradient all-reduce hook and batch sharding."""
import os
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from real_time_helmet_detection_b200.parallel import FlatAllReduce, shard_batch, broadcast_parameters
    flat = torch.arange(10, dtype=torch.float32) * (rank + 1)
    hook = FlatAllReduce()
    hook(flat)
    # the two-bucket protocol StackedHourglass._run_backward drives (early: stacks' tail of the buffer, late: PreLayer head)
    flat2 = torch.arange(10, dtype=torch.float32) * (rank + 1)
    hook2 = FlatAllReduce()
    assert hook2.overlap and hook2.active()
    hook2.early(flat2[4:], None)
    hook2.late(flat2[:4])
    assert flat2.tolist() == [1.5 * i for i in range(10)] and hook2.calls == 2 and hook2.elements == 10 and hook2.steps == 1
    lin = torch.nn.Linear(3, 2)
    with torch.no_grad():
        lin.weight.fill_(float(rank + 5))
    broadcast_parameters(lin, src=0)
    q.put((rank, flat.tolist(), hook.calls, list(shard_batch(8, rank, world)), lin.weight[0, 0].item()))
    dist.destroy_process_group()


def test_flat_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    expect = [1.5 * i for i in range(10)]           # mean of 1x and 2x
    for rank, flat, calls, shard, w in res:
        assert flat == expect and calls == 1 and w == 5.0
    assert res[0][3] == [0, 1, 2, 3] and res[1][3] == [4, 5, 6, 7]
