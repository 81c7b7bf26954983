#!/usr/bin/env python
"""Where the multi-GPU step time goes: CUDA-event time of every gradient collective and of the step,
per rank, for the overlapped two-bucket exchange and for the single flat all-reduce.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \\
        tools/allreduce_timing.py > profiles/r02_allreduce_timing.txt

Timing mode (`FlatAllReduce(timing=True)`) brackets each collective with events on the stream it is issued from and waits
for it there, so "stacks" = time from "bucket ready on the communication stream" to "all-reduce complete" (it overlaps the
PreLayer backward), "pre_layer" / "flat" = exposed time on the compute stream at the end of the backward pass.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from real_time_helmet_detection_b200.hourglass import StackedHourglass
    from real_time_helmet_detection_b200.loss import LossCalculator
    from real_time_helmet_detection_b200.parallel import attach_flat_allreduce, broadcast_parameters
    from real_time_helmet_detection_b200.synthetic import synthetic_targets
    from real_time_helmet_detection_b200.train import train_step
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B, steps, warm = 32, 12, 5
    torch.manual_seed(777)
    net = StackedHourglass(1, 128, 6).to(dev).train()
    broadcast_parameters(net)
    crit = LossCalculator(1.0, 1.0, 0.1, 2.0, 4.0).to(dev)
    x = torch.randn(B, 3, 512, 512, device=dev, generator=torch.Generator(dev).manual_seed(rank))
    gts = [torch.from_numpy(a).to(dev) for a in synthetic_targets(B, imsize=512)]

    def run(mode):
        hook = None
        if mode != "none":
            hook = attach_flat_allreduce(net, overlap=(mode in ("overlap", "overlap+events")), timing=mode.endswith("events"))
        else:
            net.grad_sync = None
        times = []
        for i in range(warm + steps):
            for p in net.parameters():
                p.grad = None
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier()
            torch.cuda.synchronize()
            if hook is not None:
                hook.events.clear()
            e0.record()
            train_step(net, crit, x, *gts)
            e1.record()
            torch.cuda.synchronize()
            if i >= warm:
                row = {"step_ms": e0.elapsed_time(e1)}
                if hook is not None:
                    for name, a, b in hook.events:
                        row[name + "_ms"] = a.elapsed_time(b)
                times.append(row)
        keys = sorted({k for r in times for k in r})
        mean = torch.tensor([sum(r.get(k, 0.0) for r in times) / len(times) for k in keys], device=dev)
        allr = [torch.empty_like(mean) for _ in range(world)]
        dist.all_gather(allr, mean)
        if rank == 0:
            m = torch.stack(allr).cpu()
            print(f"== mode {mode}: world {world}, batch {B}/GPU, {steps} steps after {warm} warm-up (barrier + sync before every step)")
            for j, k in enumerate(keys):
                col = m[:, j]
                print(f"   {k:14s} per rank: " + " ".join(f"{v:7.3f}" for v in col.tolist()) +
                      f"   | max {col.max():7.3f}  min {col.min():7.3f}  mean {col.mean():7.3f}")
            sys.stdout.flush()

    modes = os.environ.get("HD_TIMING_MODES", "none,flat,flat+events,overlap,overlap+events").split(",")
    if rank == 0:
        print("## env:", {k: v for k, v in os.environ.items() if k.startswith(("HD_", "NCCL_"))})
    for mode in modes:
        run(mode)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
