#!/usr/bin/env python
"""Benchmark of the B200 hot path: images/sec @512x512, train fwd+bwd (BASELINE.json `metric`).

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun for N>1)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU cores

One JSON line on stdout (rank 0). Workload at N=1: BASELINE.json configs[1] — 1-stack hourglass, 2 classes, 512x512,
batch 32 per GPU, bf16 activations, forward + fused loss + backward (+ one flat NCCL gradient all-reduce for N>1).
`value`   : device-resident synthetic inputs, CUDA-event timed, max over ranks (weak scaling: 32 images per GPU).
`e2e`     : the same step through the public API `real_time_helmet_detection_b200.train.train_step` with PINNED HOST
            inputs: H2D copy of image + targets and a D2H read of the loss inside the timed region, every step.
`roofline`: the dominant kernel (tcgen05 implicit-GEMM conv 3x3 128->128 @128x128, B=32), timed live here with CUDA
            events on its launch stream; achieved = algorithmic FLOPs / launch duration vs the measured bf16 peak.
`cpu_baseline`: the oracle port of the reference path (oracle/, fp32 PyTorch CPU) on the host cores, bounded sample.
Timing hygiene: >= 3 warm-up steps; the per-step working set (~10 GB of activations) is far larger than the 126 MB
L2, so no explicit L2 flush is needed between timed iterations (stated in config.l2).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_IMG_TRAIN = {1: 239.189e9, 2: 329.421e9}     # conv FLOPs per 512^2 image, fwd + bwd
METRIC = "images/sec @512x512 (train fwd+bwd)"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9 or not (t0 <= ts <= t1 + 0.2):
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def synthetic_batch(torch, B, size, seed):
    from real_time_helmet_detection_b200.synthetic import synthetic_targets
    g = torch.Generator().manual_seed(777 + seed)
    image = torch.randn(B, 3, size, size, generator=g)
    gts = [torch.from_numpy(a) for a in synthetic_targets(B, imsize=size)]
    return image, gts


# ------------------------------------------------------------------------------------------------ reference arm
class CpuArm:
    """The reference's train-loop body (train.py:99-134, with the out-of-place squeeze) on the host CPU, fp32.
    kind "reference": the UNMODIFIED modules staged under baseline/_ref/ (hourglass.StackedHourglass, loss.LossCalculator,
    optim.get_optimizer); kind "port": the oracle restatement (oracle/), used only when the reference is not staged."""

    def __init__(self, torch, S):
        from baseline import refload
        self.torch, self.S = torch, S
        self.kind = "reference" if refload.available() else "port"
        if self.kind == "reference":
            mods = refload.load_reference(("hourglass", "loss", "optim"))
            torch.manual_seed(777)
            self.net = mods["hourglass"].StackedHourglass(num_stack=S, in_ch=128, out_ch=6).train()
            self.crit = mods["loss"].LossCalculator(1.0, 1.0, 0.1, 2.0, 4.0)
            self.opt, _ = mods["optim"].get_optimizer(self.net, 5e-4, None, 0.1)
        else:
            self.sd = cpu_model(torch, S)
            self.opt = torch.optim.Adam([v for v in self.sd.values() if v.requires_grad], lr=5e-4)

    def step(self, x, gts, adam=False):
        torch = self.torch
        if self.kind == "port":
            loss = cpu_reference_step(torch, self.sd, x, gts, self.S)
        else:
            outputs = self.net(x)
            total = 0
            for output in outputs.split(1, dim=1):
                output = output.squeeze(1)
                phm, poff, psz = output.split([2, 2, 2], dim=1)
                total = total + self.crit(torch.sigmoid(phm), poff, psz, *gts)
            total.backward()
            loss = float(total.detach())
        if adam:
            self.opt.step()
        self.opt.zero_grad(set_to_none=True)
        return loss


def cpu_reference_step(torch, sd, x, gts, S):
    """Oracle-port flavour of the same loop body (only when baseline/_ref is absent)."""
    from oracle import hourglass_ref, loss_ref
    for v in sd.values():
        if v.requires_grad:
            v.grad = None
    out = hourglass_ref.stacked_hourglass_forward(sd, x, training=True)
    tot = sum(loss_ref.losses_from_logits(out[:, s], *gts)[3] for s in range(S))
    tot.backward()
    return float(tot.detach())


def cpu_model(torch, S):
    from real_time_helmet_detection_b200.hourglass import StackedHourglass
    torch.manual_seed(777)
    net = StackedHourglass(S, 128, 6)
    return {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone())
            for k, v in net.state_dict().items()}


def usable_cores():
    """Host cores this process may really use: affinity mask capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(float(quota) / float(period))))
    except Exception:
        pass
    return max(1, n)


def cpu_name():
    try:
        return [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        return ""


def time_cpu(torch, S, size, batch, steps, warmup, adam=False, arm=None):
    # a batch-2 fp32 train step has parallel slack for a few dozen threads at most: beyond that PyTorch's intra-op
    # pool only adds contention (measured 50x slower with 128 threads on the pool's host), so threads are capped at 32
    cores = min(usable_cores(), 32)
    torch.set_num_threads(cores)
    arm = arm or CpuArm(torch, S)
    x, gts = synthetic_batch(torch, batch, size, 0)
    for i in range(warmup):
        t = time.perf_counter()
        arm.step(x, gts, adam)
        if time.perf_counter() - t > 20.0:        # very slow host: keep the run bounded
            steps = 1
            break
    t0 = time.perf_counter()
    for _ in range(steps):
        arm.step(x, gts, adam)
    dt = (time.perf_counter() - t0) / steps
    return batch / dt, dt, cores, arm.kind


def cpu_decode_us(torch, S=1, runs=200):
    """Config 5 on the host: the reference's `Prediction` minus the network (evaluate.py:126-182: per-stack split ->
    sigmoid -> hm2box -> cat -> torchvision NMS) on the synthetic blob head, median wall time over `runs` calls.
    kind "reference": the unmodified evaluate.Prediction / transform.hm2box from baseline/_ref; "port": oracle/decode_ref."""
    from baseline import refload
    from real_time_helmet_detection_b200.synthetic import synthetic_head
    head = synthetic_head(S=S)
    torch.set_num_threads(min(usable_cores(), 32))
    if refload.available():
        ev = refload.load_reference(("evaluate",))["evaluate"]
        t = torch.from_numpy(head)

        class Fixed(torch.nn.Module):
            def forward(self, x):
                return t.clone()

        pred = ev.Prediction(Fixed(), 100, 4, 0.2, "nms", 0.2).eval()
        fn, kind = (lambda: pred(None)), "reference"
    else:
        from oracle import decode_ref
        fn, kind = (lambda: decode_ref.predict(head)), "port"
    wall = []
    with torch.no_grad():
        for _ in range(10):
            out = fn()
        for _ in range(runs):
            t0 = time.perf_counter()
            out = fn()
            wall.append(time.perf_counter() - t0)
    return {"us": statistics.median(wall) * 1e6, "kind": kind, "runs": runs, "boxes": int(len(out[0][0])),
            "what": "Prediction minus network (sigmoid, hm2box, cat, NMS) on the host CPU, fp32, median wall time"}


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    S, size, batch = args.num_stack, args.imsize, 2
    steps, warmup = max(1, min(args.steps, 30)), max(1, min(args.warmup, 3))
    ips, dt, cores, kind = time_cpu(torch, S, size, batch, steps, warmup)
    what = ("the UNMODIFIED reference modules (baseline/_ref: hourglass.StackedHourglass + loss.LossCalculator, "
            "train.py:99-134 loop body)" if kind == "reference" else "oracle port of the reference path")
    sample = (f"{what}, fp32 PyTorch CPU, {steps} timed steps of a bounded batch-{batch} "
              f"sample of the 512x512 workload (fwd + loss + bwd), {cores} threads, {cpu_name()}")
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": "img/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{S}-stack hourglass, 2 classes, {size}x{size}, train fwd+bwd", "batch_per_step": batch},
            "cpu_baseline": {"value": ips, "unit": "img/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": ips, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ library bar
def library_bar(torch, dev, S, B, size, steps=6, warmup=3):
    """The only perf comparator that matters: the reference's network executed by the libraries the
    reference itself calls (nn.Conv2d -> cuDNN, nn.BatchNorm2d, ... hourglass.py:100-103) on THIS GPU, bf16 autocast
    (train.py:97 with the dtype BASELINE names), cudnn.benchmark (train.py:34), same step (fwd + loss + bwd), same batch,
    CUDA-event timed. NCHW (the reference as written) and channels_last (the best stock layout)."""
    from baseline import torch_eager
    from real_time_helmet_detection_b200.loss import LossCalculator
    res = {"workload": f"{S}-stack hourglass, {size}x{size}, batch {B}, train fwd + loss + bwd, bf16 autocast, "
                       "cudnn.benchmark", "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}
    try:
        net, kind = torch_eager.reference_network(S, None, dev)
        net.train()
        crit = torch_eager.reference_loss(dev) or LossCalculator(1.0, 1.0, 0.1, 2.0, 4.0).to(dev)
        res["kind"] = kind + (" (unmodified hourglass.py / loss.py from baseline/_ref)" if kind == "reference"
                              else " (torch.nn.functional port over the same parameter tree; reference not staged)")
        x, gts = synthetic_batch(torch, B, size, 0)
        x, gts = x.to(dev), [g.to(dev) for g in gts]
        torch.backends.cudnn.benchmark = True

        def step(inp):
            for p in net.parameters():
                p.grad = None
            torch_eager.train_loop_body(net, crit, inp, gts, autocast_dtype=torch.bfloat16)

        for name in ("nchw", "channels_last"):
            inp = x
            if name == "channels_last":
                net.to(memory_format=torch.channels_last)
                inp = x.contiguous(memory_format=torch.channels_last)
            for _ in range(warmup):
                step(inp)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step(inp)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            res[name] = {"ms_per_step": ms, "img_s": B / ms * 1e3}
        res["best_img_s"] = max(res["nchw"]["img_s"], res["channels_last"]["img_s"])
    except Exception as exc:                      # e.g. out of memory on a shared box: keep the bench line alive
        res["error"] = f"{type(exc).__name__}: {str(exc)[:200]}"
    finally:
        net = crit = None
        torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------------------ our arm
def dominant_kernel_roofline(torch, dev, peaks, reps=20):
    """conv 3x3 128->128 @128x128, B=32 (8 forward + 8 dgrad instances per step): live CUDA-event timing."""
    from real_time_helmet_detection_b200 import ops
    B, H, W, C = 32, 128, 128, 128
    xs = [torch.randn(B, H, W, C, device=dev).to(torch.bfloat16) for _ in range(3)]   # 3 x 134 MB > L2
    w = ops.pack_weight(torch.randn(C, C, 3, 3, device=dev) * 0.03)
    out = torch.empty(B, H, W, C, device=dev, dtype=torch.bfloat16)
    stats = torch.zeros(2, C, device=dev)
    for i in range(3):
        ops.conv2d_igemm(xs[i % 3], w, C, 3, stats=stats, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(reps):
        ops.conv2d_igemm(xs[i % 3], w, C, 3, stats=stats, out=out)
    e1.record()
    torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) * 1e-3 / reps
    flops = 2.0 * B * H * W * C * C * 9
    achieved = flops / dt / 1e12
    peak = peaks["bf16_tflops"]
    return {"bound": "tensor", "kernel": "conv_igemm_halo_kernel 3x3 128->128 @128x128 B=32 (fwd, BN-stat epilogue)",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "peak_source": f"{peaks['source']} bf16 burst (kernel timed alone)", "us_per_launch": dt * 1e6,
            "flops_per_launch": flops,
            # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this shape from one `ncu --set full` capture,
            # recorded (with the name of the capture's summary file) in profiles/conv_halo_traffic.json - a profiler
            # figure cannot be taken inside a timed run; algorithmic bytes = 134.2 MB in + 134.2 MB out + 0.3 MB weights
            **recorded_traffic(), "traffic_unit": "bytes/launch", "algorithmic_bytes_per_launch": 268.7e6}


def recorded_traffic():
    p = os.path.join(ROOT, "profiles", "conv_halo_traffic.json")
    try:
        d = json.load(open(p))
        return {"traffic": float(d["dram_bytes_read"]) + float(d["dram_bytes_write"]),
                "traffic_source": f"profiles/conv_halo_traffic.json <- {d['source']}"}
    except Exception:
        return {"traffic": None, "traffic_source": "profiles/conv_halo_traffic.json missing"}


def decode_latency(torch, dev, S=1, runs=300):
    """Config 5: decode + NMS at batch 1 (topk 100, conf 0.2, nms 0.2) on the synthetic blob head tensor.
    device_us: the fused kernel alone, back-to-back launches on pre-allocated outputs, CUDA events on its stream;
    api_wall_us: `Prediction.decode` as a user calls it (allocations + launch + the one count read-back sync)."""
    from real_time_helmet_detection_b200.synthetic import synthetic_head
    from real_time_helmet_detection_b200.evaluate import Prediction
    from real_time_helmet_detection_b200.transform import _decode_call, _decode_buffers
    head = torch.from_numpy(synthetic_head(S=S)).to(dev)
    pred = Prediction(None, topk=100, scale_factor=4, conf_th=0.2, nms="nms", nms_th=0.2)
    for _ in range(10):
        b, _, _ = pred.decode(head)
    B, S_, O, H, W = head.shape
    C, hw = O - 4, H * W
    strides = ((S_ * O * hw, O * hw),) * 3
    bufs = _decode_buffers(dev, B, S_, C, H, W, 100)
    off_v, wh_v = head[:, :, C:], head[:, :, C + 2:]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(runs):
        _decode_call(head, off_v, wh_v, strides, B, S_, C, H, W, 100, 4, 0.2, 0.2, False, True, True, bufs=bufs)
    e1.record()
    torch.cuda.synchronize()
    dev_us = e0.elapsed_time(e1) * 1e3 / runs
    # the same two launches replayed from a CUDA graph ("also report with CUDA-graph replay")
    graph_us = None
    try:
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            _decode_call(head, off_v, wh_v, strides, B, S_, C, H, W, 100, 4, 0.2, 0.2, False, True, True, bufs=bufs)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            _decode_call(head, off_v, wh_v, strides, B, S_, C, H, W, 100, 4, 0.2, 0.2, False, True, True, bufs=bufs)
        for _ in range(5):
            g.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(runs):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        graph_us = e0.elapsed_time(e1) * 1e3 / runs
    except Exception as exc:      # keep the benchmark line alive if capture is not possible on some driver
        graph_us = f"unavailable: {type(exc).__name__}"
    wall = []
    for _ in range(runs):
        t0 = time.perf_counter()
        b, c, s = pred.decode(head)       # includes the count D2H read (the one sync the API's output shapes need)
        wall.append(time.perf_counter() - t0)
    return {"workload": f"decode+NMS 512x512 batch 1, {S} stack, topk 100, conf 0.2, nms 0.2", "device_us": dev_us,
            "graph_replay_us": graph_us, "api_wall_us": statistics.median(wall) * 1e6, "boxes": int(b[0].shape[0]),
            "algorithmic_bytes": 6 * H * W * 4 * S}


def inference_latency(torch, dev, S=1, runs=50):
    """Batch-1 inference as the reference's demo times it (PytorchToCpp/main.cpp:60-67, README.md:76: "100 FPS" on a
    GTX 1080 Ti): network forward (eval mode) + fused decode/NMS, host wall clock ending with the result read-back."""
    from real_time_helmet_detection_b200.hourglass import StackedHourglass
    from real_time_helmet_detection_b200.evaluate import Prediction
    torch.manual_seed(0)
    net = StackedHourglass(S, 128, 6).to(dev).eval().freeze_weights()      # inference: parameters are static
    x = torch.randn(1, 3, 512, 512, device=dev)
    res = {}
    for name, graph in (("eager", False), ("cuda_graph", True)):
        pred = Prediction(net, topk=100, scale_factor=4, conf_th=0.2, nms="nms", nms_th=0.2, cuda_graph=graph)
        for _ in range(5):
            pred(x)
        torch.cuda.synchronize()
        wall = []
        for _ in range(runs):
            t0 = time.perf_counter()
            pred(x)
            wall.append(time.perf_counter() - t0)
        res[name] = statistics.median(wall) * 1e3
    ms = res["cuda_graph"]
    return {"workload": f"forward (eval) + decode + NMS, 512x512 batch 1, {S} stack, Prediction(cuda_graph=True)",
            "ms": ms, "fps": 1e3 / ms, "eager_ms": res["eager"],
            "reference": "README.md:76: 100 FPS (10 ms) on a GTX 1080 Ti, TorchScript C++ app"}


def config3_line(torch, dev, peaks, steps=10, warmup=4):
    """BASELINE.json configs[2] in the same run: 2-stack hourglass, 512x512, batch 16, bf16, train fwd + loss + bwd."""
    from real_time_helmet_detection_b200.hourglass import StackedHourglass
    from real_time_helmet_detection_b200.loss import LossCalculator
    from real_time_helmet_detection_b200.train import train_step
    S, B, size = 2, 16, 512
    try:
        torch.manual_seed(777)
        net = StackedHourglass(S, 128, 6).to(dev).train()
        crit = LossCalculator(1.0, 1.0, 0.1, 2.0, 4.0).to(dev)
        x, gts = synthetic_batch(torch, B, size, 0)
        x, gts = x.to(dev), [g.to(dev) for g in gts]

        def step():
            for p in net.parameters():
                p.grad = None
            train_step(net, crit, x, *gts)

        def time_it(fn):
            for _ in range(warmup):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / steps

        ms_eager = time_it(step)
        ms, launch = ms_eager, "eager (kernel by kernel)"
        try:                                              # the same iteration replayed from a CUDA graph, as for config 2
            from real_time_helmet_detection_b200.train import GraphedTrainStep
            graphed = GraphedTrainStep(net, crit, x, *gts, warmup=0)
            ms = time_it(lambda: graphed(*graphed.static_in, log=False))
            launch = "CUDA graph replay (train.GraphedTrainStep)"
        except Exception as exc:
            launch += f" (graph capture failed: {type(exc).__name__})"
        ips = B / ms * 1e3
        return {"workload": "2-stack hourglass, 2 classes, 512x512, batch 16, bf16, train fwd + loss + bwd",
                "value": ips, "unit": "img/s", "ms_per_step": ms, "steps": steps, "warmup": warmup, "launch": launch,
                "eager_img_s": B / ms_eager * 1e3,
                "step_frac_of_conv_roofline": ips * FLOPS_PER_IMG_TRAIN[2] / (peaks["bf16_tflops_sustained"] * 1e12)}
    except Exception as exc:
        return {"error": f"{type(exc).__name__}: {str(exc)[:200]}"}
    finally:
        net = None
        torch.cuda.empty_cache()


def run_ours(args):
    import torch
    import torch.distributed as dist
    from real_time_helmet_detection_b200 import _lib
    from real_time_helmet_detection_b200.hourglass import StackedHourglass
    from real_time_helmet_detection_b200.loss import LossCalculator
    from real_time_helmet_detection_b200.parallel import attach_flat_allreduce, broadcast_parameters
    from real_time_helmet_detection_b200.train import train_step, DevicePrefetcher

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.pop("NCCL_DEBUG", None)                   # NCCL prints its version banner to stdout at any debug level: keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    S, size, B = args.num_stack, args.imsize, args.batch
    peaks = load_peaks()

    torch.manual_seed(777)
    net = StackedHourglass(S, 128, 6).to(dev).train()
    crit = LossCalculator(1.0, 1.0, 0.1, 2.0, 4.0).to(dev)
    hook = None
    if world > 1:
        broadcast_parameters(net)
        hook = attach_flat_allreduce(net, overlap=not args.no_overlap)
    image_h, gts_h = synthetic_batch(torch, B, size, rank)
    image_d, gts_d = image_h.to(dev), [g.to(dev) for g in gts_h]
    image_p, gts_p = image_h.pin_memory(), [g.pin_memory() for g in gts_h]

    def step_device():
        for p in net.parameters():
            p.grad = None
        return train_step(net, crit, image_d, *gts_d)

    def host_batches():                                         # an endless "dataloader" of pinned host batches
        while True:
            yield (image_p, *gts_p)

    loader = DevicePrefetcher(host_batches(), dev)              # H2D of step i+1 overlaps the compute of step i

    # The loss of every step is read back to the host (4 bytes, D2H) through a pinned double buffer with a one-step
    # lag - the copy of step i is issued in step i and consumed in step i+1 - so the host never stalls the device.
    loss_host = [torch.zeros(1).pin_memory() for _ in range(2)]
    loss_evt = [torch.cuda.Event() for _ in range(2)]
    e2e_state = {"i": 0, "last": None}

    def step_e2e():
        i = e2e_state["i"]
        g = e2e_state.get("graphed")
        if g is not None:
            # double-buffered graph replay: this step's batch was staged (pinned host -> the graph's idle input set, on a
            # copy stream) during the previous step; stage the next one now, then replay
            if i == 0:
                g.stage((image_p, *gts_p))
            loss = g.run(log=False)
            g.stage((image_p, *gts_p))                          # H2D of step i+1 overlaps the compute of step i
        else:
            for p in net.parameters():
                p.grad = None
            batch = next(loader)                                # this step's inputs, copied from pinned host memory
            loss = train_step(net, crit, *batch)
        loss_host[i & 1].copy_(loss.reshape(1), non_blocking=True)
        loss_evt[i & 1].record()
        if i > 0:
            loss_evt[(i - 1) & 1].synchronize()
            e2e_state["last"] = float(loss_host[(i - 1) & 1])   # the previous step's loss, on the host
        e2e_state["i"] = i + 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        l0 = _lib.lib().hd_launch_count()
        t0 = time.time()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        t1 = time.time()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, _lib.lib().hd_launch_count() - l0, t0, t1

    warmup = max(3, args.warmup)
    for _ in range(warmup):
        step_device()
    # The measured step is the same iteration replayed from a CUDA graph (train.GraphedTrainStep: one cudaGraphLaunch
    # instead of ~260 kernel launches + ~100 cross-stream event edges; identical kernels and work; for N > 1 the NCCL
    # exchange inside the backward node is part of the recording). The eager step is timed as well and reported beside it.
    graphed, graph_launches, eager = None, 0, None
    if not args.no_graph:
        try:
            from real_time_helmet_detection_b200.train import GraphedTrainStep
            graphed = GraphedTrainStep(net, crit, image_d, *gts_d, warmup=0, buffers=2)
            graph_launches = graphed.kernels_per_replay                 # library kernels recorded into one graph = per replay
        except Exception as exc:                                        # capture not possible: stay eager, say so
            graphed, eager = None, {"graph_error": f"{type(exc).__name__}: {str(exc)[:160]}"}
    if world > 1:                                # every rank must take the same path
        ok = torch.tensor([1 if graphed is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok) == 0:
            graphed = None
    if graphed is not None:
        ms_eager, launches_eager, _, _ = timed(step_device, args.steps)
        eager = {"value": world * B * args.steps / (ms_eager * 1e-3), "unit": "img/s", "ms_per_step": ms_eager / args.steps,
                 "gpu_launches": int(launches_eager), "what": "the same step launched kernel by kernel (no CUDA graph)"}

        def step_measured():                             # inputs resident in HBM: the graph's own static tensors
            return graphed(*graphed.static_in, log=False)
        for _ in range(3):
            step_measured()
    else:
        step_measured = step_device
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ms, launches, t0, t1 = timed(step_measured, args.steps)
    if graphed is not None:
        launches = graph_launches * args.steps          # replayed kernels (hd_launch_count only sees the capture)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    e2e_state["graphed"] = graphed
    for _ in range(max(args.warmup, 5)):    # the prefetch stream's allocator blocks need a few steps to settle
        step_e2e()
    ms_e2e, _, _, _ = timed(step_e2e, args.steps)
    crit.log = {k: [] for k in ("hm", "offset", "size", "total")}

    # Same step fed the way a dataloader with the device-side collate would feed it: uint8 HWC
    # images + box lists on the host -> DeviceCollate (H2D of 25 MB instead of 115 MB, normalise + GT-encode kernels)
    # -> train_step -> lagged loss read-back. Informational: the contract's `e2e` is the float-tensor path above.
    from real_time_helmet_detection_b200.data import DeviceCollate
    from real_time_helmet_detection_b200.synthetic import synthetic_boxes
    import numpy as np
    rs = np.random.RandomState(rank)
    imgs_u8 = [rs.randint(0, 256, (size, size, 3)).astype(np.uint8) for _ in range(B)]
    bl = synthetic_boxes(B, imsize=size)
    bbs, ids = [b[0] for b in bl], [b[1] for b in bl]
    collate = DeviceCollate(dev, num_cls=2, max_boxes=8)
    u8_state = {"i": 0}

    def u8_batches():                                           # the "dataloader": collate runs under the prefetch stream
        while True:
            yield collate(imgs_u8, bbs, ids)

    loader_u8 = DevicePrefetcher(u8_batches(), dev)

    def step_u8():
        i = u8_state["i"]
        for p in net.parameters():
            p.grad = None
        loss = train_step(net, crit, *next(loader_u8))
        loss_host[i & 1].copy_(loss.reshape(1), non_blocking=True)
        loss_evt[i & 1].record()
        if i > 0:
            loss_evt[(i - 1) & 1].synchronize()
            float(loss_host[(i - 1) & 1])
        u8_state["i"] = i + 1

    for _ in range(max(args.warmup, 5)):
        step_u8()
    ms_u8, _, _, _ = timed(step_u8, args.steps)
    crit.log = {k: [] for k in ("hm", "offset", "size", "total")}

    value = world * B * args.steps / (ms * 1e-3)
    e2e_value = world * B * args.steps / (ms_e2e * 1e-3)
    h2d = image_h.numel() * 4 + sum(g.numel() * 4 for g in gts_h)

    if rank == 0:
        roof = dominant_kernel_roofline(torch, dev, peaks)
        per_gpu = value / world
        step_frac = per_gpu * FLOPS_PER_IMG_TRAIN.get(S, 0.0) / (peaks["bf16_tflops_sustained"] * 1e12)
        roof["step_frac_of_conv_roofline"] = step_frac
        roof["step_peak"] = f"{peaks['bf16_tflops_sustained']} TFLOP/s sustained ({peaks['source']})"
        dec = decode_latency(torch, dev, S=1)
        infer = inference_latency(torch, dev, S=1)
        cpu = None
        lib = cfg3 = None
        if world == 1 and not args.no_cpu_baseline:
            # the reference path on this box's host cores, in the same run: a bounded batch-2 sample of the
            # metric's own workload (512x512 train fwd+loss+bwd), config 1 exactly (128x128, batch 2, with / without
            # Adam), and the decode of config 5
            ips, dt, cores, kind = time_cpu(torch, S, size, 2, 6, 1)
            arm1 = CpuArm(torch, 1)
            c1 = time_cpu(torch, 1, 128, 2, 20, 3, adam=False, arm=arm1)
            c1a = time_cpu(torch, 1, 128, 2, 20, 3, adam=True, arm=arm1)
            what = ("unmodified reference modules from baseline/_ref (hourglass.py, loss.py, optim.py; train.py:99-136 "
                    "loop body with the out-of-place squeeze)" if kind == "reference"
                    else "oracle port (fp32 PyTorch CPU restatement of the reference path; reference not staged)")
            cpu = {"value": ips, "unit": "img/s", "cores": cores, "kind": kind, "cpu": cpu_name(),
                   "sample": f"{what}, 6 timed steps of a batch-2 sample of the 512x512 train fwd+loss+bwd workload, "
                             f"{cores} threads",
                   "config1": {"workload": "1-stack, 2 classes, 128x128, batch 2, fp32 CPU, 20 timed steps",
                               "fwd_loss_bwd": {"img_s": c1[0], "ms_per_step": c1[1] * 1e3},
                               "fwd_loss_bwd_adam": {"img_s": c1a[0], "ms_per_step": c1a[1] * 1e3}}}
            dcpu = cpu_decode_us(torch, S=1)
            dec["reference_cpu_us"] = dcpu["us"]
            dec["reference_cpu"] = dcpu
            cpu["decode_config5_us"] = dcpu["us"]
        if world == 1 and not args.no_library_bar:
            lib = library_bar(torch, dev, S, B, size)
            if "best_img_s" in lib:
                lib["ours_over_best_library"] = per_gpu / lib["best_img_s"]
        if world == 1 and S == 1 and not args.no_config3:
            cfg3 = config3_line(torch, dev, peaks)
        line = {"metric": METRIC, "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps,
                "warmup": warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"{S}-stack hourglass, 2 classes, {size}x{size}, batch {B}/GPU, bf16, "
                                       "train fwd + fused focal/L1 loss + bwd" + (", NCCL grad all-reduce (2 overlapped buckets)" if world > 1 else ""),
                           "global_batch": B * world, "parallelism": f"dp{world}",
                           "l2": "no explicit flush: ~10 GB of activations per step >> 126 MB L2"},
                "e2e": {"value": e2e_value, "unit": "img/s", "h2d_bytes_per_step": h2d * world,
                        "d2h_bytes_per_step": 4 * world, "ms_per_step": ms_e2e / args.steps,
                        "how": ("train.GraphedTrainStep.stage / run: pinned host batch -> H2D into the idle one of two static "
                                "input sets on a copy stream (overlapping the previous step) -> CUDA-graph replay of the "
                                "iteration" if graphed is not None else
                                "train.train_step on pinned host batches via train.DevicePrefetcher (H2D of step i+1 "
                                "overlaps step i)") + "; every step's loss is copied D2H and read on the host one step later"},
                "e2e_device_collate": {"value": world * B * args.steps / (ms_u8 * 1e-3), "unit": "img/s",
                                       "h2d_bytes_per_step": collate.h2d_bytes * world, "ms_per_step": ms_u8 / args.steps,
                                       "how": "uint8 HWC images + padded box lists staged in pinned memory -> data.DeviceCollate "
                                              "(hd_normalize_u8 + hd_encode_targets on the device, issued one step ahead on "
                                              "the prefetch stream) -> train.train_step; the host-side staging copy of the "
                                              "images is inside the timed region"},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "decode": dec,
                "inference_b1": infer}
        line["config"]["launch"] = ("CUDA graph replay of the whole iteration (train.GraphedTrainStep)" if graphed is not None
                                    else "eager (kernel by kernel)")
        if eager is not None:
            line["eager"] = eager
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if lib is not None:
            line["library_bar"] = lib
        if cfg3 is not None:
            line["config3"] = cfg3
        if hook is not None:
            line["allreduce"] = {"collectives_per_step": hook.calls / max(1, hook.steps), "bytes_per_step": hook.elements * 4,
                                 "overlap": bool(hook.overlap), "op": "ncclAvg (no scaling kernel)",
                                 "how": ("two buckets: everything but the 256x256 level (96 % of the buffer) on a communication stream "
                                         "under that level's backward, the rest after the last wgrad"
                                         if hook.overlap else "one flat all-reduce after the backward pass")}
        print(json.dumps(line), flush=True)
    if world > 1:
        # Teardown. CUDA graphs that recorded NCCL collectives keep communicator resources alive, and destroying the
        # process group under them was seen to hang (round 2: the JSON line was out, then the run sat until its timeout).
        # So: drop the graphs first, and bound the whole teardown - the measurement is complete at this point.
        import gc
        sys.stdout.flush()
        graphed = None
        e2e_state.clear()
        gc.collect()
        torch.cuda.synchronize()
        threading.Timer(20.0, lambda: os._exit(0)).start()
        try:
            dist.barrier()
            dist.destroy_process_group()
        finally:
            sys.stdout.flush()
            os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--num-stack", type=int, default=1)
    ap.add_argument("--imsize", type=int, default=512)
    ap.add_argument("--batch", type=int, default=32, help="per-GPU batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-bar", action="store_true")
    ap.add_argument("--no-config3", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="N=1: time the eager step only (no CUDA-graph replay)")
    ap.add_argument("--no-overlap", action="store_true", help="N>1: one flat all-reduce after backward (round-1 behaviour)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
