"""Batch-1 inference latency: network forward (eval) + fused decode/NMS, as the reference's demo measures it
(PytorchToCpp/main.cpp:60-67: second model.forward, wall clock)."""
import os, sys, time, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from real_time_helmet_detection_b200.hourglass import StackedHourglass
from real_time_helmet_detection_b200.evaluate import Prediction
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = StackedHourglass(1, 128, 6).to(dev).eval()
for B, graph in ((1, False), (1, True), (8, False), (8, True)):
    pred = Prediction(net, 100, 4, 0.2, "nms", 0.2, cuda_graph=graph)
    x = torch.randn(B, 3, 512, 512, device=dev)
    for _ in range(5):
        pred(x)
    torch.cuda.synchronize()
    wall = []
    for _ in range(50):
        t0 = time.perf_counter()
        b, c, s = pred(x)          # ends with the count read-back (a sync)
        wall.append(time.perf_counter() - t0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        with torch.no_grad():
            net(x)
    e1.record(); torch.cuda.synchronize()
    print(f"B={B} cuda_graph={graph}: predict wall {statistics.median(wall)*1e3:.3f} ms/batch ({B/statistics.median(wall):.0f} img/s) | forward-only device {e0.elapsed_time(e1)/50:.3f} ms")
