"""Two warm batch-1 eval forwards + decode for ncu: `python tools/profile_infer.py [B]`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from real_time_helmet_detection_b200.hourglass import StackedHourglass
from real_time_helmet_detection_b200.evaluate import Prediction
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = StackedHourglass(1, 128, 6).to(dev).eval()
pred = Prediction(net, 100, 4, 0.2, "nms", 0.2)
x = torch.randn(B, 3, 512, 512, device=dev)
for _ in range(3):
    pred(x)
torch.cuda.synchronize()
print("done")
