"""One batch-32 / 512x512 forward + decode, eager launches (for per-launch ncu metrics): python tools/run_forward_once.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centernet_lightning_b200.model import CenterNet  # noqa: E402

dev = torch.device("cuda:0")
net = CenterNet(80, box_multiplier=16.0).init_synthetic_(0).to(dev)
x = torch.rand((32, 3, 512, 512), device=dev)
for _ in range(2):
    net.detect(x, use_graph=False)
torch.cuda.synchronize()
print("ops:", " ".join(op.name for op in net.model.engine_for(x).plan.ops))
