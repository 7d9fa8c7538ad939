"""Run one op of the batch-32 / 512x512 engine repeatedly (for ncu captures): python tools/run_op.py <op-name> [reps] [precision]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centernet_lightning_b200.model import CenterNet  # noqa: E402

name = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
prec = sys.argv[3] if len(sys.argv) > 3 else "split"
dev = torch.device("cuda:0")
net = CenterNet(80, box_multiplier=16.0, precision=prec).init_synthetic_(0).to(dev)
x = torch.rand((32, 3, 512, 512), device=dev)
eng = net.model.engine_for(x)
eng.forward(x)                       # populate every activation buffer once
torch.cuda.synchronize()
idx = [op.name for op in eng.plan.ops].index(name)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(reps):
    eng.forward(x, idx, idx + 1)
e.record()
torch.cuda.synchronize()
print(name, prec, "ms/launch", s.elapsed_time(e) / reps)
