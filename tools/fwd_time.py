"""Whole-step time of the headline configuration (CUDA-graph replays of detect(), resident inputs): the number A/B experiments
on kernel-selection knobs (environment variables, read once per process) should compare.  python tools/fwd_time.py [reps] [batch] [size] [backbone] [neck]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centernet_lightning_b200.model import CenterNet  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 32
size = int(sys.argv[3]) if len(sys.argv) > 3 else 512
dev = torch.device("cuda:0")
backbone = sys.argv[4] if len(sys.argv) > 4 else "resnet34"
neck = sys.argv[5] if len(sys.argv) > 5 else "FPN"
net = CenterNet(80, backbone, neck=neck, box_multiplier=16.0).init_synthetic_(0).to(dev)
xs = [torch.rand((batch, 3, size, size), device=dev) for _ in range(2)]
for i in range(4):
    net.detect(xs[i & 1], static_input=True)
torch.cuda.synchronize()
time.sleep(0.5)
best = []
for rnd in range(3):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(reps):
        net.detect(xs[i & 1], static_input=True)
    e.record()
    torch.cuda.synchronize()
    best.append(s.elapsed_time(e) / reps)
knobs = {k: v for k, v in os.environ.items() if k.startswith("CNL_")}
print(f"step ms (3 rounds of {reps}): " + " ".join(f"{t:.4f}" for t in best) + f"  images/s {batch / min(best) * 1e3:.1f}  {backbone}+{neck} batch {batch} @{size}  knobs {knobs}", flush=True)
