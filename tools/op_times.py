"""Per-op device times of the batch-32 / 512x512 engine: python tools/op_times.py [precision] [reps] [name-filter]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centernet_lightning_b200.model import CenterNet  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "split"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
flt = sys.argv[3] if len(sys.argv) > 3 else ""
dev = torch.device("cuda:0")
net = CenterNet(80, box_multiplier=16.0, precision=prec).init_synthetic_(0).to(dev)
x = torch.rand((32, 3, 512, 512), device=dev)
eng = net.model.engine_for(x)
for _ in range(3):
    eng.forward(x)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(reps):
    eng.forward(x)
e.record()
torch.cuda.synchronize()
print(f"lib={os.environ.get('CNL_LIB', 'default')} precision={prec} forward {s.elapsed_time(e) / reps:.4f} ms", flush=True)
tot = 0.0
for i, op in enumerate(eng.plan.ops):
    if flt and flt not in op.name:
        continue
    eng.forward(x, i, i + 1)
    torch.cuda.synchronize()
    s.record()
    for _ in range(reps):
        eng.forward(x, i, i + 1)
    e.record()
    torch.cuda.synchronize()
    t = s.elapsed_time(e) / reps
    tot += t
    hw = 512 // eng.plan.buffers[op.dst].stride
    gmac = (256 * 256 if op.kind == "stem" else hw * hw) * op.macs_per_out_pixel * 32 / 1e9
    print(f"  {op.name:34s} {t:8.4f} ms  {2 * gmac / t:8.1f} TFLOP/s(alg)", flush=True)
print(f"  sum {tot:.4f} ms")
