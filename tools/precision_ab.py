"""A/B of the engine precision modes: head-map error vs the CPU fp32 oracle (and fp64), and forward time at batch 32."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from centernet_lightning_b200.model import CenterNet  # noqa: E402
from centernet_lightning_b200 import plan as P  # noqa: E402
from oracle import spec_model  # noqa: E402
from plan_emulator import run_plan  # noqa: E402

dev = torch.device("cuda:0")
spec = spec_model.synth_init(spec_model.build_spec_model(80), seed=0)
g = torch.Generator().manual_seed(11)
x = torch.rand((2, 3, 256, 256), generator=g)
with torch.no_grad():
    ref32 = spec(x)
ref64 = run_plan(P.build_plan(spec.state_dict()), x.double(), conv_dtype=torch.float64)
for prec in ("split", "split_fused", "fast"):
    net = CenterNet(80, box_multiplier=16.0, precision=prec)
    net.model.load_state_dict(spec.state_dict())
    net = net.to(dev)
    out = {k: v.cpu() for k, v in net.model(x.to(dev)).items()}
    e32 = {k: (out[k] - ref32[k]).abs().max().item() for k in out}
    e64 = {k: (out[k].double() - ref64[k]).abs().max().item() for k in out}
    m64 = {k: (out[k].double() - ref64[k]).abs().mean().item() for k in out}
    perr = (out["heatmap"].sigmoid() - ref32["heatmap"].sigmoid()).abs().max().item()
    xb = torch.rand((32, 3, 512, 512), device=dev)
    eng = net.model.engine_for(xb)
    for _ in range(2):
        eng.forward(xb)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        eng.forward(xb)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    print(json.dumps(dict(precision=prec, max_err_vs_fp32=e32, max_err_vs_fp64=e64, mean_err_vs_fp64=m64, prob_err=perr,
                          forward_ms_b32=ms, img_s=32 / ms * 1e3)), flush=True)
    if prec == "split":
        # per-op timing of the batch-32 engine
        rows = []
        for i, op in enumerate(eng.plan.ops):
            eng.forward(xb, i, i + 1)
            torch.cuda.synchronize()
            s.record()
            for _ in range(3):
                eng.forward(xb, i, i + 1)
            e.record()
            torch.cuda.synchronize()
            t = s.elapsed_time(e) / 3
            hw = 512 // eng.plan.buffers[op.dst].stride
            gmac = (256 * 256 if op.kind == "stem" else hw * hw) * op.macs_per_out_pixel * 32 / 1e9
            rows.append((op.name, t, gmac, 2 * gmac / t if t else 0))
        tot = sum(r[1] for r in rows)
        print(f"per-op total {tot:.3f} ms", flush=True)
        for name, t, gmac, tf in rows:
            print(f"  {name:34s} {t:8.4f} ms  {gmac:8.2f} GMAC  {tf:8.1f} TFLOP/s(alg)  {100 * t / tot:5.1f}%", flush=True)
    net.invalidate()
    del net, eng
    torch.cuda.empty_cache()
