"""CPU study: which parity bar could a scheme with FEWER tensor-core passes on some layers meet?

The engine carries activations and weights as fp16 hi+lo pairs and issues three passes per product (hi*hi, hi*lo, lo*hi).
Dropping `lo*hi` for an op is the same as rounding its INPUT ACTIVATIONS to one fp16 ("2w": weights stay split), dropping
`hi*lo` rounds its WEIGHTS ("2a").  This script emulates exactly that operand rounding per op (float64 accumulation, so only
the operand precision is studied, not the accumulator's) on the 50-op ResNet-34+FPN plan and reports the head-map error
against the float64 reference.  python tools/precision_mixed_cpu.py [size]"""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from centernet_lightning_b200 import plan as P  # noqa: E402
from oracle import spec_model  # noqa: E402


def q16(x, split):
    hi = x.to(torch.float16).double()
    if split:
        hi = hi + (x - hi).to(torch.float16).double()
    return hi


@torch.no_grad()
def run(plan, image, modes):
    """modes[op.name] in {"3", "2w", "2a", "1"}; storage is always the hi+lo pair (except exact = no rounding at all)."""
    bufs = {"image": image.double()}
    for op in plan.ops:
        mode = modes.get(op.name, modes.get("*", "3"))
        x = bufs[op.src]
        if op.kind != "stem":
            x = x[:, op.src_c_off:op.src_c_off + op.cin]
        w = op.weight.double()
        if mode != "exact":
            x = q16(x, split=mode in ("3", "2a"))
            w = q16(w, split=mode in ("3", "2w"))
        y = F.conv2d(x, w, None, op.stride, op.pad) + op.bias.double().view(1, -1, 1, 1)
        if op.residual is not None:
            r = bufs[op.residual]
            if op.residual_up == 2:
                r = F.interpolate(r, scale_factor=2.0, mode="nearest")
            y = y + r
        if op.relu:
            y = F.relu(y)
        if op.kind == "stem":
            y = F.max_pool2d(y, 3, 2, 1)
        if op.dst not in bufs:
            bufs[op.dst] = torch.zeros((y.shape[0], plan.buffers[op.dst].channels, y.shape[2], y.shape[3]), dtype=torch.float64)
        bufs[op.dst][:, op.dst_c_off:op.dst_c_off + op.cout] = y
    return {h: bufs[b] for h, b in plan.outputs.items()}


if __name__ == "__main__":
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    torch.set_num_threads(os.cpu_count() or 1)
    spec = spec_model.synth_init(spec_model.build_spec_model(80), seed=0)
    plan = P.build_plan(spec.state_dict())
    g = torch.Generator().manual_seed(11)
    x = torch.rand((2, 3, size, size), generator=g)
    ref = run(plan, x, {"*": "exact"})
    tower = [op.name for op in plan.ops if op.name.startswith("heads.") and "block_" in op.name]
    big = tower + ["neck.output.0"]
    macs = {op.name: op.macs_per_out_pixel * (size // (2 if op.kind == "stem" else plan.buffers[op.dst].stride)) ** 2 for op in plan.ops}
    total = sum(macs.values())
    schemes = {
        "3 passes everywhere (shipped)": {"*": "3"},
        "1 pass everywhere (fp16)": {"*": "1"},
        "2 passes everywhere, activations single (drop lo*hi)": {"*": "2w"},
        "2 passes everywhere, weights single (drop hi*lo)": {"*": "2a"},
        "2 passes (drop lo*hi) on the 6 head-tower convs + heads.block_1 + neck.output.0": {**{n: "2w" for n in big}, "*": "3"},
        "2 passes (drop hi*lo) on the same ops": {**{n: "2a" for n in big}, "*": "3"},
        "2 passes (drop lo*hi) on block_3 of both heads only": {"heads.heatmap.block_3": "2w", "heads.box_2d.block_3": "2w", "*": "3"},
        "2 passes (drop lo*hi) on block_2 and block_3 of both heads": {**{n: "2w" for n in tower if "block_1" not in n}, "*": "3"},
    }
    for name, modes in schemes.items():
        out = run(plan, x, modes)
        err = {k: (out[k] - ref[k]).abs().max().item() for k in out}
        mean = {k: (out[k] - ref[k]).abs().mean().item() for k in out}
        passes = sum(macs[op.name] * {"3": 3, "2w": 2, "2a": 2, "1": 1}[modes.get(op.name, modes["*"])] for op in plan.ops) / total
        print(json.dumps({"scheme": name, "avg_passes": round(passes, 3), "roofline_ceiling": round(1 / passes, 3),
                          "max_err": {k: float(f"{v:.3g}") for k, v in err.items()}, "mean_err": {k: float(f"{v:.3g}") for k, v in mean.items()}}), flush=True)
