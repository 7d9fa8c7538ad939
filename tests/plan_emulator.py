"""Torch emulator of a centernet_lightning_b200.plan.Plan (CPU).  Test helper.

Executes the fused-op list with ATen ops so the host-side lowering (BN folding, head
fusion, residual / upsample-add wiring, channel offsets) can be checked on CPU against
the forward oracle, and so the effect of 16-bit operand storage can be studied
(``act_dtype``: activations and weights are rounded to that dtype between ops,
products accumulate in fp32 - exactly what tcgen05 kind::f16 does).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F


def _q(x: torch.Tensor, dtype: Optional[torch.dtype], split: bool) -> torch.Tensor:
    if dtype is None:
        return x
    hi = x.to(dtype).float()
    if split:                       # hi + lo pair keeps ~2x the mantissa bits
        hi = hi + (x - hi).to(dtype).float()
    return hi


@torch.no_grad()
def run_plan(plan, image: torch.Tensor, act_dtype: Optional[torch.dtype] = None, split: bool = False,
             conv_dtype: torch.dtype = torch.float32, extra: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
    bufs: Dict[str, torch.Tensor] = {"image": image}
    bufs.update(extra or {})
    for op in plan.ops:
        if op.kind == "fuse":                       # dst = sum_i scale_i * src_i, the last source resized first
            terms = []
            for j, (name, sc) in enumerate(zip(op.srcs, op.scales)):
                t = bufs[name]
                if j == len(op.srcs) - 1 and op.resize == 1:
                    t = F.interpolate(t, scale_factor=2.0, mode="nearest")
                elif j == len(op.srcs) - 1 and op.resize == 2:
                    t = F.max_pool2d(t, 2, 2)
                terms.append(t * sc)
            bufs[op.dst] = _q(sum(terms), act_dtype, split)
            continue
        x = bufs[op.src]
        if op.kind == "dw":                         # depthwise 3x3, pad 1
            w = _q(op.weight, act_dtype, split)
            y = F.conv2d(x, w, op.bias, op.stride, 1, groups=x.shape[1])
            y = F.relu6(y) if op.relu == 2 else (F.relu(y) if op.relu else y)
            bufs[op.dst] = _q(y, act_dtype, split)
            continue
        if op.kind not in ("stem", "stem3x3"):
            x = x[:, op.src_c_off:op.src_c_off + op.cin]
        w = _q(op.weight, act_dtype, split)
        if op.kh:                                   # explicit window with top/left padding (transposed-conv phases)
            xp = F.pad(x, (op.pad_w, op.kw - 1 - op.pad_w, op.pad_h, op.kh - 1 - op.pad_h))
            y = F.conv2d(xp.to(conv_dtype), w.to(conv_dtype), None, op.stride, 0).float()
        else:
            y = F.conv2d(x.to(conv_dtype), w.to(conv_dtype), None, op.stride, op.pad).float()
        y = y + op.bias.view(1, -1, 1, 1)
        if op.residual is not None:
            r = bufs[op.residual]
            if op.residual_up == 2:
                r = F.interpolate(r, scale_factor=2.0, mode="nearest")
            y = y + r
        if op.relu == 2:
            y = F.relu6(y)
        elif op.relu:
            y = F.relu(y)
        if op.kind == "stem":
            y = F.max_pool2d(y, 3, 2, 1)
        dst = plan.buffers[op.dst]
        if not dst.fp32_nchw:
            y = _q(y, act_dtype, split)
        up = op.dst_up
        if op.dst not in bufs:
            n, _, h, wd = y.shape
            bufs[op.dst] = torch.zeros((n, dst.channels, h * up, wd * up))
        if up == 2 and op.dst_phase < 0:            # conv + nearest x2
            bufs[op.dst][:, op.dst_c_off:op.dst_c_off + op.cout] = F.interpolate(y, scale_factor=2.0, mode="nearest")
        elif up == 2:                               # one sub-pixel phase
            py, px = op.dst_phase >> 1, op.dst_phase & 1
            bufs[op.dst][:, op.dst_c_off:op.dst_c_off + op.cout, py::2, px::2] = y
        else:
            bufs[op.dst][:, op.dst_c_off:op.dst_c_off + op.cout] = y
    return {h: bufs[b] for h, b in plan.outputs.items()}
