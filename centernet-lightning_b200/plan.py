"""Execution plan for the CenterNet forward graph: a flat list of fused conv launches.

The reference runs ``GenericModel.forward`` (reference models/meta.py:41-47) as ~50
cuDNN convolutions plus separate BatchNorm / ReLU / add / upsample / max-pool kernels.
Here the same graph is lowered ONCE, at model-build time, into a list of ``ConvOp``
records that the sm_100a engine (csrc/cnl_conv.cu) executes back to back:

* eval-mode BatchNorm is folded into the conv weights and a per-channel bias
  (SURVEY Appendix B8), computed in float64 on the host;
* ReLU, the ResNet identity/downsample add, and the FPN "lateral + nearest-upsample(x)"
  add become epilogue flags of the producing conv;
* the first 3x3 conv of every head reads the same neck output, so the heads' first
  layers are concatenated into one launch with Cout = heads x width;
* the stem 7x7/2 conv + BN + ReLU + 3x3/2 max-pool is one special op.

This module is pure host logic (numpy/torch on CPU, no CUDA) so it is covered by the
``-m "not gpu"`` tests through a torch emulator of the op list (tests/plan_emulator.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

RESNET_DEPTHS = {"resnet18": (2, 2, 2, 2), "resnet34": (3, 4, 6, 3), "resnet50": (3, 4, 6, 3)}
RESNET_WIDTHS = (64, 128, 256, 512)


def resnet_out_channels(backbone: str) -> Tuple[int, ...]:
    """Channels of the four stage outputs: BasicBlock trunks keep the width, Bottleneck trunks (resnet50) expand x4."""
    return tuple(w * (4 if backbone == "resnet50" else 1) for w in RESNET_WIDTHS)


@dataclass
class Buffer:
    """One activation tensor in HBM, NHWC (channels innermost), 2 bytes per element unless fp32_nchw."""
    name: str
    channels: int
    stride: int                 # spatial down-scale relative to the input image
    fp32_nchw: bool = False     # head outputs are handed to decode / the caller as fp32 NCHW


@dataclass
class ConvOp:
    name: str
    kind: str                   # "stem" | "conv"
    src: str
    dst: str
    cin: int
    cout: int                   # real output channels
    ksize: int
    stride: int
    pad: int
    weight: torch.Tensor        # (cout, cin, k, k) float32, BN folded
    bias: torch.Tensor          # (cout,) float32, BN folded
    relu: bool
    src_c_off: int = 0          # first input channel inside src (grouped head towers)
    dst_c_off: int = 0
    residual: Optional[str] = None      # buffer added before ReLU
    residual_up: int = 1                # 2 = residual is the half-resolution map, nearest-upsampled (FPN)
    kh: int = 0                         # explicit (non-square) window with its own top/left padding; 0 = ksize x ksize
    kw: int = 0
    pad_h: int = 0
    pad_w: int = 0
    dst_up: int = 1                     # 2 = dst has twice the conv's output resolution
    dst_phase: int = -1                 # dst_up == 2: -1 = fill each 2x2 block (nearest x2), 0..3 = one sub-pixel (py*2+px)

    @property
    def window(self) -> Tuple[int, int]:
        return (self.kh or self.ksize, self.kw or self.ksize)

    @property
    def macs_per_out_pixel(self) -> int:
        kh, kw = self.window
        return self.cin * self.cout * kh * kw


@dataclass
class Plan:
    buffers: Dict[str, Buffer] = field(default_factory=dict)
    ops: List[ConvOp] = field(default_factory=list)
    outputs: Dict[str, str] = field(default_factory=dict)   # head name -> buffer name
    model_stride: int = 4

    def add_buffer(self, name, channels, stride, fp32_nchw=False) -> str:
        self.buffers[name] = Buffer(name, channels, stride, fp32_nchw)
        return name

    def macs(self, height: int, width: int) -> int:
        """Algorithmic conv MACs per image (bias/BN/ReLU/add/pool excluded), SURVEY 8d."""
        total = 0
        for op in self.ops:
            s = self.buffers[op.dst].stride
            if op.kind == "stem":
                s = 2                                   # conv output is at stride 2, pooled to 4
            s *= op.dst_up                              # upsampling stores: the conv runs at half the dst resolution
            total += (height // s) * (width // s) * op.macs_per_out_pixel
        return total


def fold_bn(weight: torch.Tensor, bn: Optional[Dict[str, torch.Tensor]], conv_bias: Optional[torch.Tensor] = None,
            eps: float = 1e-5) -> Tuple[torch.Tensor, torch.Tensor]:
    """w' = w * gamma / sqrt(var + eps);  b' = beta + (b - mean) * gamma / sqrt(var + eps)."""
    w = weight.detach().double()
    cout = w.shape[0]
    b = conv_bias.detach().double() if conv_bias is not None else torch.zeros(cout, dtype=torch.float64)
    if bn is not None:
        scale = bn["weight"].detach().double() / torch.sqrt(bn["running_var"].detach().double() + eps)
        w = w * scale.view(-1, 1, 1, 1)
        b = bn["bias"].detach().double() + (b - bn["running_mean"].detach().double()) * scale
    return w.float().contiguous(), b.float().contiguous()


def _bn(sd: Dict[str, torch.Tensor], prefix: str) -> Dict[str, torch.Tensor]:
    return {k: sd[f"{prefix}.{k}"] for k in ("weight", "bias", "running_mean", "running_var")}


def lower_conv_transpose(name: str, src: str, dst: str, weight: torch.Tensor, bn: Optional[Dict[str, torch.Tensor]],
                         relu: bool = True) -> List[ConvOp]:
    """ConvTranspose2d(C_in, C_out, k, stride=2, padding=p, output_padding=k%2, bias=False) [+ BN + ReLU] as the
    reference builds it (models/layers.py:86-96: k=3 -> p=1, output_padding=1; k=4 -> p=1) lowered to four ordinary
    convolutions, one per output sub-pixel (py, px):

        out[2y+py, 2x+px] = sum_{dy,dx} in[y+dy, x+dx] * W[:, :, py+p-2dy, px+p-2dx]     (taps inside the kernel only)

    i.e. a (1..2)x(1..2) correlation window with top/left padding -dy_min/-dx_min whose result lands on sub-pixel
    (py, px) of the double-resolution destination.  ``weight`` is torch's (C_in, C_out, k, k)."""
    cin, cout, k, k2 = weight.shape
    if k != k2 or k not in (3, 4):
        raise ValueError(f"conv_transpose kernel {k}x{k2}: 3 and 4 are implemented (reference models/layers.py:86-96)")
    pad = (k + k % 2) // 2 - 1
    # fold BN (per output channel = dim 1 of the transposed-conv weight)
    w_f, b_f = fold_bn(weight.permute(1, 0, 2, 3).contiguous(), bn)          # (cout, cin, k, k)
    ops = []
    for py in range(2):
        dys = [dy for dy in (-1, 0, 1) if 0 <= py + pad - 2 * dy < k]
        for px in range(2):
            dxs = [dx for dx in (-1, 0, 1) if 0 <= px + pad - 2 * dx < k]
            wp = torch.empty((cout, cin, len(dys), len(dxs)), dtype=torch.float32)
            for r, dy in enumerate(dys):
                for c, dx in enumerate(dxs):
                    wp[:, :, r, c] = w_f[:, :, py + pad - 2 * dy, px + pad - 2 * dx]
            ops.append(ConvOp(f"{name}.phase{py}{px}", "conv", src, dst, cin, cout, 0, 1, 0, wp.contiguous(), b_f, relu=relu,
                              kh=len(dys), kw=len(dxs), pad_h=-dys[0], pad_w=-dxs[0], dst_up=2, dst_phase=py * 2 + px))
    return ops


def build_plan(sd: Dict[str, torch.Tensor], *, backbone: str = "resnet34", neck: str = "FPN",
               head_names: Sequence[str] = ("heatmap", "box_2d"), head_depth: int = 3,
               prefix: str = "") -> Plan:
    """Lower a state dict with the G2 key layout (``backbone.*``, ``neck.*``, ``heads.<h>.block_<i>.*``,
    ``heads.<h>.out_conv.*``; reference models/meta.py:26-28,36-38,92-95) into a Plan."""
    sd = {k[len(prefix):]: v.detach().cpu() for k, v in sd.items() if k.startswith(prefix)}
    p = Plan()
    p.add_buffer("image", 3, 1, fp32_nchw=True)

    # ---- stem: conv7x7/2 + BN + ReLU + maxpool3x3/2 ------------------------------------------
    w, b = fold_bn(sd["backbone.conv1.weight"], _bn(sd, "backbone.bn1"))
    cur = p.add_buffer("stem", 64, 4)
    p.ops.append(ConvOp("backbone.stem", "stem", "image", cur, 3, 64, 7, 2, 3, w, b, relu=True))

    # ---- residual stages ----------------------------------------------------------------------
    feats: List[str] = []
    cin = 64
    bottleneck = backbone == "resnet50"
    out_ch = resnet_out_channels(backbone)
    for li, (width, depth) in enumerate(zip(RESNET_WIDTHS, RESNET_DEPTHS[backbone])):
        s_out = 4 * (2 ** li)
        cout = out_ch[li]
        for bi in range(depth):
            stride = 2 if (bi == 0 and li > 0) else 1
            base = f"backbone.layer{li + 1}.{bi}"
            idt = cur
            if f"{base}.downsample.0.weight" in sd:
                w, b = fold_bn(sd[f"{base}.downsample.0.weight"], _bn(sd, f"{base}.downsample.1"))
                idt = p.add_buffer(f"{base}.idt", cout, s_out)
                p.ops.append(ConvOp(f"{base}.downsample", "conv", cur, idt, cin, cout, 1, stride, 0, w, b, relu=False))
            if bottleneck:
                # torchvision v1.5 Bottleneck: 1x1 reduce -> 3x3 (carries the stride) -> 1x1 expand (+ identity) -> ReLU
                w, b = fold_bn(sd[f"{base}.conv1.weight"], _bn(sd, f"{base}.bn1"))
                m1 = p.add_buffer(f"{base}.mid1", width, s_out // stride)
                p.ops.append(ConvOp(f"{base}.conv1", "conv", cur, m1, cin, width, 1, 1, 0, w, b, relu=True))
                w, b = fold_bn(sd[f"{base}.conv2.weight"], _bn(sd, f"{base}.bn2"))
                m2 = p.add_buffer(f"{base}.mid2", width, s_out)
                p.ops.append(ConvOp(f"{base}.conv2", "conv", m1, m2, width, width, 3, stride, 1, w, b, relu=True))
                w, b = fold_bn(sd[f"{base}.conv3.weight"], _bn(sd, f"{base}.bn3"))
                out = p.add_buffer(f"{base}.out", cout, s_out)
                p.ops.append(ConvOp(f"{base}.conv3", "conv", m2, out, width, cout, 1, 1, 0, w, b, relu=True, residual=idt))
            else:
                w, b = fold_bn(sd[f"{base}.conv1.weight"], _bn(sd, f"{base}.bn1"))
                mid = p.add_buffer(f"{base}.mid", width, s_out)
                p.ops.append(ConvOp(f"{base}.conv1", "conv", cur, mid, cin, width, 3, stride, 1, w, b, relu=True))
                w, b = fold_bn(sd[f"{base}.conv2.weight"], _bn(sd, f"{base}.bn2"))
                out = p.add_buffer(f"{base}.out", width, s_out)
                p.ops.append(ConvOp(f"{base}.conv2", "conv", mid, out, width, width, 3, 1, 1, w, b, relu=True, residual=idt))
            cur, cin = out, cout
        feats.append(cur)

    # ---- neck -----------------------------------------------------------------------------------
    if neck == "FPN":
        d = sd["neck.lateral.0.weight"].shape[0]
        w, b = fold_bn(sd["neck.lateral.3.weight"], None, sd["neck.lateral.3.bias"])
        x = p.add_buffer("neck.p5", d, 32)
        p.ops.append(ConvOp("neck.lateral.3", "conv", feats[3], x, out_ch[3], d, 1, 1, 0, w, b, relu=False))
        for i in (2, 1, 0):
            s = 4 * (2 ** i)
            w, b = fold_bn(sd[f"neck.lateral.{i}.weight"], None, sd[f"neck.lateral.{i}.bias"])
            fused = p.add_buffer(f"neck.sum{i}", d, s)
            p.ops.append(ConvOp(f"neck.lateral.{i}", "conv", feats[i], fused, out_ch[i], d, 1, 1, 0, w, b,
                                relu=False, residual=x, residual_up=2))
            w, b = fold_bn(sd[f"neck.output.{i}.conv.weight"], _bn(sd, f"neck.output.{i}.bn"))
            x = p.add_buffer(f"neck.out{i}", d, s)
            p.ops.append(ConvOp(f"neck.output.{i}", "conv", fused, x, d, d, 3, 1, 1, w, b, relu=True))
        neck_out, neck_c, neck_stride = x, d, 4
    elif neck in ("simple", "SimpleNeck"):
        # G1 "simple" neck (reference configs/base_resnet34.yaml:7-11, models/layers.py:71-99): on C5 only,
        # n x [conv3x3-BN-ReLU -> x2 upsample]; the upsample is nearest (fused into the conv's store: every output pixel
        # is written to its 2x2 block) or ConvTranspose2d(k, stride 2)-BN-ReLU (four sub-pixel phase convs).
        x, cin_n, s = feats[3], out_ch[3], 32
        i = 0
        while f"neck.blocks.{i}.conv.weight" in sd:
            w, b = fold_bn(sd[f"neck.blocks.{i}.conv.weight"], _bn(sd, f"neck.blocks.{i}.bn"))
            c = w.shape[0]
            if f"neck.up.{i}.0.weight" in sd:
                y = p.add_buffer(f"neck.conv{i}", c, s)
                p.ops.append(ConvOp(f"neck.blocks.{i}", "conv", x, y, cin_n, c, 3, 1, 1, w, b, relu=True))
                up = p.add_buffer(f"neck.up{i}", c, s // 2)
                for op in lower_conv_transpose(f"neck.up.{i}", y, up, sd[f"neck.up.{i}.0.weight"], _bn(sd, f"neck.up.{i}.1")):
                    p.ops.append(op)
            else:
                up = p.add_buffer(f"neck.up{i}", c, s // 2)
                p.ops.append(ConvOp(f"neck.blocks.{i}", "conv", x, up, cin_n, c, 3, 1, 1, w, b, relu=True, dst_up=2, dst_phase=-1))
            x, cin_n, s = up, c, s // 2
            i += 1
        if i == 0:
            raise ValueError("simple neck: no neck.blocks.* parameters in the state dict")
        neck_out, neck_c, neck_stride = x, cin_n, s
    else:
        raise ValueError(f"neck {neck!r} is not lowered by the sm_100a engine (FPN and simple; SURVEY 8a F2, 8f rank 4)")
    p.model_stride = neck_stride

    # ---- heads: first tower layers fused across heads ------------------------------------------
    width = sd[f"heads.{head_names[0]}.out_conv.weight"].shape[1]
    nh = len(head_names)
    if head_depth >= 1:
        ws, bs = [], []
        for h in head_names:
            w, b = fold_bn(sd[f"heads.{h}.block_1.conv.weight"], _bn(sd, f"heads.{h}.block_1.bn"))
            ws.append(w); bs.append(b)
        t = p.add_buffer("heads.t1", width * nh, neck_stride)
        p.ops.append(ConvOp("heads.block_1", "conv", neck_out, t, neck_c, width * nh, 3, 1, 1,
                            torch.cat(ws), torch.cat(bs), relu=True))
        for li in range(2, head_depth + 1):
            t2 = p.add_buffer(f"heads.t{li}", width * nh, neck_stride)
            for hi, h in enumerate(head_names):
                w, b = fold_bn(sd[f"heads.{h}.block_{li}.conv.weight"], _bn(sd, f"heads.{h}.block_{li}.bn"))
                p.ops.append(ConvOp(f"heads.{h}.block_{li}", "conv", t, t2, width, width, 3, 1, 1, w, b, relu=True,
                                    src_c_off=hi * width, dst_c_off=hi * width))
            t = t2
        tower, tower_c = t, width
    else:
        tower, tower_c = neck_out, neck_c
    for hi, h in enumerate(head_names):
        w = sd[f"heads.{h}.out_conv.weight"].float().contiguous()
        b = sd[f"heads.{h}.out_conv.bias"].float().contiguous()
        o = p.add_buffer(f"out.{h}", w.shape[0], neck_stride, fp32_nchw=True)
        p.ops.append(ConvOp(f"heads.{h}.out_conv", "conv", tower, o, tower_c, w.shape[0], 1, 1, 0, w, b, relu=False,
                            src_c_off=(hi * width if head_depth >= 1 else 0)))
        p.outputs[h] = o
    return p
