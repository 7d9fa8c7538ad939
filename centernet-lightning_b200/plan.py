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
    kind: str                   # "stem" | "conv" | "dw" (depthwise 3x3) | "fuse" (weighted sum of maps) | "stem3x3" (3x3/2 from the image)
    src: str
    dst: str
    cin: int
    cout: int                   # real output channels
    ksize: int
    stride: int
    pad: int
    weight: torch.Tensor        # (cout, cin, k, k) float32, BN folded
    bias: torch.Tensor          # (cout,) float32, BN folded
    relu: int                   # 0 / False = linear, 1 / True = ReLU, 2 = ReLU6 (separable convs, MobileNetV2)
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
    # kind == "fuse": dst = sum_i scales[i] * srcs[i]; the LAST source is resized first (reference models/layers.py:138-177)
    srcs: Optional[List[str]] = None
    scales: Optional[List[float]] = None
    resize: int = 0                     # 0 = same size, 1 = nearest x2 up-sample, 2 = max-pool 2x2 / stride 2
    real_macs: Optional[int] = None     # MACs per output pixel without the zero channel padding (None: cin*cout*kh*kw)

    @property
    def window(self) -> Tuple[int, int]:
        return (self.kh or self.ksize, self.kw or self.ksize)

    @property
    def macs_per_out_pixel(self) -> int:
        if self.real_macs is not None:
            return self.real_macs
        if self.kind == "fuse":
            return 0
        kh, kw = self.window
        return (self.cout if self.kind == "dw" else self.cin * self.cout) * kh * kw


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


MOBILENET_V2_OUT = ((3, 24), (6, 32), (13, 96), (17, 320))      # (features index, channels) of the stride 4/8/16/32 maps


def backbone_out_channels(backbone: str) -> Tuple[int, ...]:
    if backbone == "mobilenet_v2":
        return tuple(c for _, c in MOBILENET_V2_OUT)
    return resnet_out_channels(backbone)


def pad64(c: int) -> int:
    """Channel counts of NHWC buffers are multiples of 64 (one 128-byte swizzle row of fp16); narrower layers
    (MobileNetV2: 16 / 24 / 32 / 96 / 144 / 160 channels) are zero-padded - zero weights and biases keep the padding at 0."""
    return (c + 63) // 64 * 64


def _pad_wb(w: torch.Tensor, b: torch.Tensor, cin_pad: int, cout_pad: int) -> Tuple[torch.Tensor, torch.Tensor]:
    cout, cin = w.shape[0], w.shape[1]
    if cin == cin_pad and cout == cout_pad:
        return w, b
    wp = torch.zeros((cout_pad, cin_pad, *w.shape[2:]), dtype=w.dtype)
    wp[:cout, :cin] = w
    bp = torch.zeros((cout_pad,), dtype=b.dtype)
    bp[:cout] = b
    return wp.contiguous(), bp.contiguous()


class _Lowering:
    """Shared helpers of the backbone / neck lowerings: every NHWC buffer carries pad64(channels) channels."""

    def __init__(self, plan: Plan, sd: Dict[str, torch.Tensor]):
        self.p, self.sd = plan, sd

    def conv(self, name: str, src: str, dst: str, w: torch.Tensor, b: torch.Tensor, *, ksize: int, stride: int = 1, relu: int = 0,
             residual: Optional[str] = None, residual_up: int = 1, src_c_off: int = 0, dst_c_off: int = 0, **kw) -> ConvOp:
        """A dense conv whose weights are zero-padded to the (padded) channel counts of its buffers."""
        cout, cin = w.shape[0], w.shape[1]
        dstb = self.p.buffers[dst]
        cin_p = pad64(cin)
        cout_p = cout if dstb.fp32_nchw else pad64(cout)
        wp, bp = _pad_wb(w, b, cin_p, cout_p)
        kh, kw_ = w.shape[2], w.shape[3]
        op = ConvOp(name, "conv", src, dst, cin_p, cout_p, ksize, stride, (ksize - 1) // 2, wp, bp, relu=relu, residual=residual,
                    residual_up=residual_up, src_c_off=src_c_off, dst_c_off=dst_c_off,
                    real_macs=(cin * cout * kh * kw_ if (cin_p, cout_p) != (cin, cout) else None), **kw)
        self.p.ops.append(op)
        return op

    def depthwise(self, name: str, src: str, dst: str, w: torch.Tensor, b: torch.Tensor, *, stride: int, relu: int) -> ConvOp:
        """Depthwise 3x3 conv, pad 1 (weight (C,1,3,3), BN folded)."""
        c = w.shape[0]
        cp = pad64(c)
        wp = torch.zeros((cp, 1, 3, 3), dtype=w.dtype); wp[:c] = w
        bp = torch.zeros((cp,), dtype=b.dtype); bp[:c] = b
        op = ConvOp(name, "dw", src, dst, cp, cp, 3, stride, 1, wp.contiguous(), bp.contiguous(), relu=relu, real_macs=c * 9)
        self.p.ops.append(op)
        return op

    def make_conv(self, name: str, key: str, src: str, dst: str, c_in: int, c_out: int, stride_of_map: int) -> None:
        """reference models/layers.py:40-79 ``make_conv`` (Sequential indices 0..2 normal, 0..5 separable)."""
        sd = self.sd
        if f"{key}.3.weight" in sd:                                  # separable: dw3x3-BN-ReLU6, pw1x1-BN-ReLU6
            w, b = fold_bn(sd[f"{key}.0.weight"], _bn(sd, f"{key}.1"))
            mid = self.p.add_buffer(f"{name}.dw", pad64(c_in), stride_of_map)
            self.depthwise(f"{name}.dw", src, mid, w, b, stride=1, relu=2)
            w, b = fold_bn(sd[f"{key}.3.weight"], _bn(sd, f"{key}.4"))
            self.conv(f"{name}.pw", mid, dst, w, b, ksize=1, relu=2)
        else:
            w, b = fold_bn(sd[f"{key}.0.weight"], _bn(sd, f"{key}.1"))
            self.conv(name, src, dst, w, b, ksize=3, relu=1)

    def fuse(self, name: str, key: str, inputs: Sequence[Tuple[str, int]], out_c: int, stride_out: int, resize: str) -> str:
        """reference models/layers.py:138-177 ``Fuse``: inputs = [(buffer, real channels)], the last one is resized (its
        map has stride_out * 2 for "up", stride_out / 2 for "down").  Returns the output buffer."""
        sd, p = self.sd, self.p
        n = len(inputs)
        if f"{key}.weights" in sd:
            wts = torch.relu(sd[f"{key}.weights"].double())
            scales = [float(x) for x in (wts / (wts.sum() + 1e-6))]
        else:
            scales = [1.0] * n
        terms: List[Tuple[str, float]] = []
        fused_first = None                                           # (w, b) of a projection that can carry the sum in its epilogue
        for i, (buf, c) in enumerate(inputs):
            s_map = stride_out if i < n - 1 else (stride_out * 2 if resize == "up" else stride_out // 2)
            if f"{key}.project.{i}.weight" in sd:                    # 1x1 conv with bias; the fusion weight folds into it
                w = sd[f"{key}.project.{i}.weight"].double() * scales[i]
                b = sd[f"{key}.project.{i}.bias"].double() * scales[i]
                if i == 0 and n == 2 and resize == "up":
                    fused_first = (w.float(), b.float())
                    continue
                t = p.add_buffer(f"{name}.proj{i}", pad64(out_c), s_map)
                self.conv(f"{name}.project.{i}", buf, t, w.float(), b.float(), ksize=1, relu=0)
                terms.append((t, 1.0))
            else:
                terms.append((buf, scales[i]))
        summed = p.add_buffer(f"{name}.sum", pad64(out_c), stride_out)
        if fused_first is not None and terms[0][1] == 1.0:
            # project(x0) + nearest_up(x1): the FPN pattern - the half-resolution map is added in the 1x1 conv's epilogue
            self.conv(f"{name}.project.0", inputs[0][0], summed, *fused_first, ksize=1, relu=0, residual=terms[0][0], residual_up=2)
        else:
            if fused_first is not None:
                t = p.add_buffer(f"{name}.proj0", pad64(out_c), stride_out)
                self.conv(f"{name}.project.0", inputs[0][0], t, *fused_first, ksize=1, relu=0)
                terms.insert(0, (t, 1.0))
            cp = pad64(out_c)
            p.ops.append(ConvOp(f"{name}.fuse", "fuse", terms[0][0], summed, cp, cp, 1, 1, 0, torch.zeros(0), torch.zeros(0), relu=0,
                                srcs=[t for t, _ in terms], scales=[sc for _, sc in terms],
                                resize={"up": 1, "down": 2}[resize]))
        out = p.add_buffer(f"{name}.out", pad64(out_c), stride_out)
        self.make_conv(f"{name}.output_conv", f"{key}.output_conv", summed, out, out_c, out_c, stride_out)
        return out


def _lower_resnet(lw: _Lowering, backbone: str) -> List[str]:
    p, sd = lw.p, lw.sd
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
    return feats


def _lower_mobilenet_v2(lw: _Lowering) -> List[str]:
    """torchvision mobilenet_v2 features[0:18] (keys ``backbone.features.<i>...``): conv3x3/2-BN-ReLU6, then
    InvertedResidual blocks = [1x1 expand-BN-ReLU6] -> depthwise 3x3 (stride s)-BN-ReLU6 -> 1x1 project-BN (linear)
    [+ input when stride 1 and the channel count is unchanged]."""
    p, sd = lw.p, lw.sd
    w, b = fold_bn(sd["backbone.features.0.0.weight"], _bn(sd, "backbone.features.0.1"))
    cur, cin, s_map = p.add_buffer("stem", pad64(32), 2), 32, 2
    wp = torch.zeros((pad64(32), 3, 3, 3)); wp[:32] = w
    bp = torch.zeros((pad64(32),)); bp[:32] = b
    p.ops.append(ConvOp("backbone.features.0", "stem3x3", "image", cur, 3, pad64(32), 3, 2, 1, wp, bp, relu=2, real_macs=27 * 32))
    feats: List[str] = []
    out_at = dict(MOBILENET_V2_OUT)
    i = 1
    while f"backbone.features.{i}.conv.0.0.weight" in sd:
        base = f"backbone.features.{i}"
        x = cur
        j = 0
        if f"{base}.conv.2.weight" in sd and sd[f"{base}.conv.2.weight"].dim() == 4:      # expand-ratio != 1: conv.0 is the 1x1 expansion
            w, b = fold_bn(sd[f"{base}.conv.0.0.weight"], _bn(sd, f"{base}.conv.0.1"))
            hid = w.shape[0]
            e = p.add_buffer(f"{base}.exp", pad64(hid), s_map)
            lw.conv(f"{base}.expand", x, e, w, b, ksize=1, relu=2)
            x, j = e, 1
        wd = sd[f"{base}.conv.{j}.0.weight"]
        hid = wd.shape[0]
        w, b = fold_bn(wd, _bn(sd, f"{base}.conv.{j}.1"))
        wproj = sd[f"{base}.conv.{j + 1}.weight"]
        cout = wproj.shape[0]
        stride = MOBILENET_V2_STRIDES[i]
        s_out = s_map * stride
        d = p.add_buffer(f"{base}.dw", pad64(hid), s_out)
        lw.depthwise(f"{base}.dw", x, d, w, b, stride=stride, relu=2)
        w, b = fold_bn(wproj, _bn(sd, f"{base}.conv.{j + 2}"))
        out = p.add_buffer(f"{base}.out", pad64(cout), s_out)
        lw.conv(f"{base}.project", d, out, w, b, ksize=1, relu=0, residual=(cur if (stride == 1 and cin == cout) else None))
        cur, cin, s_map = out, cout, s_out
        if i in out_at:
            feats.append(cur)
        i += 1
    if len(feats) != 4:
        raise ValueError("mobilenet_v2: expected backbone.features.0 .. 17 in the state dict")
    return feats


# stride of the depthwise conv of torchvision's mobilenet_v2 features[i] (inverted_residual_setting t,c,n,s)
MOBILENET_V2_STRIDES = {1: 1, 2: 2, 3: 1, 4: 2, 5: 1, 6: 1, 7: 2, 8: 1, 9: 1, 10: 1, 11: 1, 12: 1, 13: 1, 14: 2, 15: 1, 16: 1, 17: 1}


def build_plan(sd: Dict[str, torch.Tensor], *, backbone: str = "resnet34", neck: str = "FPN",
               head_names: Sequence[str] = ("heatmap", "box_2d"), head_depth: int = 3,
               prefix: str = "") -> Plan:
    """Lower a state dict with the G2 key layout (``backbone.*``, ``neck.*``, ``heads.<h>.block_<i>.*``,
    ``heads.<h>.out_conv.*``; reference models/meta.py:26-28,36-38,92-95) into a Plan."""
    sd = {k[len(prefix):]: v.detach().cpu() for k, v in sd.items() if k.startswith(prefix)}
    p = Plan()
    p.add_buffer("image", 3, 1, fp32_nchw=True)
    lw = _Lowering(p, sd)
    feats = _lower_mobilenet_v2(lw) if backbone == "mobilenet_v2" else _lower_resnet(lw, backbone)
    out_ch = backbone_out_channels(backbone)

    # ---- neck -----------------------------------------------------------------------------------
    if neck == "FPN":
        d = sd["neck.lateral.0.weight"].shape[0]
        w, b = fold_bn(sd["neck.lateral.3.weight"], None, sd["neck.lateral.3.bias"])
        x = p.add_buffer("neck.p5", pad64(d), 32)
        lw.conv("neck.lateral.3", feats[3], x, w, b, ksize=1, relu=0)
        for i in (2, 1, 0):
            s = 4 * (2 ** i)
            w, b = fold_bn(sd[f"neck.lateral.{i}.weight"], None, sd[f"neck.lateral.{i}.bias"])
            fused = p.add_buffer(f"neck.sum{i}", pad64(d), s)
            lw.conv(f"neck.lateral.{i}", feats[i], fused, w, b, ksize=1, relu=0, residual=x, residual_up=2)
            w, b = fold_bn(sd[f"neck.output.{i}.conv.weight"], _bn(sd, f"neck.output.{i}.bn"))
            x = p.add_buffer(f"neck.out{i}", pad64(d), s)
            lw.conv(f"neck.output.{i}", fused, x, w, b, ksize=3, relu=1)
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
                y = p.add_buffer(f"neck.conv{i}", pad64(c), s)
                lw.conv(f"neck.blocks.{i}", x, y, w, b, ksize=3, relu=1)
                up = p.add_buffer(f"neck.up{i}", pad64(c), s // 2)
                for op in lower_conv_transpose(f"neck.up.{i}", y, up, sd[f"neck.up.{i}.0.weight"], _bn(sd, f"neck.up.{i}.1")):
                    p.ops.append(op)
            else:
                up = p.add_buffer(f"neck.up{i}", pad64(c), s // 2)
                lw.conv(f"neck.blocks.{i}", x, up, w, b, ksize=3, relu=1, dst_up=2, dst_phase=-1)
            x, cin_n, s = up, c, s // 2
            i += 1
        if i == 0:
            raise ValueError("simple neck: no neck.blocks.* parameters in the state dict")
        neck_out, neck_c, neck_stride = x, cin_n, s
    elif neck in ("ida", "IDANeck"):
        # iterative deep aggregation (reference docs/implementation.md:43) from Fuse nodes (models/layers.py:138-177):
        # every level fuses consecutive maps pairwise until one stride-4 map is left
        cur = [(f, c) for f, c in zip(feats, out_ch)]
        lvl = 0
        while len(cur) > 1:
            nxt = []
            for i in range(len(cur) - 1):
                s = 4 * (2 ** i)
                o = lw.fuse(f"neck.levels.{lvl}.{i}", f"neck.levels.{lvl}.{i}", [cur[i], cur[i + 1]], cur[i][1], s, "up")
                nxt.append((o, cur[i][1]))
            cur, lvl = nxt, lvl + 1
        neck_out, neck_c, neck_stride = cur[0][0], cur[0][1], 4
    elif neck in ("bifpn", "BiFPNNeck"):
        # BiFPN (reference docs/implementation.md:42) from Fuse nodes: 1x1 projections to D channels, then per layer a
        # top-down pass (Fuse "up") and a bottom-up pass (Fuse "down" over p_i, td_i, out_{i-1})
        d = sd["neck.project.0.weight"].shape[0]
        n = len(feats)
        pl: List[str] = []
        for i in range(n):
            w, b = fold_bn(sd[f"neck.project.{i}.weight"], None, sd[f"neck.project.{i}.bias"])
            t = p.add_buffer(f"neck.p{i}", pad64(d), 4 * (2 ** i))
            lw.conv(f"neck.project.{i}", feats[i], t, w, b, ksize=1, relu=0)
            pl.append(t)
        layer = 0
        while f"neck.top_down.{layer}.0.output_conv.0.weight" in sd:
            td: List[Optional[str]] = [None] * n
            td[n - 1] = pl[n - 1]
            for i in range(n - 2, -1, -1):
                td[i] = lw.fuse(f"neck.top_down.{layer}.{i}", f"neck.top_down.{layer}.{i}", [(pl[i], d), (td[i + 1], d)], d, 4 * (2 ** i), "up")
            out: List[Optional[str]] = [td[0]] + [None] * (n - 1)
            for i in range(1, n):
                ins = [(pl[i], d), (td[i], d), (out[i - 1], d)] if i < n - 1 else [(pl[i], d), (out[i - 1], d)]
                out[i] = lw.fuse(f"neck.bottom_up.{layer}.{i - 1}", f"neck.bottom_up.{layer}.{i - 1}", ins, d, 4 * (2 ** i), "down")
            pl = out
            layer += 1
        if layer == 0:
            raise ValueError("bifpn neck: no neck.top_down.* parameters in the state dict")
        neck_out, neck_c, neck_stride = pl[0], d, 4
    else:
        raise ValueError(f"neck {neck!r} is not lowered by the sm_100a engine (FPN, simple, ida, bifpn; SURVEY 8a F2, 8f rank 4)")
    p.model_stride = neck_stride

    # ---- heads: first tower layers fused across heads ------------------------------------------
    width = sd[f"heads.{head_names[0]}.out_conv.weight"].shape[1]
    if width % 64:
        raise ValueError("head width must be a multiple of 64")
    nh = len(head_names)
    if head_depth >= 1:
        ws, bs = [], []
        for h in head_names:
            w, b = fold_bn(sd[f"heads.{h}.block_1.conv.weight"], _bn(sd, f"heads.{h}.block_1.bn"))
            ws.append(w); bs.append(b)
        t = p.add_buffer("heads.t1", width * nh, neck_stride)
        lw.conv("heads.block_1", neck_out, t, torch.cat(ws), torch.cat(bs), ksize=3, relu=1)
        for li in range(2, head_depth + 1):
            t2 = p.add_buffer(f"heads.t{li}", width * nh, neck_stride)
            for hi, h in enumerate(head_names):
                w, b = fold_bn(sd[f"heads.{h}.block_{li}.conv.weight"], _bn(sd, f"heads.{h}.block_{li}.bn"))
                lw.conv(f"heads.{h}.block_{li}", t, t2, w, b, ksize=3, relu=1, src_c_off=hi * width, dst_c_off=hi * width)
            t = t2
        tower = t
    else:
        tower = neck_out
    for hi, h in enumerate(head_names):
        w = sd[f"heads.{h}.out_conv.weight"].float().contiguous()
        b = sd[f"heads.{h}.out_conv.bias"].float().contiguous()
        o = p.add_buffer(f"out.{h}", w.shape[0], neck_stride, fp32_nchw=True)
        lw.conv(f"heads.{h}.out_conv", tower, o, w, b, ksize=1, relu=0, src_c_off=(hi * width if head_depth >= 1 else 0))
        p.outputs[h] = o
    return p
