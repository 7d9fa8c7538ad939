"""Forward oracle: CPU fp32 torch restatement of the reference's model graph.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
path (``centernet-lightning_b200/``); only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may use it.

What it restates (paths relative to /root/reference):

* ``centernet_lightning/models/meta.py:41-47``  GenericModel.forward
  (backbone.forward_features -> neck -> {name: head(x)})
* ``centernet_lightning/models/meta.py:21-30``  GenericHead
  (depth x ConvBnAct named block_1..block_d, then out_conv = Conv2d(width, out, 1)
  with bias filled with init_bias)
* ``centernet_lightning/models/meta.py:87-96``  backbone/neck/head construction
  and ``stride = backbone.stride // neck.stride``
* ``centernet_lightning/models/centernet.py:102-105``  the two CenterNet heads
  (heatmap bias = log(p/(1-p)), box_2d bias = box_init_bias)

The backbone / neck / ConvBnAct arithmetic lives in the un-vendored, un-pinned
third-party package ``vision_toolbox`` (requirements.txt:11, setup.cfg:21:
``git+https://github.com/gau-nernst/vision-toolbox.git`` with no commit).  Its
behaviour is reconstructed from the reference's call sites and from the parameter
counts the reference publishes (docs/experiments.md:24-27: ResNet-34 21.3M,
FPN(256) 2.0M, heads(256x3) 3.6M) - SURVEY.md Appendix B lists every decision.
PARITY STATUS: "parity unpinned" for vision_toolbox internals (no reference test
or golden vector pins conv arithmetic); the GenericModel/GenericHead wiring IS
pinned against the reference's own classes by tests/golden/gen_golden.py.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
from torch import nn
import torch.nn.functional as F


class ConvBnAct(nn.Sequential):
    """3x3 conv (no bias) -> BatchNorm2d -> ReLU.  SURVEY Appendix B4.

    Stand-in for vision_toolbox.components.ConvBnAct (call sites
    models/meta.py:22,26); the 3x3 kernel is pinned by the 3.6M head parameter
    count (docs/experiments.md:27)."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, 3, padding=1, bias=False)
        self.bn = nn.BatchNorm2d(out_channels)
        self.act = nn.ReLU(inplace=True)


class BasicBlock(nn.Module):
    """torchvision-compatible ResNet BasicBlock (same parameter names)."""

    def __init__(self, cin: int, cout: int, stride: int):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        y = F.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        return F.relu(y + idt)


class Bottleneck(nn.Module):
    """torchvision-compatible ResNet-50 Bottleneck (v1.5: the stride sits on the 3x3 conv), same parameter names."""
    expansion = 4

    def __init__(self, cin: int, width: int, stride: int):
        super().__init__()
        cout = width * self.expansion
        self.conv1 = nn.Conv2d(cin, width, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(width)
        self.conv2 = nn.Conv2d(width, width, 3, stride, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(width)
        self.conv3 = nn.Conv2d(width, cout, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        y = F.relu(self.bn1(self.conv1(x)))
        y = F.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        return F.relu(y + idt)


_RESNET_DEPTHS = {"resnet18": (2, 2, 2, 2), "resnet34": (3, 4, 6, 3), "resnet50": (3, 4, 6, 3)}


class ResNetTrunk(nn.Module):
    """ResNet-18/34/50 trunk without avgpool/fc (resnet50 is one of the backbones the reference's
    tests name, tests/test_models.py:37-39).

    Interface mirrors what models/meta.py:42,87-88,96 needs from a
    vision_toolbox backbone: forward_features, get_out_channels, stride.
    Parameter names equal torchvision.models.resnet34 so torchvision / reference
    checkpoints load unchanged."""

    stride = 32

    def __init__(self, name: str = "resnet34"):
        super().__init__()
        depths = _RESNET_DEPTHS[name]
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        widths = (64, 128, 256, 512)
        bottleneck = name == "resnet50"
        self.out_channels = [w * (4 if bottleneck else 1) for w in widths]
        cin = 64
        for i, (w, d) in enumerate(zip(widths, depths)):
            blocks = []
            for j in range(d):
                stride = 2 if (j == 0 and i > 0) else 1
                blocks.append(Bottleneck(cin, w, stride) if bottleneck else BasicBlock(cin, w, stride))
                cin = w * (4 if bottleneck else 1)
            setattr(self, f"layer{i + 1}", nn.Sequential(*blocks))

    def get_out_channels(self) -> List[int]:
        return list(self.out_channels)

    def forward_features(self, x: torch.Tensor) -> List[torch.Tensor]:
        x = self.maxpool(F.relu(self.bn1(self.conv1(x))))
        outs = []
        for i in range(4):
            x = getattr(self, f"layer{i + 1}")(x)
            outs.append(x)
        return outs


class MobileNetV2Trunk(nn.Module):
    """torchvision mobilenet_v2 ``features[0:18]`` as a 4-level backbone (one of the backbones the reference's tests
    name, tests/test_models.py:37).  vision_toolbox's own wrapper is absent from the snapshot; like the ResNet trunks
     it returns the stride 4 / 8 / 16 / 32 maps: outputs of features[3], [6], [13], [17]
    (24 / 32 / 96 / 320 channels); the final 1x1 conv to 1280 channels is the classifier's and is dropped.
    Parameter names equal torchvision's (``features.<i>...``)."""

    stride = 32
    out_indices = (3, 6, 13, 17)

    def __init__(self):
        super().__init__()
        import torchvision
        self.features = torchvision.models.mobilenet_v2(weights=None).features[:18]
        self.out_channels = [24, 32, 96, 320]

    def get_out_channels(self) -> List[int]:
        return list(self.out_channels)

    def forward_features(self, x: torch.Tensor) -> List[torch.Tensor]:
        outs = []
        for i, f in enumerate(self.features):
            x = f(x)
            if i in self.out_indices:
                outs.append(x)
        return outs


def make_conv(in_channels: int, out_channels: int, conv_type: str = "normal") -> nn.Sequential:
    """models/layers.py:40-79 restated (kernel 3, the default): "normal" = conv3x3(no bias)-BN-ReLU; "separable" =
    depthwise conv3x3(no bias)-BN-ReLU6 then pointwise conv1x1(no bias)-BN-ReLU6 (depth_multiplier 1).  Same Sequential
    indices as the reference so that state dicts interchange."""
    assert conv_type in ("separable", "normal")
    if conv_type == "separable":
        return nn.Sequential(
            nn.Conv2d(in_channels, in_channels, 3, padding=1, groups=in_channels, bias=False), nn.BatchNorm2d(in_channels),
            nn.ReLU6(inplace=True),
            nn.Conv2d(in_channels, out_channels, 1, bias=False), nn.BatchNorm2d(out_channels), nn.ReLU6(inplace=True))
    return nn.Sequential(nn.Conv2d(in_channels, out_channels, 3, padding=1, bias=False), nn.BatchNorm2d(out_channels),
                         nn.ReLU(inplace=True))


class Fuse(nn.Module):
    """models/layers.py:138-177 restated: the fusion node of BiFPNNeck / IDANeck.  Every input is projected to ``out``
    channels by a 1x1 conv WITH bias when its channel count differs (:149-151); the LAST input is resized - nearest x2
    (make_upsample's default, :81-99) or MaxPool2d(2, 2) (make_downsample's default, :119-136; Fuse passes its
    ``downsample`` argument under a keyword make_downsample does not have, so the default always applies) - then
    ``out = output_conv(sum_i w_i x_i / (sum_i w_i + eps))`` with w = relu(weights) (weighted_fusion) or the plain sum."""

    def __init__(self, in_channels: List[int], out: int, resize: str, conv_type: str = "normal", weighted_fusion: bool = False):
        super().__init__()
        assert resize in ("up", "down")
        self.project = nn.ModuleList([nn.Conv2d(c, out, 1) if c != out else None for c in in_channels])
        self.weights = nn.Parameter(torch.ones(len(in_channels))) if weighted_fusion else None
        self.resize = nn.Upsample(scale_factor=2, mode="nearest") if resize == "up" else nn.MaxPool2d(2, 2)
        self.output_conv = make_conv(out, out, conv_type=conv_type)

    def forward(self, *features, eps: float = 1e-6):
        out = [p(x) if p is not None else x for p, x in zip(self.project, features)]
        out[-1] = self.resize(out[-1])
        if self.weights is not None:
            w = F.relu(self.weights)
            out = torch.stack([out[i] * w[i] for i in range(len(out))], dim=-1)
            out = torch.sum(out, dim=-1) / (torch.sum(w) + eps)
        else:
            out = torch.sum(torch.stack(out, dim=-1), dim=-1)
        return self.output_conv(out)


class IDANeck(nn.Module):
    """Iterative deep aggregation over the backbone's feature maps ("iteratively fuse consecutive feature maps from
    backbone until there is only 1 feature map left", reference docs/implementation.md:43; the class itself is not in the
    snapshot - models/__init__.py:2 - so the wiring is reconstructed from that sentence, from the Fuse node
    (models/layers.py:138-177) and from DLA's IDA): level l+1 = [Fuse([c_i, c_{i+1}], c_i, "up")(f_i, f_{i+1}) for every
    consecutive pair]; after len(in_channels)-1 levels one stride-4 map with in_channels[0] channels is left."""

    def __init__(self, in_channels: List[int], conv_type: str = "normal", weighted_fusion: bool = False):
        super().__init__()
        self.stride = 2 ** (len(in_channels) - 1)
        self.out_channels = in_channels[0]
        self.levels = nn.ModuleList()
        chans = list(in_channels)
        while len(chans) > 1:
            self.levels.append(nn.ModuleList([Fuse([chans[i], chans[i + 1]], chans[i], "up", conv_type=conv_type, weighted_fusion=weighted_fusion)
                                              for i in range(len(chans) - 1)]))
            chans = chans[:-1]

    def get_out_channels(self) -> int:
        return self.out_channels

    def forward(self, feats: List[torch.Tensor]) -> torch.Tensor:
        feats = list(feats)
        for nodes in self.levels:
            feats = [node(feats[i], feats[i + 1]) for i, node in enumerate(nodes)]
        return feats[0]


class BiFPNNeck(nn.Module):
    """BiFPN (EfficientDet, reference docs/implementation.md:42) from the reference's Fuse nodes; class absent from the
    snapshot, reconstructed: every level is first projected to ``out_channels`` by a 1x1 conv (with bias); each of the
    ``num_layers`` layers runs a top-down pass td_i = Fuse([D, D], D, "up")(p_i, td_{i+1}) and a bottom-up pass
    out_i = Fuse([D, D, D], D, "down")(p_i, td_i, out_{i-1}) (out_0 = td_0; the top level has no td input:
    Fuse([D, D], D, "down")(p_top, out_{top-1})); the finest level of the last layer is returned (stride 4)."""

    def __init__(self, in_channels: List[int], out_channels: int = 64, num_layers: int = 2, conv_type: str = "normal",
                 weighted_fusion: bool = True):
        super().__init__()
        n = len(in_channels)
        self.stride = 2 ** (n - 1)
        self.out_channels = out_channels
        d = out_channels
        self.project = nn.ModuleList([nn.Conv2d(c, d, 1) for c in in_channels])
        self.top_down = nn.ModuleList([nn.ModuleList([Fuse([d, d], d, "up", conv_type=conv_type, weighted_fusion=weighted_fusion) for _ in range(n - 1)])
                                       for _ in range(num_layers)])
        self.bottom_up = nn.ModuleList([nn.ModuleList([Fuse([d, d, d] if i < n - 2 else [d, d], d, "down", conv_type=conv_type, weighted_fusion=weighted_fusion)
                                                       for i in range(n - 1)]) for _ in range(num_layers)])

    def get_out_channels(self) -> int:
        return self.out_channels

    def forward(self, feats: List[torch.Tensor]) -> torch.Tensor:
        p = [proj(f) for proj, f in zip(self.project, feats)]
        n = len(p)
        for td_nodes, bu_nodes in zip(self.top_down, self.bottom_up):
            td = [None] * n
            td[n - 1] = p[n - 1]
            for i in range(n - 2, -1, -1):
                td[i] = td_nodes[i](p[i], td[i + 1])
            out = [td[0]] + [None] * (n - 1)
            for i in range(1, n):
                node = bu_nodes[i - 1]
                out[i] = node(p[i], td[i], out[i - 1]) if i < n - 1 else node(p[i], out[i - 1])
            p = out
        return p[0]


class FPN(nn.Module):
    """Top-down FPN that returns only the finest level.

    lateral = Conv2d(c_i, D, 1, bias=True); x = lateral(C5); for C4, C3, C2:
    x = ConvBnAct3x3(lateral(C_i) + nearest_up2x(x)).  2.018M params at D=256
    (docs/experiments.md:27)."""

    def __init__(self, in_channels: List[int], out_channels: int = 256, fuse_fn: str = "sum"):
        super().__init__()
        if fuse_fn != "sum":
            raise ValueError("only fuse_fn='sum' is on the hot path (configs/centernet.yaml:9)")
        self.out_channels = out_channels
        self.stride = 2 ** (len(in_channels) - 1)
        self.lateral = nn.ModuleList([nn.Conv2d(c, out_channels, 1) for c in in_channels])
        self.output = nn.ModuleList([ConvBnAct(out_channels, out_channels) for _ in in_channels[:-1]])

    def get_out_channels(self) -> int:
        return self.out_channels

    def forward(self, feats: List[torch.Tensor]) -> torch.Tensor:
        x = self.lateral[-1](feats[-1])
        for i in range(len(feats) - 2, -1, -1):
            x = self.lateral[i](feats[i]) + F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = self.output[i](x)
        return x


class SimpleNeck(nn.Module):
    """BASELINE config 1 neck: 3 x [ConvBnAct3x3 -> x2 upsample]
    on C5 only, channels 512->256->128->64 (configs/base_resnet34.yaml:7-11,
    tests/test_necks.py:23-38).  ``upsample_type``: "nearest" (the shipped config) or
    "conv_transpose" = ConvTranspose2d(c, c, k, stride=2, padding, output_padding,
    bias=False) -> BN -> ReLU exactly as models/layers.py:86-96 builds it
    (output_padding = k % 2, padding = (k + output_padding) // 2 - 1)."""

    def __init__(self, in_channels: List[int], upsample_channels=(256, 128, 64), upsample_type: str = "nearest",
                 deconv_kernel: int = 3):
        super().__init__()
        assert upsample_type in ("nearest", "conv_transpose")
        self.stride = 2 ** len(upsample_channels)
        chans = [in_channels[-1], *upsample_channels]
        self.blocks = nn.ModuleList([ConvBnAct(chans[i], chans[i + 1]) for i in range(len(upsample_channels))])
        self.out_channels = chans[-1]
        self.upsample_type = upsample_type
        if upsample_type == "conv_transpose":
            output_padding = deconv_kernel % 2
            padding = (deconv_kernel + output_padding) // 2 - 1
            self.up = nn.ModuleList([
                nn.Sequential(nn.ConvTranspose2d(c, c, deconv_kernel, stride=2, padding=padding,
                                                 output_padding=output_padding, bias=False),
                              nn.BatchNorm2d(c), nn.ReLU(inplace=True))
                for c in upsample_channels])

    def get_out_channels(self) -> int:
        return self.out_channels

    def forward(self, feats: List[torch.Tensor]) -> torch.Tensor:
        x = feats[-1]
        for i, blk in enumerate(self.blocks):
            x = blk(x)
            x = self.up[i](x) if self.upsample_type == "conv_transpose" else F.interpolate(x, scale_factor=2.0, mode="nearest")
        return x


class Head(nn.Sequential):
    """models/meta.py:21-30 restated (block_1..block_depth + out_conv with bias)."""

    def __init__(self, in_channels: int, out_channels: int, width: int = 256, depth: int = 3,
                 init_bias: Optional[float] = None):
        super().__init__()
        for i in range(depth):
            self.add_module(f"block_{i + 1}", ConvBnAct(in_channels if i == 0 else width, width))
        self.out_conv = nn.Conv2d(width, out_channels, 1)
        if init_bias is not None:
            self.out_conv.bias.data.fill_(init_bias)


class SpecModel(nn.Module):
    """models/meta.py:33-47 restated.  forward() returns {head_name: logits}."""

    def __init__(self, backbone: nn.Module, neck: nn.Module, heads: nn.Module):
        super().__init__()
        self.backbone = backbone
        self.neck = neck
        self.heads = heads

    def forward(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        feats = self.backbone.forward_features(x)
        y = self.neck(feats)
        return {name: head(y) for name, head in self.heads.named_children()}


def build_spec_model(num_classes: int = 80, backbone: str = "resnet34", neck: str = "FPN",
                     neck_config: Optional[dict] = None, head_config: Optional[dict] = None,
                     heatmap_prior: float = 0.01, box_init_bias: Optional[float] = None,
                     reid_dim: int = 0) -> SpecModel:
    """Construct the graph the way models/meta.py:87-96 + models/centernet.py:102-105 do.

    reid_dim > 0 adds the tracking head (GenericHead(256, 64, **head_config))."""
    neck_config = dict(neck_config or {})
    head_config = dict(head_config or {})
    bb = MobileNetV2Trunk() if backbone == "mobilenet_v2" else ResNetTrunk(backbone)
    if neck == "FPN":
        neck_config.setdefault("out_channels", 256)
        nk = FPN(bb.get_out_channels(), **neck_config)
    elif neck in ("simple", "SimpleNeck"):
        nk = SimpleNeck(bb.get_out_channels(), **neck_config)
    elif neck in ("ida", "IDANeck"):
        nk = IDANeck(bb.get_out_channels(), **neck_config)
    elif neck in ("bifpn", "BiFPNNeck"):
        nk = BiFPNNeck(bb.get_out_channels(), **neck_config)
    else:
        raise ValueError(f"unknown neck {neck!r}")
    heads = nn.Module()
    c = nk.get_out_channels()
    heads.add_module("heatmap", Head(c, num_classes, init_bias=math.log(heatmap_prior / (1 - heatmap_prior)), **head_config))
    heads.add_module("box_2d", Head(c, 4, init_bias=box_init_bias, **head_config))
    if reid_dim:
        heads.add_module("reid", Head(c, reid_dim, init_bias=None, **head_config))
    m = SpecModel(bb, nk, heads)
    m.stride = bb.stride // nk.stride          # models/meta.py:96
    return m.eval()


@torch.no_grad()
def synth_init(model: SpecModel, seed: int = 0, calib_size: int = 128, calib_batch: int = 4,
               out_gain: float = 3.0) -> SpecModel:
    """Deterministic synthetic weights that behave like a trained network.

    There is no network access for checkpoints, so benchmarks and parity tests
    use random weights.  Plain default-init + identity BN statistics lets the
    activations collapse towards the biases after ~40 layers, which hides
    numerical error (SURVEY section 7 'hard parts').  Instead: He-normal conv
    weights, gamma ~ U(0.5,1.5), beta ~ N(0,0.1), and BN running statistics
    *calibrated* on a seeded random batch so every layer sees unit-scale inputs
    as it would after training.  Final 1x1 weights are scaled by ``out_gain`` so
    the heatmap has distinct peaks."""
    g = torch.Generator().manual_seed(seed)
    for mod in model.modules():
        if isinstance(mod, nn.Conv2d):
            fan_in = mod.in_channels // mod.groups * mod.kernel_size[0] * mod.kernel_size[1]
            mod.weight.copy_(torch.randn(mod.weight.shape, generator=g) * math.sqrt(2.0 / fan_in))
            if mod.bias is not None and mod is not model.heads.heatmap.out_conv:
                mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.1)    # heatmap keeps its prior bias
        elif isinstance(mod, nn.ConvTranspose2d):
            # every output pixel of a stride-2 transposed conv sees about k*k/4 taps per input channel
            fan_in = mod.in_channels * mod.kernel_size[0] * mod.kernel_size[1] / 4.0
            mod.weight.copy_(torch.randn(mod.weight.shape, generator=g) * math.sqrt(2.0 / fan_in))
        elif isinstance(mod, nn.BatchNorm2d):
            mod.weight.copy_(torch.rand(mod.weight.shape, generator=g) + 0.5)
            mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.1)
    for blk in model.modules():
        if isinstance(blk, BasicBlock):        # damp residual branches like a trained ResNet
            blk.bn2.weight.mul_(0.5)
        elif isinstance(blk, Bottleneck):
            blk.bn3.weight.mul_(0.5)
        elif type(blk).__name__ == "InvertedResidual" and blk.use_res_connect:
            blk.conv[-1].weight.mul_(0.5)
        elif type(blk).__name__ == "Fuse" and blk.weights is not None:      # fusion weights: positive, not all equal
            blk.weights.copy_(torch.rand(blk.weights.shape, generator=g) + 0.5)
    for head in model.heads.children():
        head.out_conv.weight.mul_(out_gain / math.sqrt(2.0))
    # calibrate BN running stats with one train-mode pass (momentum=1 -> stats of this batch)
    bns = [m for m in model.modules() if isinstance(m, nn.BatchNorm2d)]
    for bn in bns:
        bn.momentum = 1.0
    model.train()
    x = torch.rand((calib_batch, 3, calib_size, calib_size), generator=g)
    model(x)
    model.eval()
    for bn in bns:
        bn.momentum = 0.1
    return model


def count_params(mod: nn.Module) -> int:
    return sum(p.numel() for p in mod.parameters())
