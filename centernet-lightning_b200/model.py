"""Drop-in model object: the reference's Python surface over the sm_100a engine.

G2 surface (authoritative semantics; reference centernet_lightning/models/centernet.py:68-121, 229-304 and
models/meta.py:41-47, 96):
    CenterNet(num_classes, backbone, pretrained_backbone, neck, neck_config, head_config, box_log, box_multiplier,
              heatmap_prior, nms_kernel, num_detections, ...)
    .model(images) -> {"heatmap": logits, "box_2d": ltrb[, "reid": emb]}      (GenericModel.forward)
    .decode_detections(heatmap_prob, box_offsets, normalize_boxes=False) -> {"boxes","scores","labels"}
    .get_topk_from_heatmap(heatmap, pseudo_nms=True) -> (scores, indices, labels)
    CenterNet.gather_and_decode_boxes(box_offsets, indices, ...)               (staticmethod)
    .stride, .num_classes, .hparams
G1 aliases (README.md:27-101, tests/test_models.py:61-99, models/fairmot.py:138-151, utils/image_annotate.py:220-223):
    forward(images) -> (heatmap_sigmoid, box_2d[, reid]), get_encoded_outputs / get_output_dict, gather_detection2d,
    gather_tracking2d, inference_detection(img_dir), output_stride.
Plus the fused entry point this package adds: detect(images) = forward + sigmoid + NMS + top-k + gather in one
CUDA-graph replay (what reference validation_step :202-205 does in ~70 eager kernels).

State-dict keys are the G2 ones: model.backbone.*, model.neck.*, model.heads.<head>.block_<i>.{conv,bn}.*,
model.heads.<head>.out_conv.{weight,bias} (reference models/meta.py:26-28, 36-38, 92-95).  The nn.Conv2d /
nn.BatchNorm2d objects below only HOLD parameters (default torch init, load_state_dict); they are never called.
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict, namedtuple
from types import SimpleNamespace
from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch
from torch import nn

from . import decode as _decode
from . import _lib
from .engine import Engine, PRECISION_FAST, PRECISION_SPLIT
from .plan import MOBILENET_V2_STRIDES, RESNET_DEPTHS, RESNET_WIDTHS, backbone_out_channels, build_plan, resnet_out_channels

_PRECISIONS = {"split": PRECISION_SPLIT, "fp32": PRECISION_SPLIT, "split_fused": 2, "fast": PRECISION_FAST, "fp16": PRECISION_FAST}


# ----------------------------------------------------------------------------------------------------------------
# parameter containers (same module tree / key names as the reference graph)
# ----------------------------------------------------------------------------------------------------------------
class _ConvBn(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 3, padding=1, bias=False)
        self.bn = nn.BatchNorm2d(cout)


class _Block(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))


class _Bottleneck(nn.Module):
    def __init__(self, cin, width, stride):
        super().__init__()
        cout = 4 * width
        self.conv1 = nn.Conv2d(cin, width, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(width)
        self.conv2 = nn.Conv2d(width, width, 3, stride, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(width)
        self.conv3 = nn.Conv2d(width, cout, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(cout)
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))


class _InvertedResidual(nn.Module):
    """Parameter layout of torchvision's mobilenet_v2 InvertedResidual (``conv.<i>...`` keys)."""

    def __init__(self, cin, cout, stride, expand):
        super().__init__()
        hid = cin * expand
        layers = []
        if expand != 1:
            layers.append(nn.Sequential(nn.Conv2d(cin, hid, 1, bias=False), nn.BatchNorm2d(hid)))
        layers += [nn.Sequential(nn.Conv2d(hid, hid, 3, stride, 1, groups=hid, bias=False), nn.BatchNorm2d(hid)),
                   nn.Conv2d(hid, cout, 1, bias=False), nn.BatchNorm2d(cout)]
        self.conv = nn.Sequential(*layers)
        self.use_res_connect = stride == 1 and cin == cout


class _MobileNetV2(nn.Module):
    """torchvision mobilenet_v2 ``features[0:18]`` parameter tree (one of the backbones the reference's tests name,
    tests/test_models.py:37); features at strides 4 / 8 / 16 / 32 = outputs of features[3], [6], [13], [17]."""
    stride = 32
    name = "mobilenet_v2"

    def __init__(self):
        super().__init__()
        feats = [nn.Sequential(nn.Conv2d(3, 32, 3, 2, 1, bias=False), nn.BatchNorm2d(32))]
        cin = 32
        for t, c, n, s_ in ((1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1)):
            for i in range(n):
                feats.append(_InvertedResidual(cin, c, s_ if i == 0 else 1, t))
                cin = c
        self.features = nn.Sequential(*feats)
        assert all(MOBILENET_V2_STRIDES[i] == f.conv[-3][0].stride[0] for i, f in enumerate(self.features) if i > 0)

    def get_out_channels(self):
        return list(backbone_out_channels("mobilenet_v2"))


class _Backbone(nn.Module):
    stride = 32

    def __init__(self, name):
        super().__init__()
        if name not in RESNET_DEPTHS:
            raise ValueError(f"backbone {name!r}: the sm_100a engine implements {sorted(RESNET_DEPTHS) + ['mobilenet_v2']}")
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        cin = 64
        self.name = name
        bottleneck = name == "resnet50"
        for i, (w, d) in enumerate(zip(RESNET_WIDTHS, RESNET_DEPTHS[name])):
            blocks = []
            for j in range(d):
                stride = 2 if (j == 0 and i > 0) else 1
                blocks.append(_Bottleneck(cin, w, stride) if bottleneck else _Block(cin, w, stride))
                cin = 4 * w if bottleneck else w
            setattr(self, f"layer{i + 1}", nn.Sequential(*blocks))

    def get_out_channels(self):
        return list(resnet_out_channels(self.name))


class _FPN(nn.Module):
    def __init__(self, in_channels, out_channels=256, fuse_fn="sum"):
        super().__init__()
        if fuse_fn != "sum":
            raise ValueError("the sm_100a engine implements fuse_fn='sum' (reference configs/centernet.yaml:9)")
        self.stride = 2 ** (len(in_channels) - 1)
        self.out_channels = out_channels
        self.lateral = nn.ModuleList([nn.Conv2d(c, out_channels, 1) for c in in_channels])
        self.output = nn.ModuleList([_ConvBn(out_channels, out_channels) for _ in in_channels[:-1]])

    def get_out_channels(self):
        return self.out_channels


class _SimpleNeck(nn.Module):
    """G1 "simple" neck (reference configs/base_resnet34.yaml:7-11, models/layers.py:71-99):
    on C5 only, n x [conv3x3-BN-ReLU -> x2 upsample], upsample = nearest or ConvTranspose2d-BN-ReLU."""

    def __init__(self, in_channels, upsample_channels=(256, 128, 64), upsample_type="nearest", deconv_kernel=3,
                 conv_type="normal"):
        super().__init__()
        if upsample_type not in ("nearest", "conv_transpose"):
            raise ValueError(f"upsample_type {upsample_type!r}: the sm_100a engine lowers 'nearest' and 'conv_transpose'")
        if conv_type != "normal":
            raise ValueError(f"conv_type {conv_type!r}: the sm_100a engine lowers 'normal' convolutions")
        if deconv_kernel not in (3, 4):
            raise ValueError("deconv_kernel must be 3 or 4")
        if any(c % 64 for c in upsample_channels):
            raise ValueError("upsample_channels must be multiples of 64")
        self.stride = 2 ** len(upsample_channels)
        chans = [in_channels[-1], *upsample_channels]
        self.blocks = nn.ModuleList([_ConvBn(chans[i], chans[i + 1]) for i in range(len(upsample_channels))])
        self.out_channels = chans[-1]
        if upsample_type == "conv_transpose":
            op = deconv_kernel % 2
            self.up = nn.ModuleList([
                nn.Sequential(nn.ConvTranspose2d(c, c, deconv_kernel, stride=2, padding=(deconv_kernel + op) // 2 - 1,
                                                 output_padding=op, bias=False), nn.BatchNorm2d(c))
                for c in upsample_channels])

    def get_out_channels(self):
        return self.out_channels


def _make_conv(cin, cout, conv_type):
    """Parameter layout of the reference's make_conv (models/layers.py:40-79): Sequential indices 0,1 (normal) or 0,1,3,4
    (separable: depthwise 3x3 + BN, pointwise 1x1 + BN; index 2 / 5 are the parameter-free activations)."""
    if conv_type == "separable":
        return nn.Sequential(nn.Conv2d(cin, cin, 3, padding=1, groups=cin, bias=False), nn.BatchNorm2d(cin), nn.Identity(),
                             nn.Conv2d(cin, cout, 1, bias=False), nn.BatchNorm2d(cout))
    if conv_type != "normal":
        raise ValueError(f"conv_type {conv_type!r}: 'normal' and 'separable' are lowered (deformable convs are not)")
    return nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout))


class _Fuse(nn.Module):
    """Parameter layout of the reference's Fuse node (models/layers.py:138-177): project.<i> (1x1 conv with bias where the
    channel count differs), weights (weighted fusion), output_conv (make_conv)."""

    def __init__(self, in_channels, out, conv_type="normal", weighted_fusion=False):
        super().__init__()
        self.project = nn.ModuleList([nn.Conv2d(c, out, 1) if c != out else None for c in in_channels])
        self.weights = nn.Parameter(torch.ones(len(in_channels))) if weighted_fusion else None
        self.output_conv = _make_conv(out, out, conv_type)


class _IDANeck(nn.Module):
    """IDANeck (reference docs/implementation.md:43; class absent from the snapshot, reconstructed from its Fuse node):
    consecutive maps are fused pairwise, level after level, until one stride-4 map with in_channels[0] channels is left."""

    def __init__(self, in_channels, conv_type="normal", weighted_fusion=False):
        super().__init__()
        self.stride = 2 ** (len(in_channels) - 1)
        self.out_channels = in_channels[0]
        self.levels = nn.ModuleList()
        chans = list(in_channels)
        while len(chans) > 1:
            self.levels.append(nn.ModuleList([_Fuse([chans[i], chans[i + 1]], chans[i], conv_type, weighted_fusion)
                                              for i in range(len(chans) - 1)]))
            chans = chans[:-1]

    def get_out_channels(self):
        return self.out_channels


class _BiFPNNeck(nn.Module):
    """BiFPNNeck (reference docs/implementation.md:42; reconstructed from the Fuse node): 1x1 projections to ``out_channels``,
    then ``num_layers`` x (top-down Fuse "up" pass, bottom-up Fuse "down" pass); the stride-4 map is returned."""

    def __init__(self, in_channels, out_channels=64, num_layers=2, conv_type="normal", weighted_fusion=True):
        super().__init__()
        n, d = len(in_channels), out_channels
        self.stride = 2 ** (n - 1)
        self.out_channels = d
        self.project = nn.ModuleList([nn.Conv2d(c, d, 1) for c in in_channels])
        self.top_down = nn.ModuleList([nn.ModuleList([_Fuse([d, d], d, conv_type, weighted_fusion) for _ in range(n - 1)])
                                       for _ in range(num_layers)])
        self.bottom_up = nn.ModuleList([nn.ModuleList([_Fuse([d, d, d] if i < n - 2 else [d, d], d, conv_type, weighted_fusion)
                                                       for i in range(n - 1)]) for _ in range(num_layers)])

    def get_out_channels(self):
        return self.out_channels


_NECKS = {"FPN": "FPN", "fpn": "FPN", "FPNNeck": "FPN", "simple": "simple", "SimpleNeck": "simple", "ida": "ida", "IDANeck": "ida",
          "bifpn": "bifpn", "BiFPNNeck": "bifpn"}


class _Head(nn.Module):
    def __init__(self, cin, cout, width=256, depth=3, init_bias=None):
        super().__init__()
        for i in range(depth):
            self.add_module(f"block_{i + 1}", _ConvBn(cin if i == 0 else width, width))
        self.out_conv = nn.Conv2d(width, cout, 1)
        if init_bias is not None:
            self.out_conv.bias.data.fill_(init_bias)          # reference models/meta.py:29-30
        self.depth = depth


class EngineModel(nn.Module):
    """Stands where the reference's GenericModel stands (``CenterNet.model``): same parameters, same call contract
    ``model(images) -> Dict[str, Tensor]`` of raw head outputs, executed by the sm_100a engine."""

    def __init__(self, backbone: _Backbone, neck: nn.Module, heads: nn.Module, precision: int, backbone_name: str = "resnet34",
                 neck_name: str = "FPN"):
        super().__init__()
        self.backbone, self.neck, self.heads = backbone, neck, heads
        self.backbone_name = backbone_name
        self.neck_name = neck_name
        self.precision = precision
        # one engine (plan + packed weights + activation arena: 7.6 GB at 32 x 512^2) per input shape, least recently used
        # evicted beyond `max_engines` (ragged last batches / multi-scale inputs would otherwise accumulate arenas)
        self._engines: "OrderedDict[Tuple, Engine]" = OrderedDict()
        self.max_engines = 4
        self.generation = 0            # bumped whenever the parameters may have changed: keys every cached plan / graph
        self.register_load_state_dict_post_hook(lambda m, keys: m.invalidate())

    def invalidate(self) -> None:
        """Drop packed weights / plans (call after changing parameters in place; load_state_dict does it itself)."""
        for e in self._engines.values():
            e.close()
        self._engines.clear()
        self.generation += 1

    def head_names(self) -> List[str]:
        return [n for n, _ in self.heads.named_children()]

    def engine_for(self, images: torch.Tensor) -> Engine:
        n, c, h, w = images.shape
        key = (n, h, w, images.device.index, self.precision)
        eng = self._engines.get(key)
        if eng is None:
            while len(self._engines) >= max(1, self.max_engines):
                _, old = self._engines.popitem(last=False)
                old.close()
            names = self.head_names()
            depth = getattr(self.heads, names[0]).depth
            plan = build_plan(self.state_dict(), backbone=self.backbone_name, neck=self.neck_name, head_names=names,
                              head_depth=depth)
            eng = Engine(plan, n, h, w, images.device, precision=self.precision)
            self._engines[key] = eng
        else:
            self._engines.move_to_end(key)
        return eng

    def forward(self, images: torch.Tensor, sigmoid_head: Optional[str] = None) -> Dict[str, torch.Tensor]:
        """``sigmoid_head`` (an extension): that head is returned as probabilities, the logistic fused into its out-conv."""
        if images.dim() != 4 or images.shape[1] != 3:
            raise ValueError(f"images must be (N,3,H,W), got {tuple(images.shape)}")
        if not images.is_cuda:
            raise RuntimeError("images are on the CPU: the cnl_b200 forward runs on CUDA (sm_100a) only, there is no fallback")
        images = images.float().contiguous()
        out = self.engine_for(images).forward(images, sigmoid_head=sigmoid_head)
        # fresh tensors, as the reference's GenericModel returns: the engine's own buffers are overwritten by the next call
        # with this shape (detect() reads them in place instead)
        return {k: v.clone() for k, v in out.items()}


_Det = namedtuple("EncodedOutputs", ["heatmap", "box_2d"])
_Trk = namedtuple("EncodedOutputs", ["heatmap", "box_2d", "reid"])


class CenterNet(nn.Module):
    def __init__(self, num_classes: int, backbone: str = "resnet34", pretrained_backbone: bool = False, neck: str = "FPN",
                 neck_config: Optional[Dict[str, Any]] = None, head_config: Optional[Dict[str, Any]] = None,
                 box_init_bias: Optional[float] = None, box_log: bool = False, box_multiplier: float = 1.0,
                 heatmap_prior: float = 0.01, nms_kernel: int = 3, num_detections: int = 100,
                 reid_dim: int = 0, precision: str = "split", **training_only: Any):
        super().__init__()
        if pretrained_backbone:
            raise RuntimeError("pretrained_backbone=True needs a download; load weights with load_state_dict() instead")
        if neck not in _NECKS:
            raise ValueError(f"neck {neck!r}: the sm_100a engine lowers {sorted(set(_NECKS.values()))} (SURVEY 8a F2, 8f rank 4)")
        neck = _NECKS[neck]
        if nms_kernel % 2 != 1:
            raise ValueError("nms_kernel must be odd")
        neck_config = dict(neck_config or {})
        head_config = dict(head_config or {})
        self.hparams = SimpleNamespace(num_classes=num_classes, backbone=backbone, neck=neck, neck_config=neck_config,
                                       head_config=head_config, box_log=box_log, box_multiplier=box_multiplier,
                                       heatmap_prior=heatmap_prior, nms_kernel=nms_kernel, num_detections=num_detections,
                                       reid_dim=reid_dim, precision=precision, **training_only)
        bb = _MobileNetV2() if backbone == "mobilenet_v2" else _Backbone(backbone)
        nk = {"FPN": _FPN, "simple": _SimpleNeck, "ida": _IDANeck, "bifpn": _BiFPNNeck}[neck](bb.get_out_channels(), **neck_config)
        heads = nn.Module()
        c = nk.get_out_channels()
        heads.add_module("heatmap", _Head(c, num_classes, init_bias=math.log(heatmap_prior / (1 - heatmap_prior)), **head_config))
        heads.add_module("box_2d", _Head(c, 4, init_bias=box_init_bias, **head_config))
        if reid_dim:
            heads.add_module("reid", _Head(c, reid_dim, **head_config))
        self.model = EngineModel(bb, nk, heads, _PRECISIONS[precision], backbone, neck)
        self.stride = bb.stride // nk.stride                                   # reference models/meta.py:96
        self.num_classes = num_classes
        self._graphs: "OrderedDict[Tuple, Any]" = OrderedDict()
        self._stage: Dict[Tuple, List[Optional[torch.Tensor]]] = {}      # detect_host_batches: device staging buffers per shape
        self.eval()

    # ---- G1 names -------------------------------------------------------------------------------------------
    @property
    def output_stride(self) -> int:
        return self.stride

    def get_encoded_outputs(self, images: torch.Tensor) -> Dict[str, torch.Tensor]:
        return self.model(images)

    get_output_dict = get_encoded_outputs

    def forward(self, images: torch.Tensor):
        """G1 contract: (sigmoid(heatmap), box_2d[, reid])  (reference tests/test_models.py:88-99, models/tracker.py:100)."""
        out = self.model(images, sigmoid_head="heatmap")          # the logistic runs in the heat-map out-conv's epilogue
        if "reid" in out:
            return _Trk(out["heatmap"], out["box_2d"], out["reid"])
        return _Det(out["heatmap"], out["box_2d"])

    # ---- decode (reference models/centernet.py:229-304) ----------------------------------------------------
    def decode_detections(self, heatmap: torch.Tensor, box_offsets: torch.Tensor, normalize_boxes: bool = False,
                          from_logits: bool = False) -> Dict[str, torch.Tensor]:
        hp = self.hparams
        out = _decode.decode_detections(heatmap, box_offsets, num_detections=hp.num_detections, nms_kernel=hp.nms_kernel,
                                        normalize_boxes=normalize_boxes, box_log=hp.box_log,
                                        box_multiplier=hp.box_multiplier, stride=self.stride, from_logits=from_logits)
        return {"boxes": out["boxes"], "scores": out["scores"], "labels": out["labels"]}

    def get_topk_from_heatmap(self, heatmap: torch.Tensor, pseudo_nms: bool = True):
        return _decode.get_topk_from_heatmap(heatmap, self.hparams.num_detections, self.hparams.nms_kernel, pseudo_nms)

    gather_and_decode_boxes = staticmethod(_decode.gather_and_decode_boxes)

    def gather_detection2d(self, heatmap, box_2d=None, num_detections: Optional[int] = None, nms_kernel: Optional[int] = None,
                           normalize_bbox: bool = False) -> Dict[str, torch.Tensor]:
        """G1 decode: probabilities + ltrb map (or the tuple/dict forward() returned) -> {"bboxes","labels","scores"}."""
        heatmap, box_2d, _ = _unpack(heatmap, box_2d, None)
        hp = self.hparams
        out = _decode.decode_detections(heatmap, box_2d, num_detections=num_detections or hp.num_detections,
                                        nms_kernel=nms_kernel or hp.nms_kernel, normalize_boxes=normalize_bbox,
                                        box_log=hp.box_log, box_multiplier=hp.box_multiplier, stride=self.stride)
        return {"bboxes": out["boxes"], "labels": out["labels"], "scores": out["scores"]}

    def gather_tracking2d(self, heatmap, box_2d=None, reid=None, num_detections: int = 100, nms_kernel: int = 3,
                          normalize_bbox: bool = False) -> Dict[str, torch.Tensor]:
        """reference models/fairmot.py:138-151: + "embeddings" (N,k,E)."""
        heatmap, box_2d, reid = _unpack(heatmap, box_2d, reid)
        hp = self.hparams
        out = _decode.decode_detections(heatmap, box_2d, reid=reid, num_detections=num_detections, nms_kernel=nms_kernel,
                                        normalize_boxes=normalize_bbox, box_log=hp.box_log,
                                        box_multiplier=hp.box_multiplier, stride=self.stride)
        return {"bboxes": out["boxes"], "labels": out["labels"], "scores": out["scores"], "embeddings": out["embeddings"]}

    # ---- fused forward + decode ----------------------------------------------------------------------------
    @torch.no_grad()
    def detect(self, images: torch.Tensor, normalize_boxes: bool = False, use_graph: bool = True,
               static_input: bool = False, packed_out: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """images (N,3,H,W) float32 on the GPU -> {"boxes","scores","labels"[, "embeddings"]}: what the reference's
        validation_step computes at models/centernet.py:204-205, as one CUDA-graph replay of this package's kernels.
        Returned tensors are static buffers overwritten by the next call with the same shape.

        ``static_input``: the caller promises that ``images`` is a long-lived buffer it refills in place (a double-buffered
        loader); the graph then reads it directly instead of copying it into a private input first.
        ``packed_out``: (N, k, 8 + reid_dim) float32 buffer that additionally receives the packed rows of
        distributed.DetectionGather (written by the select kernel)."""
        if not images.is_cuda:
            raise RuntimeError("images are on the CPU: copy them to the GPU first (no CPU fallback)")
        if images.dtype != torch.float32 or not images.is_contiguous():
            if static_input:
                raise ValueError("static_input needs a contiguous float32 tensor")
            images = images.float().contiguous()
        hp = self.hparams
        # everything a captured graph freezes is part of its key: shape, weights generation, decode hyper-parameters
        key = (tuple(images.shape), images.device.index, bool(normalize_boxes), self.model.precision, self.model.generation,
               int(hp.num_detections), int(hp.nms_kernel), bool(hp.box_log), float(hp.box_multiplier), bool(use_graph),
               images.data_ptr() if static_input else None, packed_out.data_ptr() if packed_out is not None else None)
        g = self._graphs.get(key)
        if g is not None and g.engine.handle is None:       # its engine was evicted (arena released): the graph is dead
            del self._graphs[key]
            g = None
        if g is None:
            for old in [k for k in self._graphs if k[4] != self.model.generation]:      # graphs of replaced weights
                del self._graphs[old]
            while len(self._graphs) >= 2 * max(1, self.model.max_engines):
                self._graphs.popitem(last=False)
            g = _DetectGraph(self, images, normalize_boxes, use_graph, static_input, packed_out)
            self._graphs[key] = g
        else:
            self._graphs.move_to_end(key)
        return g.run(images)

    @torch.no_grad()
    def detect_host_batches(self, host_batches, device: Optional[torch.device] = None):
        """Generator over detections for an iterable of PINNED host batches (N,3,H,W) float32.  The H2D copy of batch
        i+1 runs on a side stream while batch i is computed (double-buffered device inputs); results are copied into
        pinned host tensors and yielded once they have landed (the yielded dict is reused every second call)."""
        dev = device or next(self.parameters()).device
        copy_stream = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        copy_stream.wait_stream(main)          # an abandoned earlier generator may still be reading the shared staging buffers
        ready, consumed, host_out, out_done = [None, None], [None, None], [None, None], [None, None]
        it = iter(host_batches)

        def stage(slot, hb):
            # the two device staging buffers of a shape live with the model: the captured graphs read them in place
            # (static_input), so they must be the same tensors on every call
            dev_in = self._stage.setdefault((tuple(hb.shape), dev.index), [None, None])
            if dev_in[slot] is None:
                dev_in[slot] = torch.empty(hb.shape, dtype=torch.float32, device=dev)
            with torch.cuda.stream(copy_stream):
                if consumed[slot] is not None:
                    copy_stream.wait_event(consumed[slot])
                dev_in[slot].copy_(hb, non_blocking=True)
                ready[slot] = torch.cuda.Event()
                ready[slot].record(copy_stream)

        nxt = next(it, None)
        if nxt is not None:
            stage(0, nxt)
        i = 0
        while nxt is not None:
            slot = i & 1
            cur = nxt
            nxt = next(it, None)
            if nxt is not None:
                stage(slot ^ 1, nxt)
            main.wait_event(ready[slot])
            det = self.detect(self._stage[(tuple(cur.shape), dev.index)][slot], static_input=True)
            consumed[slot] = torch.cuda.Event()
            consumed[slot].record(main)
            if host_out[slot] is None:
                host_out[slot] = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in det.items()}
            for k, v in det.items():
                host_out[slot][k].copy_(v, non_blocking=True)
            out_done[slot] = torch.cuda.Event()
            out_done[slot].record(main)
            if i > 0:
                out_done[slot ^ 1].synchronize()
                yield host_out[slot ^ 1]
            i += 1
        if i > 0:
            out_done[(i - 1) & 1].synchronize()
            yield host_out[(i - 1) & 1]

    @torch.no_grad()
    def predict_step(self, images: torch.Tensor) -> List[Dict[str, Any]]:
        """The inference half of the reference's validation_step (models/centernet.py:202-209): forward, decode,
        boxes xyxy -> xywh (COCO format), one dict of numpy arrays per image - ready for CocoEvaluator.update."""
        det = self.detect(images)
        boxes = _decode.boxes_xyxy_to_xywh(det["boxes"])
        host = {"boxes": boxes.cpu().numpy(), "scores": det["scores"].cpu().numpy(), "labels": det["labels"].cpu().numpy()}
        if "embeddings" in det:
            host["embeddings"] = det["embeddings"].cpu().numpy()
        return [{k: v[i] for k, v in host.items()} for i in range(images.shape[0])]

    # ---- validation tail (reference models/centernet.py:202-218) ------------------------------------------------
    # Any object with update(preds, targets) / get_metrics() / reset().  The reference builds CocoEvaluator(num_classes) in its
    # constructor (models/centernet.py:115); here the pycocotools-free evaluate.CocoEvaluator is created on first use.
    evaluator = None

    def validation_step(self, batch, batch_idx: int = 0):
        """``images, targets = batch``: forward + decode + xyxy->xywh on the GPU (predict_step), per-image numpy dicts,
        targets filtered to ``boxes`` / ``labels`` and handed to ``self.evaluator.update`` exactly as the reference does
        (models/centernet.py:202-212)."""
        import numpy as np
        images, targets = batch
        preds = self.predict_step(images.to(next(self.parameters()).device))
        targets = [{k: np.array(t[k]) for k in ("boxes", "labels")} for t in targets]
        if self.evaluator is None:
            from .evaluate import CocoEvaluator
            self.evaluator = CocoEvaluator(self.num_classes)
        self.evaluator.update(preds, targets)
        return preds

    def validation_epoch_end(self, outputs=None) -> Dict[str, float]:
        """reference models/centernet.py:214-218: metrics of every rank's predictions (gather_and_merge), prefixed val/."""
        metrics = self.evaluator.get_metrics()
        self.evaluator.reset()
        return {f"val/{k}": v for k, v in metrics.items()}

    def invalidate(self) -> None:
        """Drop every cached graph, plan and packed weight set (after changing parameters in place)."""
        self._graphs.clear()
        self._stage.clear()
        self.model.invalidate()

    @torch.no_grad()
    def init_synthetic_(self, seed: int = 0) -> "CenterNet":
        """Deterministic random weights with healthy activation statistics (no network access for checkpoints):
        He-normal convs, BN gamma ~ U(0.5,1.5) (residual-branch bn2 damped), beta ~ N(0,0.1), unit running variance.
        Dense throughput does not depend on the values, but dead (all-zero) activations would flatter power/clocks."""
        g = torch.Generator().manual_seed(seed)
        for name, mod in self.model.named_modules():
            if isinstance(mod, nn.Conv2d):
                fan_in = mod.in_channels // mod.groups * mod.kernel_size[0] * mod.kernel_size[1]
                mod.weight.copy_(torch.randn(mod.weight.shape, generator=g) * math.sqrt(2.0 / fan_in))
                if name.endswith("out_conv"):
                    mod.weight.mul_(0.2)             # head logits with std ~1.5 around the prior, like a trained heatmap
                if mod.bias is not None and not name.endswith("heatmap.out_conv"):
                    mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.1)
            elif isinstance(mod, nn.BatchNorm2d):
                mod.weight.copy_(torch.rand(mod.weight.shape, generator=g) + 0.5)
                if name.endswith("bn3" if self.hparams.backbone == "resnet50" else "bn2") and ".layer" in name:
                    mod.weight.mul_(0.3)
                mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.1)
                mod.running_mean.zero_()
                mod.running_var.fill_(1.0)
        self.invalidate()
        return self

    # ---- checkpoints -----------------------------------------------------------------------------------------
    def load_reference_state_dict(self, state_dict: Dict[str, torch.Tensor], key_map: Optional[Dict[str, str]] = None,
                                  strict: bool = True):
        """Load a reference checkpoint's ``state_dict`` (Lightning layout: ``model.backbone.*``, ``model.neck.*``,
        ``model.heads.<head>.block_<i>.*`` / ``.out_conv.*``; reference models/meta.py:26-28,36-38,92-95).

        The names INSIDE ``backbone`` / ``neck`` / ``ConvBnAct`` belong to vision_toolbox, which is not part of the
        reference snapshot; this package uses torchvision's ResNet names, ``neck.lateral.<i>``,
        ``neck.output.<i>.{conv,bn}`` and ``block_<i>.{conv,bn}``.  ``key_map`` = ordered ``{regex: replacement}`` rules
        (``re.sub``) applied to every key bridges a checkpoint with other inner names.  Keys of modules that do not exist
        at inference (``fc.*``, ``classifier.*``, loss / evaluator state) are dropped; anything else that does not match
        raises under ``strict``."""
        import re
        sd = {}
        for k, v in state_dict.items():
            for pat, rep in (key_map or {}).items():
                k = re.sub(pat, rep, k)
            if re.search(r"(^|\.)(fc|classifier|evaluator|loss\w*)\.", k):
                continue
            sd[k] = v
        return self.load_state_dict(sd, strict=strict)

    # ---- folder inference (reference README.md:49-65) ------------------------------------------------------
    @torch.no_grad()
    def inference_detection(self, img_dir: str, img_names: Optional[Sequence[str]] = None, batch_size: int = 4,
                            num_detections: Optional[int] = None, img_size: int = 512, device: str = "cuda:0",
                            workers: Optional[int] = None):
        from .inference import run_folder
        return run_folder(self, img_dir, img_names, batch_size, num_detections, img_size, torch.device(device), workers)


def _unpack(heatmap, box_2d, reid):
    if isinstance(heatmap, dict):
        return heatmap["heatmap"], heatmap["box_2d"], heatmap.get("reid", reid)
    if isinstance(heatmap, (tuple, list)):
        t = tuple(heatmap)
        return t[0], t[1], (t[2] if len(t) > 2 else reid)
    return heatmap, box_2d, reid


class _DetectGraph:
    """forward + decode for one input shape, captured once into a CUDA graph with static in/out buffers."""

    def __init__(self, net: CenterNet, example: torch.Tensor, normalize_boxes: bool, use_graph: bool,
                 static_input: bool = False, packed_out: Optional[torch.Tensor] = None):
        self.net = net
        dev = example.device
        hp = net.hparams
        self.engine = net.model.engine_for(example)
        n = example.shape[0]
        h, w = example.shape[2] // net.stride, example.shape[3] // net.stride
        k = hp.num_detections
        self.bound = static_input                           # the graph reads the caller's own buffer
        self.static_in = example if static_input else torch.empty_like(example)
        outs = self.engine.outputs
        reid = outs.get("reid")
        self.bufs = _decode.DecodeBuffers(n, h, w, k, reid.shape[1] if reid is not None else 0, dev, packed=packed_out)
        # the heat-map out-conv stores probabilities (the reference's `.sigmoid()`, models/centernet.py:205, in its epilogue),
        # so the decode is the reference's own probability-space decode: no second pass over the map, no logit-space shortcut
        self.kw = dict(num_detections=k, nms_kernel=hp.nms_kernel, normalize_boxes=normalize_boxes, box_log=hp.box_log,
                       box_multiplier=hp.box_multiplier, stride=net.stride, from_logits=False)
        self.launches = 0
        self.graph = None
        if not static_input:
            self.static_in.copy_(example)
        with torch.cuda.device(dev):
            self._body()                                        # warm-up (also sets function attributes)
            torch.cuda.synchronize(dev)
            if use_graph:
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    self._body()
                torch.cuda.current_stream(dev).wait_stream(side)
                torch.cuda.synchronize(dev)
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self._body()

    def _body(self):
        outs = self.engine.forward(self.static_in, sigmoid_head="heatmap")
        self.launches = self.engine.last_launches + _decode.decode_into(self.bufs, outs["heatmap"], outs["box_2d"],
                                                                         outs.get("reid"), **self.kw)

    def run(self, images: torch.Tensor) -> Dict[str, torch.Tensor]:
        if not self.bound:
            self.static_in.copy_(images, non_blocking=True)
        with torch.cuda.device(self.static_in.device):
            if self.graph is not None:
                self.graph.replay()
            else:
                self._body()
        return self.bufs.as_dict()
