"""Import the UNMODIFIED reference decode (and GenericModel/GenericHead) under stub modules.

TEST INFRASTRUCTURE ONLY.  Works only where /root/reference exists (the authoring
container); used by tests/golden/gen_golden.py to produce the committed golden
vectors and by tests/test_oracle.py (skipped when the reference is absent, e.g. on
the GPU box).  Nothing here copies reference source: the reference's own files are
imported from where they lie.

The reference imports pytorch_lightning, vision_toolbox, albumentations and
pycocotools at module import time (models/centernet.py:12-15, models/meta.py:7-10);
none is installed, so empty stand-ins are registered in sys.modules first.
"""
from __future__ import annotations

import os
import sys
import types
from types import SimpleNamespace

REFERENCE_ROOT = os.environ.get("CNL_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "centernet_lightning"))


def _stub(name: str, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []          # behaves as a package for "import a.b"
    sys.modules[name] = m
    return m


def import_reference():
    """Returns the reference's ``centernet_lightning.models.centernet`` and ``.meta`` modules."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    from torch import nn

    class _LightningModule(nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

    _stub("pytorch_lightning", LightningModule=_LightningModule)
    from oracle import spec_model as sm
    vt = _stub("vision_toolbox")
    vt.backbones = _stub("vision_toolbox.backbones", BaseBackbone=nn.Module)
    vt.necks = _stub("vision_toolbox.necks", BaseNeck=nn.Module)
    vt.components = _stub("vision_toolbox.components", ConvBnAct=sm.ConvBnAct)
    a = _stub("albumentations", Compose=object, OneOf=object, BboxParams=object)
    a.pytorch = _stub("albumentations.pytorch", ToTensorV2=object)
    pc = _stub("pycocotools")
    pc.coco = _stub("pycocotools.coco", COCO=object)
    pc.cocoeval = _stub("pycocotools.cocoeval", COCOeval=object)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    cn = importlib.import_module("centernet_lightning.models.centernet")
    meta = importlib.import_module("centernet_lightning.models.meta")
    return cn, meta


def import_reference_layers():
    """The reference's own ``centernet_lightning.models.layers`` (make_conv :40-79, make_upsample :81-99, Fuse :138-177);
    it only needs torch + torchvision.ops.DeformConv2d, both installed."""
    import_reference()
    import importlib
    return importlib.import_module("centernet_lightning.models.layers")


def reference_decode(heatmap, box_offsets, *, num_detections=100, nms_kernel=3, normalize_boxes=False,
                     box_log=False, box_multiplier=1.0, stride=4):
    """Run the reference's own CenterNet.decode_detections (models/centernet.py:229-241) unbound,
    with a SimpleNamespace standing in for ``self`` (only .hparams and .stride are read)."""
    cn, _ = import_reference()
    ns = SimpleNamespace(
        hparams=SimpleNamespace(nms_kernel=nms_kernel, num_detections=num_detections,
                                box_log=box_log, box_multiplier=box_multiplier),
        stride=stride)
    ns.get_topk_from_heatmap = lambda h, pseudo_nms=True: cn.CenterNet.get_topk_from_heatmap(ns, h, pseudo_nms)
    out = cn.CenterNet.decode_detections(ns, heatmap, box_offsets, normalize_boxes=normalize_boxes)
    _, indices, _ = cn.CenterNet.get_topk_from_heatmap(ns, heatmap)
    out["indices"] = indices
    return out


def reference_generic_model(spec):
    """Wrap a spec model's backbone/neck with the reference's own GenericModel and rebuild its
    heads with the reference's own GenericHead (models/meta.py:21-47), copying the weights."""
    _, meta = import_reference()
    from torch import nn
    heads = nn.Module()
    for name, h in spec.heads.named_children():
        depth = len([k for k, _ in h.named_children() if k.startswith("block_")])
        rh = meta.GenericHead(h.block_1.conv.in_channels, h.out_conv.out_channels,
                              width=h.out_conv.in_channels, depth=depth)
        rh.load_state_dict(h.state_dict())
        heads.add_module(name, rh)
    return meta.GenericModel(spec.backbone, spec.neck, heads).eval()
