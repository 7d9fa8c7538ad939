"""Seeded input recipes shared by the golden generator and the tests (inputs are regenerated, never stored)."""
from __future__ import annotations

import torch


def _c(name, n, c, h, w, k=100, nms=3, seed=0, logits=True, normalize=False, box_log=False, mult=16.0, stride=4,
       reid=0, kind="randn"):
    return dict(name=name, n=n, c=c, h=h, w=w, k=k, nms=nms, seed=seed, logits=logits, normalize=normalize,
                box_log=box_log, mult=mult, stride=stride, reid=reid, kind=kind)


# SURVEY 8c "golden vectors to create": seeds x shapes x k x nms_kernel x box_log x normalize + adversarial maps.
DECODE_CASES = [
    _c("coco128_s0", 2, 80, 128, 128, seed=0),
    _c("coco128_s1", 2, 80, 128, 128, seed=1),
    _c("coco128_s2_k300", 1, 80, 128, 128, seed=2, k=300),
    _c("track128", 1, 2, 128, 128, seed=0, reid=64),
    _c("big256", 1, 80, 256, 256, seed=0),
    _c("odd17x23", 3, 5, 17, 23, seed=0, k=50),
    _c("odd17x23_k1", 3, 5, 17, 23, seed=1, k=1),
    _c("nms1", 2, 7, 32, 64, seed=0, nms=1, k=64),
    _c("nms5", 2, 7, 32, 64, seed=1, nms=5, k=20),
    _c("nms5_odd", 2, 3, 19, 21, seed=2, nms=5, k=9),
    _c("nms7_odd", 1, 3, 19, 21, seed=3, nms=7, k=4),
    _c("boxlog", 2, 80, 64, 64, seed=0, box_log=True, mult=1.0),
    _c("normalize", 2, 80, 64, 64, seed=1, normalize=True),
    _c("wide272", 1, 4, 8, 272, seed=0, k=30),
    _c("k_eq_hw", 1, 3, 6, 8, seed=0, k=48),
    _c("plateau", 1, 3, 16, 16, seed=0, k=10, kind="plateau"),
    _c("saturated", 1, 4, 16, 16, seed=0, k=10, kind="saturated"),
    _c("corners", 1, 2, 16, 32, seed=0, k=8, kind="corners"),
    _c("probs_direct", 2, 6, 32, 32, seed=4, k=40, logits=False),
]
DECODE_BY_NAME = {c["name"]: c for c in DECODE_CASES}


def make_decode_inputs(case):
    """(heat, box, reid).  heat = logits when case['logits'] else probabilities in (0,1)."""
    g = torch.Generator().manual_seed(case["seed"])
    n, c, h, w = case["n"], case["c"], case["h"], case["w"]
    kind = case["kind"]
    if kind == "randn":
        heat = torch.randn((n, c, h, w), generator=g) * 1.5 - 2.19          # SURVEY 8d decode-only benchmark recipe
    elif kind == "plateau":                                                   # every pixel equal: all are "peaks"
        heat = torch.full((n, c, h, w), -1.0)
        heat[:, 1, 4:9, 4:9] = 0.5                                            # a flat 5x5 mesa in class 1
    elif kind == "saturated":                                                 # |logit| > 20: sigmoid collapses to 0 / 1
        heat = torch.randn((n, c, h, w), generator=g) * 30.0
    elif kind == "corners":                                                   # one peak in every corner / border
        heat = torch.full((n, c, h, w), -8.0)
        for i, (y, x) in enumerate([(0, 0), (0, w - 1), (h - 1, 0), (h - 1, w - 1), (0, w // 2), (h // 2, 0)]):
            heat[:, i % c, y, x] = 1.0 + 0.25 * i
    else:
        raise ValueError(kind)
    if not case["logits"]:
        heat = torch.rand((n, c, h, w), generator=g)
    box = torch.randn((n, 4, h, w), generator=g)
    if case["box_log"]:
        box = box * 0.5
    reid = torch.randn((n, case["reid"], h, w), generator=g) if case["reid"] else None
    return heat, box, reid


FORWARD_CASES = {
    "det64": dict(model=dict(num_classes=80), seed=0, n=2, size=64, img_seed=7),
    "track64": dict(model=dict(num_classes=2, reid_dim=64), seed=1, n=1, size=64, img_seed=8),
    # BASELINE configs[0] (configs/base_resnet34.yaml): ResNet-34 + "simple" neck (no FPN), nearest x2 upsampling
    "simple64": dict(model=dict(num_classes=80, neck="simple"), seed=2, n=2, size=64, img_seed=9),
    # the same neck with the reference's ConvTranspose2d-BN-ReLU upsample layers (models/layers.py:86-96), k = 3 and 4
    "deconv3_64": dict(model=dict(num_classes=80, neck="simple", neck_config=dict(upsample_type="conv_transpose", deconv_kernel=3)),
                       seed=3, n=1, size=64, img_seed=10),
    "deconv4_64": dict(model=dict(num_classes=80, neck="simple", neck_config=dict(upsample_type="conv_transpose", deconv_kernel=4)),
                       seed=4, n=1, size=64, img_seed=11),
    # SURVEY 8f rank 4 remainder (reference tests/test_models.py:37-39): mobilenet_v2 trunk (depthwise + pointwise stages, ReLU6),
    # IDA / BiFPN necks from the reference's Fuse node (models/layers.py:138-177), make_conv's separable branch (:56-68)
    "mbv2_fpn64": dict(model=dict(num_classes=8, backbone="mobilenet_v2", neck_config=dict(out_channels=64), head_config=dict(width=64, depth=1)),
                       seed=5, n=2, size=64, img_seed=12, init=dict(out_gain=1.0)),
    "mbv2_simple64": dict(model=dict(num_classes=8, backbone="mobilenet_v2", neck="simple", head_config=dict(width=64, depth=1)),
                          seed=6, n=1, size=64, img_seed=13, init=dict(out_gain=1.0)),
    "r18_ida64": dict(model=dict(num_classes=8, backbone="resnet18", neck="ida", head_config=dict(width=64, depth=1)),
                      seed=7, n=2, size=64, img_seed=14, init=dict(out_gain=1.0)),
    "r18_bifpn128": dict(model=dict(num_classes=8, backbone="resnet18", neck="bifpn", neck_config=dict(out_channels=64, num_layers=2),
                                    head_config=dict(width=64, depth=1)), seed=8, n=1, size=128, img_seed=15, init=dict(out_gain=1.0)),
    "mbv2_ida64_sep": dict(model=dict(num_classes=8, backbone="mobilenet_v2", neck="ida", neck_config=dict(conv_type="separable", weighted_fusion=True),
                                      head_config=dict(width=64, depth=1)), seed=9, n=2, size=64, img_seed=16, init=dict(out_gain=1.0)),
    "mbv2_bifpn64_sep": dict(model=dict(num_classes=8, backbone="mobilenet_v2", neck="bifpn", neck_config=dict(conv_type="separable", num_layers=1),
                                        head_config=dict(width=64, depth=1)), seed=10, n=1, size=64, img_seed=17, init=dict(out_gain=1.0)),
}


# ---- near-tie logits (from_logits path): distinct logits that fp32 sigmoid maps to ONE probability ---------------------
# The reference compares probabilities (models/centernet.py:252-254), so such neighbours are a plateau (both kept) and such
# classes tie (first class wins).  Every structure is planted on a cleared patch of its class so that it is a local maximum.
NEARTIE_CASES = [
    dict(name="neartie_w128", n=2, c=40, h=16, w=128, k=60, seed=11),      # one warp tile per row, 4 class groups
    dict(name="neartie_w256", n=1, c=33, h=12, w=256, k=60, seed=12),      # VEC = 2 tiles
    dict(name="neartie_w272", n=1, c=8, h=12, w=272, k=60, seed=13),       # rows wider than a warp tile (halo columns), 2 groups
    dict(name="neartie_odd", n=2, c=5, h=19, w=21, k=40, seed=14),         # generic kernel geometry
    dict(name="neartie_c1", n=1, c=1, h=16, w=128, k=30, seed=15),         # single class, single group
]


def _ulps(x: float, n: int) -> float:
    import numpy as np
    v = np.float32(x)
    for _ in range(abs(n)):
        v = np.nextafter(v, np.float32(np.inf if n > 0 else -np.inf), dtype=np.float32)
    return float(v)


def make_neartie_logits(case):
    """(logits, box): randn background with planted pairs / triples whose members differ by a few ulps up to the widest
    gap that can still round to one probability, at logit levels from -3 to saturation."""
    g = torch.Generator().manual_seed(case["seed"])
    n, c, h, w = case["n"], case["c"], case["h"], case["w"]
    heat = torch.randn((n, c, h, w), generator=g) * 1.5 - 2.19
    levels = [-3.0, -0.5, 0.3, 1.7, 3.0, 5.0, 8.0, 12.0, 15.0, 16.0, 16.7, 20.0, 30.0]
    gaps = [1, 2, 5, 11, 40, 200, 900]                                   # in ulps of the smaller logit
    extra = [(8.0, 8.00001), (12.0, 12.0005), (3.0, _ulps(3.0, 5)), (16.5, 16.9), (17.5, 40.0)]
    pairs = [(a, _ulps(a, gaps[(i + j) % len(gaps)])) for i, a in enumerate(levels) for j in range(2)] + extra
    slots = [(y, x) for y in range(2, h - 2, 5) for x in range(2, w - 4, 7)]
    rng = torch.Generator().manual_seed(case["seed"] + 1000)
    order = torch.randperm(len(slots), generator=rng).tolist()
    pairs = [pairs[i] for i in torch.randperm(len(pairs), generator=rng).tolist()]      # small maps get a mix of levels
    for i, (lo, hi) in enumerate(pairs):
        if i >= len(order):
            break
        y, x = slots[order[i]]
        img = i % n
        cls = (3 * i) % c
        heat[img, cls, y - 2:y + 3, x - 2:x + 5] = -30.0
        shape = i % 5
        if shape == 0:                                                   # horizontal neighbours, smaller first
            heat[img, cls, y, x], heat[img, cls, y, x + 1] = lo, hi
        elif shape == 1:                                                 # vertical neighbours, larger first
            heat[img, cls, y, x], heat[img, cls, y + 1, x] = hi, lo
        elif shape == 2:                                                 # diagonal + a third member in between
            heat[img, cls, y, x], heat[img, cls, y + 1, x + 1], heat[img, cls, y, x + 1] = lo, hi, (lo + hi) / 2
        elif shape == 3 and c > 1:                                       # same pixel, two classes (far apart = different groups):
            other = (cls + c // 2 + 1) % c                               # the LOWER class index holds the smaller logit
            a, b = min(cls, other), max(cls, other)
            heat[img, b, y - 2:y + 3, x - 2:x + 5] = -30.0
            heat[img, a, y, x], heat[img, b, y, x] = lo, hi
        else:                                                            # same pixel, neighbouring classes (same group mostly)
            other = (cls + 1) % c
            a, b = min(cls, other), max(cls, other)
            heat[img, b, y - 2:y + 3, x - 2:x + 5] = -30.0
            heat[img, a, y, x], heat[img, b, y, x] = hi, lo
    box = torch.randn((n, 4, h, w), generator=g)
    return heat.contiguous(), box


def make_image(kw):
    g = torch.Generator().manual_seed(kw["img_seed"])
    return torch.rand((kw["n"], 3, kw["size"], kw["size"]), generator=g)


# ---- tracker sequences: seeded detection streams for Tracker.update --------------------------------
TRACK_CASES = {
    "walk_iou": dict(seed=0, frames=40, objects=6, dim=64, tracker=dict(box_cost="iou")),
    "walk_giou": dict(seed=1, frames=40, objects=8, dim=64, tracker=dict(box_cost="giou", reid_threshold=0.3)),
    "reid_only": dict(seed=2, frames=30, objects=5, dim=16, tracker=dict(box_cost=None, min_birth_age=1, max_inactive_age=3)),
    "crowded": dict(seed=3, frames=25, objects=20, dim=64, tracker=dict(detection_threshold=0.4, smoothing_factor=0.9)),
}


def make_track_sequence(case):
    """List of per-frame (bboxes (k,4) f32 normalised xyxy, labels (k,) i64, scores (k,) f32 descending, embeddings (k,E) f32):
    objects drift with constant velocity, are missed at random, and false positives with random appearance are mixed in."""
    import numpy as np
    rng = np.random.default_rng(case["seed"])
    m, e = case["objects"], case["dim"]
    centre = rng.uniform(0.2, 0.8, (m, 2))
    vel = rng.normal(0, 0.006, (m, 2))
    size = rng.uniform(0.05, 0.15, (m, 2))
    ident = rng.standard_normal((m, e))
    frames = []
    for _ in range(case["frames"]):
        centre = centre + vel
        seen = rng.random(m) > 0.15
        boxes = np.concatenate([centre - size / 2, centre + size / 2], axis=1)[seen] + rng.normal(0, 0.003, (int(seen.sum()), 4))
        emb = (ident + 0.15 * rng.standard_normal((m, e)))[seen] * rng.uniform(0.5, 2.0, (int(seen.sum()), 1))
        score = rng.uniform(0.35, 0.95, int(seen.sum()))
        n_fp = int(rng.integers(0, 4))
        fc = rng.uniform(0.1, 0.9, (n_fp, 2))
        fs = rng.uniform(0.03, 0.1, (n_fp, 2))
        boxes = np.concatenate([boxes, np.concatenate([fc - fs / 2, fc + fs / 2], axis=1)])
        emb = np.concatenate([emb, rng.standard_normal((n_fp, e))])
        score = np.concatenate([score, rng.uniform(0.05, 0.6, n_fp)])
        order = np.argsort(-score, kind="stable")                  # gather_tracking2d returns detections sorted by score
        frames.append((boxes[order].astype(np.float32), np.zeros(len(order), np.int64), score[order].astype(np.float32),
                       emb[order].astype(np.float32)))
    return frames


def run_track_sequence(tracker, frames):
    """Feeds the frames to ``tracker.update`` and records, per frame, every live track: (frame, track_id, state value,
    bbox) rows - the observable behaviour of the association."""
    import numpy as np
    rows = []
    for f, (b, l, s, e) in enumerate(frames):
        tracker.update(b, l, s, e)
        for t in tracker.tracks:
            rows.append([f, t.track_id, t.state.value, *np.asarray(t.bbox, dtype=np.float64).tolist()])
    return np.array(rows, dtype=np.float64)
