"""COCO-style detection metrics without pycocotools - the consumer of ``validation_step``'s predictions.

Mirrors the reference's ``centernet_lightning/eval/coco.py``:

* ``gather_and_merge(data)``            <- eval/coco.py:10-18 (the reference's only collective: ``dist.all_gather_object`` of the
                                           per-image lists, flattened in rank order)
* ``CocoEvaluator(num_classes)``        <- eval/coco.py:21-109, same ``update(preds, targets)`` / ``get_metrics()`` / ``reset()``
                                           contract, same 12 metric names in the same order

The reference delegates the arithmetic to ``pycocotools.cocoeval.COCOeval`` (``evaluate`` / ``accumulate`` / ``summarize``
with its default bbox parameters), a third-party dependency that is absent from /root/reference and not installed here
(requirements.txt: ``pycocotools``, unpinned).  Its published algorithm (cocoeval.py of cocoapi 2.0) is restated below:

* parameters: IoU thresholds 0.50:0.05:0.95, recall thresholds 0:0.01:1, maxDets (1, 10, 100), area ranges all / small
  (< 32^2) / medium / large (>= 96^2), categories evaluated separately (useCats);
* per image, category and area range: detections sorted by score (stable, descending) and cut to 100; ground truths with
  an area outside the range are "ignored" and sorted behind the others; greedy matching in score order - a detection takes
  the not-yet-matched ground truth of highest IoU >= threshold, preferring non-ignored ones; unmatched detections whose own
  area is outside the range are ignored;
* accumulate: per category / area / maxDets all detections of all images are merged by score (stable), cumulative TP / FP
  give recall and precision, precision is made monotonically non-increasing from the right and sampled at the 101 recall
  thresholds (``np.searchsorted(..., side="left")``); categories without ground truth stay at -1 and are left out of the means;
* summarize: the 12 numbers of ``COCOeval.stats``.

Boxes are xywh (``validation_step`` converts with cnl_boxes_xyxy_to_xywh first, reference models/centernet.py:207); every
annotation has ``iscrowd = 0`` and ``area = w * h`` exactly as ``CocoEvaluator.create_coco`` builds them (eval/coco.py:80-95).

PARITY STATUS: "parity unpinned" against pycocotools itself (not installable here); pinned by hand-computed known-answer
cases in tests/test_evaluate.py.  This is host-side bookkeeping (numpy, float64 like pycocotools), not part of the GPU path.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

METRIC_NAMES = ("mAP", "AP50", "AP75", "AP_small", "AP_medium", "AP_large",
                "AR1", "AR10", "mAR", "AR_small", "AR_medium", "AR_large")          # reference eval/coco.py:24-27

IOU_THRS = np.linspace(0.5, 0.95, int(np.round((0.95 - 0.5) / 0.05)) + 1, endpoint=True)
REC_THRS = np.linspace(0.0, 1.00, int(np.round((1.00 - 0.0) / 0.01)) + 1, endpoint=True)
MAX_DETS = (1, 10, 100)
AREA_RNG = ((0.0, 1e5 ** 2), (0.0, 32.0 ** 2), (32.0 ** 2, 96.0 ** 2), (96.0 ** 2, 1e5 ** 2))   # all, small, medium, large


def gather_and_merge(data: list) -> list:
    """reference eval/coco.py:10-18: every rank receives the concatenation (rank order) of all ranks' lists."""
    import torch.distributed as dist
    world_size = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    if world_size == 1:
        return data
    data_list = [None] * world_size
    dist.all_gather_object(data_list, data)
    return [x for part in data_list for x in part]


def box_iou_xywh(dt: np.ndarray, gt: np.ndarray) -> np.ndarray:
    """(D,4) x (G,4) xywh -> (D,G) IoU in float64 (pycocotools maskUtils.iou for boxes, iscrowd = 0)."""
    dt = np.asarray(dt, dtype=np.float64).reshape(-1, 4)
    gt = np.asarray(gt, dtype=np.float64).reshape(-1, 4)
    if len(dt) == 0 or len(gt) == 0:
        return np.zeros((len(dt), len(gt)), dtype=np.float64)
    dx1, dy1, dx2, dy2 = dt[:, 0], dt[:, 1], dt[:, 0] + dt[:, 2], dt[:, 1] + dt[:, 3]
    gx1, gy1, gx2, gy2 = gt[:, 0], gt[:, 1], gt[:, 0] + gt[:, 2], gt[:, 1] + gt[:, 3]
    iw = np.minimum(dx2[:, None], gx2[None, :]) - np.maximum(dx1[:, None], gx1[None, :])
    ih = np.minimum(dy2[:, None], gy2[None, :]) - np.maximum(dy1[:, None], gy1[None, :])
    inter = np.where((iw > 0) & (ih > 0), iw * ih, 0.0)
    union = (dt[:, 2] * dt[:, 3])[:, None] + (gt[:, 2] * gt[:, 3])[None, :] - inter
    return np.where(inter > 0, inter / np.where(union > 0, union, 1.0), 0.0)


def _evaluate_img(d_boxes, d_scores, g_boxes, area_rng):
    """COCOeval.evaluateImg for one (image, category, area range) with maxDet = 100.  Detections arrive sorted by score.
    Returns (dt_matched (T,D) bool, dt_ignore (T,D) bool, dt_scores (D,), gt_ignore (G,) bool) or None."""
    n_d, n_g = len(d_boxes), len(g_boxes)
    if n_d == 0 and n_g == 0:
        return None
    g_area = g_boxes[:, 2] * g_boxes[:, 3] if n_g else np.zeros((0,))
    g_ig = (g_area < area_rng[0]) | (g_area > area_rng[1])
    g_order = np.argsort(g_ig, kind="mergesort")                      # non-ignored ground truths first
    g_boxes, g_ig = g_boxes[g_order], g_ig[g_order]
    ious = box_iou_xywh(d_boxes, g_boxes)
    n_t = len(IOU_THRS)
    gtm = np.zeros((n_t, n_g), dtype=bool)
    dtm = np.zeros((n_t, n_d), dtype=bool)
    dt_ig = np.zeros((n_t, n_d), dtype=bool)
    if n_g and n_d:
        for ti, t in enumerate(IOU_THRS):
            for di in range(n_d):
                iou = min(t, 1 - 1e-10)
                m = -1
                for gi in range(n_g):
                    if gtm[ti, gi]:                                   # already matched (no crowd annotations here)
                        continue
                    if m > -1 and not g_ig[m] and g_ig[gi]:           # a regular match exists: stop at the ignored ones
                        break
                    if ious[di, gi] < iou:
                        continue
                    iou = ious[di, gi]
                    m = gi
                if m == -1:
                    continue
                dt_ig[ti, di] = g_ig[m]
                dtm[ti, di] = True
                gtm[ti, m] = True
    d_area = d_boxes[:, 2] * d_boxes[:, 3] if n_d else np.zeros((0,))
    out_of_range = (d_area < area_rng[0]) | (d_area > area_rng[1])
    dt_ig = dt_ig | (~dtm & out_of_range[None, :])
    return dtm, dt_ig, d_scores, g_ig


def coco_bbox_stats(preds: Sequence[Dict[str, np.ndarray]], targets: Sequence[Dict[str, np.ndarray]], num_classes: int) -> np.ndarray:
    """The 12 numbers of COCOeval.stats for per-image prediction dicts {boxes (k,4) xywh, scores (k,), labels (k,)} and
    target dicts {boxes (g,4) xywh, labels (g,)} (image i of preds <-> image i of targets, reference eval/coco.py:57-60)."""
    assert len(preds) == len(targets)
    n_t, n_r, n_a, n_m = len(IOU_THRS), len(REC_THRS), len(AREA_RNG), len(MAX_DETS)
    precision = -np.ones((n_t, n_r, num_classes, n_a, n_m))
    recall = -np.ones((n_t, num_classes, n_a, n_m))
    per_img = []
    for p, t in zip(preds, targets):
        pb = np.asarray(p["boxes"], dtype=np.float64).reshape(-1, 4)
        ps = np.asarray(p["scores"], dtype=np.float64).reshape(-1)
        pl = np.asarray(p["labels"]).reshape(-1).astype(np.int64)
        tb = np.asarray(t["boxes"], dtype=np.float64).reshape(-1, 4)
        tl = np.asarray(t["labels"]).reshape(-1).astype(np.int64)
        per_img.append((pb, ps, pl, tb, tl))
    eps = np.spacing(1)
    for c in range(num_classes):
        # detections of this category per image, sorted by score (stable) and cut to maxDets[-1] (COCOeval.computeIoU)
        img_dt = []
        for pb, ps, pl, tb, tl in per_img:
            sel = np.nonzero(pl == c)[0]
            order = sel[np.argsort(-ps[sel], kind="mergesort")][:MAX_DETS[-1]]
            img_dt.append((pb[order], ps[order], tb[tl == c]))
        for ai, rng in enumerate(AREA_RNG):
            evals = [e for e in (_evaluate_img(db, ds, gb, rng) for db, ds, gb in img_dt) if e is not None]
            if not evals:
                continue
            for mi, max_det in enumerate(MAX_DETS):
                scores = np.concatenate([e[2][:max_det] for e in evals])
                order = np.argsort(-scores, kind="mergesort")
                dtm = np.concatenate([e[0][:, :max_det] for e in evals], axis=1)[:, order]
                dt_ig = np.concatenate([e[1][:, :max_det] for e in evals], axis=1)[:, order]
                gt_ig = np.concatenate([e[3] for e in evals])
                npig = int(np.count_nonzero(~gt_ig))
                if npig == 0:
                    continue
                tp_sum = np.cumsum(dtm & ~dt_ig, axis=1).astype(np.float64)
                fp_sum = np.cumsum(~dtm & ~dt_ig, axis=1).astype(np.float64)
                for ti in range(n_t):
                    tp, fp = tp_sum[ti], fp_sum[ti]
                    nd = len(tp)
                    rc = tp / npig
                    pr = tp / (fp + tp + eps)
                    recall[ti, c, ai, mi] = rc[-1] if nd else 0
                    pr = pr.tolist()
                    for i in range(nd - 1, 0, -1):
                        if pr[i] > pr[i - 1]:
                            pr[i - 1] = pr[i]
                    inds = np.searchsorted(rc, REC_THRS, side="left")
                    q = np.zeros((n_r,))
                    for ri, pi in enumerate(inds):
                        if pi < nd:
                            q[ri] = pr[pi]
                    precision[ti, :, c, ai, mi] = q

    def summarize(ap: bool, iou_thr=None, area: int = 0, max_det: int = 100) -> float:
        mi = MAX_DETS.index(max_det)
        s = precision[:, :, :, area, mi] if ap else recall[:, :, area, mi]
        if iou_thr is not None:
            s = s[np.where(np.isclose(IOU_THRS, iou_thr))[0]]
        s = s[s > -1]
        return float(np.mean(s)) if s.size else -1.0

    return np.array([
        summarize(True), summarize(True, 0.5), summarize(True, 0.75), summarize(True, area=1), summarize(True, area=2), summarize(True, area=3),
        summarize(False, max_det=1), summarize(False, max_det=10), summarize(False), summarize(False, area=1), summarize(False, area=2),
        summarize(False, area=3)])


class CocoEvaluator:
    """Drop-in for the reference's CocoEvaluator (eval/coco.py:21-109): accumulate per-image numpy dicts, then compute the 12
    COCO bbox metrics over the predictions and targets of ALL ranks."""
    pred_keys = ("boxes", "scores", "labels")
    target_keys = ("boxes", "labels")
    metric_names = METRIC_NAMES

    def __init__(self, num_classes: int):
        self.num_classes = num_classes
        self.reset()

    def update(self, preds: List[Dict[str, np.ndarray]], targets: List[Dict[str, np.ndarray]]) -> None:
        """preds: one dict per image with boxes (xywh), scores, labels; targets: boxes (xywh), labels (eval/coco.py:48-56)."""
        assert len(preds) == len(targets)
        self.preds.extend(preds)
        self.targets.extend(targets)

    def reset(self) -> None:
        self.preds = []
        self.targets = []

    def get_metrics(self) -> Dict[str, float]:
        preds = gather_and_merge(self.preds)               # eval/coco.py:63-64
        targets = gather_and_merge(self.targets)
        stats = coco_bbox_stats(preds, targets, self.num_classes)
        return {name: float(stats[i]) for i, name in enumerate(self.metric_names)}
