"""CPU: the pycocotools-free COCO evaluator (centernet-lightning_b200/evaluate.py <- reference eval/coco.py:10-109) against
hand-computed known answers, and its cross-rank merge with world_size-2 gloo processes."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from centernet_lightning_b200 import evaluate as ev


def _p(boxes, scores, labels):
    return {"boxes": np.array(boxes, dtype=np.float32).reshape(-1, 4), "scores": np.array(scores, dtype=np.float32),
            "labels": np.array(labels, dtype=np.int64)}


def _t(boxes, labels):
    return {"boxes": np.array(boxes, dtype=np.float32).reshape(-1, 4), "labels": np.array(labels, dtype=np.int64)}


def test_metric_names_and_parameters_are_cocoeval_defaults():
    assert ev.CocoEvaluator.metric_names == ("mAP", "AP50", "AP75", "AP_small", "AP_medium", "AP_large",
                                             "AR1", "AR10", "mAR", "AR_small", "AR_medium", "AR_large")
    np.testing.assert_allclose(ev.IOU_THRS, [0.5, 0.55, 0.6, 0.65, 0.7, 0.75, 0.8, 0.85, 0.9, 0.95])
    assert len(ev.REC_THRS) == 101 and ev.REC_THRS[50] == 0.5 and ev.MAX_DETS == (1, 10, 100)


def test_box_iou_xywh():
    iou = ev.box_iou_xywh(np.array([[0, 0, 10, 10], [20, 20, 5, 5]]), np.array([[5, 0, 10, 10], [0, 0, 10, 10]]))
    np.testing.assert_allclose(iou, [[50 / 150, 1.0], [0.0, 0.0]])
    assert ev.box_iou_xywh(np.zeros((0, 4)), np.zeros((3, 4))).shape == (0, 3)


def test_perfect_predictions_score_one():
    """Predictions equal to the targets: every defined metric is 1 (one small, one medium, one large object; an empty image)."""
    targets = [_t([[5, 5, 10, 10]], [0]), _t([[0, 0, 50, 50], [60, 60, 100, 100]], [0, 1]), _t([], [])]
    preds = [_p(t["boxes"], np.linspace(0.9, 0.5, len(t["labels"])), t["labels"]) for t in targets]
    e = ev.CocoEvaluator(3)                       # class 2 never appears: left out of every mean
    e.update(preds, targets)
    m = e.get_metrics()
    assert list(m) == list(ev.METRIC_NAMES)
    for k, v in m.items():
        assert v == pytest.approx(1.0, abs=1e-9), (k, v)
    e.reset()
    assert e.preds == [] and e.targets == []


def test_hand_computed_single_image():
    """One large ground truth [0,0,100,100]; detections d1 = [0,0,100,82] (IoU 0.82, score 0.9) and d2 = the box itself
    (IoU 1, score 0.5).  IoU thresholds <= 0.8 (7 of 10): d1 is the true positive, d2 a later false positive -> AP 1.
    Thresholds >= 0.85: d1 is a false positive ranked first, d2 the true positive -> precision 0.5 at every recall -> AP 0.5.
    mAP = (7 + 3 * 0.5) / 10.  With maxDets = 1 only d1 counts: recall 1 at 7 thresholds, 0 at 3.  In the "large" range d1
    (area 8200 < 96^2) is ignored when unmatched, so the large-object AP is 1 at every threshold."""
    e = ev.CocoEvaluator(1)
    e.update([_p([[0, 0, 100, 82], [0, 0, 100, 100]], [0.9, 0.5], [0, 0])], [_t([[0, 0, 100, 100]], [0])])
    m = e.get_metrics()
    expect = dict(mAP=0.85, AP50=1.0, AP75=1.0, AP_small=-1.0, AP_medium=-1.0, AP_large=1.0,
                  AR1=0.7, AR10=1.0, mAR=1.0, AR_small=-1.0, AR_medium=-1.0, AR_large=1.0)
    for k, v in expect.items():
        assert m[k] == pytest.approx(v, abs=1e-9), (k, m[k], v)


def test_hand_computed_missed_object_and_wrong_class():
    """Class 0: two objects over two images, one found (score 0.9), one missed -> recall 0.5, precision 1 up to recall 0.5:
    AP = 51/101 at every threshold (51 of the 101 recall thresholds are <= 0.5).  Class 1: one object, one perfect
    detection -> AP 1.  The reported numbers are the means over the two classes; all objects are medium-sized."""
    targets = [_t([[10, 10, 40, 40], [100, 100, 40, 40]], [0, 1]), _t([[30, 30, 40, 40]], [0])]
    preds = [_p([[10, 10, 40, 40], [100, 100, 40, 40]], [0.9, 0.8], [0, 1]), _p([], [], [])]
    m = ev.CocoEvaluator(2)
    m.update(preds, targets)
    out = m.get_metrics()
    ap0 = 51 / 101
    assert out["mAP"] == pytest.approx((ap0 + 1.0) / 2, abs=1e-9)
    assert out["AP50"] == pytest.approx((ap0 + 1.0) / 2, abs=1e-9)
    assert out["mAR"] == pytest.approx((0.5 + 1.0) / 2, abs=1e-9) and out["AR1"] == pytest.approx(0.75, abs=1e-9)
    assert out["AP_medium"] == pytest.approx((ap0 + 1.0) / 2, abs=1e-9) and out["AP_small"] == -1.0 and out["AP_large"] == -1.0


def test_score_order_and_max_dets():
    """Detections are ranked by score across images; a high-scoring false positive in another image lowers precision at
    low recall only after the monotone envelope: sequence (FP 0.95, TP 0.9, TP 0.8) over 2 objects -> precision envelope
    2/3 everywhere -> AP 2/3.  More than 100 detections per image and category are cut to the 100 best."""
    targets = [_t([[0, 0, 50, 50]], [0]), _t([[0, 0, 60, 60]], [0])]
    preds = [_p([[0, 0, 50, 50], [200, 200, 50, 50]], [0.9, 0.95], [0, 0]), _p([[0, 0, 60, 60]], [0.8], [0])]
    e = ev.CocoEvaluator(1)
    e.update(preds, targets)
    assert e.get_metrics()["mAP"] == pytest.approx(2 / 3, abs=1e-9)
    many = _p(np.tile([[300, 300, 50, 50]], (150, 1)), np.linspace(0.99, 0.5, 150), np.zeros(150))
    many["boxes"] = np.concatenate([many["boxes"], [[0, 0, 50, 50]]]).astype(np.float32)     # the true positive has the LOWEST score
    many["scores"] = np.concatenate([many["scores"], [0.1]]).astype(np.float32)
    many["labels"] = np.zeros(151, dtype=np.int64)
    e = ev.CocoEvaluator(1)
    e.update([many], [targets[0]])
    assert e.get_metrics()["mAR"] == 0.0                                      # cut off by maxDets = 100


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    e = ev.CocoEvaluator(1)
    if rank == 0:      # the found object lives on rank 0, the missed one on rank 1: only the MERGED lists give 51/101
        e.update([_p([[10, 10, 40, 40]], [0.9], [0])], [_t([[10, 10, 40, 40]], [0])])
    else:
        e.update([_p([], [], [])], [_t([[30, 30, 40, 40]], [0])])
    merged = ev.gather_and_merge([rank])
    q.put((rank, e.get_metrics()["mAP"], merged))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_and_merge_world2_gloo():
    assert ev.gather_and_merge([1, 2]) == [1, 2]                               # no process group: identity (eval/coco.py:12-13)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ap, merged in res:
        assert ap == pytest.approx(51 / 101, abs=1e-9) and merged == [0, 1]
