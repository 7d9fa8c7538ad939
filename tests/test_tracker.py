"""Tracker.  CPU: the oracle cost functions against scipy / the reference's own box utilities, and the
host-side association logic (driven with the ORACLE's cost functions as the callables the reference API allows) against
goldens recorded from the reference's own Tracker.  GPU: the CUDA cost kernel against the oracle, the default
(CUDA-cost) tracker against the same goldens, and step_batch end to end."""
import os
import warnings

import numpy as np
import pytest
import torch

import cases
from oracle import ref_import, tracker_np

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _gold(name):
    return dict(np.load(os.path.join(GOLD, f"tracker_{name}.npz")))


def _rand_boxes(rng, n):
    c = rng.uniform(0.1, 0.9, (n, 2))
    s = rng.uniform(0.02, 0.4, (n, 2))
    return np.concatenate([c - s / 2, c + s / 2], axis=1)


def test_oracle_costs_match_scipy_and_reference_box_utils():
    from scipy.spatial import distance
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal((7, 64)), rng.standard_normal((5, 64))
    np.testing.assert_allclose(tracker_np.cosine_distance_matrix(a, b), distance.cdist(a, b, "cosine"), rtol=0, atol=1e-14)
    b1, b2 = _rand_boxes(rng, 9), _rand_boxes(rng, 4)
    d = tracker_np.box_iou_distance_matrix(b1, b2)
    assert d.shape == (9, 4) and (d >= 0).all() and (d <= 1).all()
    np.testing.assert_allclose(tracker_np.box_iou_distance_matrix(b1, b1).diagonal(), 0, atol=1e-12)
    g = tracker_np.box_giou_distance_matrix(b1, b2)
    assert (g >= d - 1e-12).all() and (g <= 2).all()
    if ref_import.reference_available():
        ref = tracker_np.import_reference_tracker()
        box = ref.box_iou_distance_matrix.__globals__
        assert np.array_equal(box["box_iou_distance_matrix"](b1, b2), d)
        assert np.array_equal(box["box_giou_distance_matrix"](b1, b2), g)


@pytest.mark.parametrize("name", list(cases.TRACK_CASES))
def test_association_logic_reproduces_reference_goldens(name):
    """Host logic only (no CUDA): reid_cost / box_cost given as callables, as the reference allows (tracker.py:61-64)."""
    from centernet_lightning_b200.tracker import Tracker
    case = cases.TRACK_CASES[name]
    kw = dict(case["tracker"])
    box = kw.pop("box_cost", "iou")
    box_fn = {None: None, "iou": tracker_np.box_iou_distance_matrix, "giou": tracker_np.box_giou_distance_matrix}[box]
    t = Tracker(model=None, reid_cost=tracker_np.cosine_distance_matrix, box_cost=box_fn, **kw)
    rows = cases.run_track_sequence(t, cases.make_track_sequence(case))
    gold = _gold(name)
    assert rows.shape == gold["rows"].shape
    assert np.array_equal(rows[:, :3], gold["rows"][:, :3])              # frame, track id, state: exact
    np.testing.assert_allclose(rows[:, 3:], gold["rows"][:, 3:], rtol=0, atol=0)
    assert t.next_track_id == int(gold["next_track_id"])


@pytest.mark.skipif(not ref_import.reference_available(), reason="needs /root/reference")
def test_goldens_are_what_the_reference_tracker_produces():
    ref = tracker_np.import_reference_tracker()
    case = cases.TRACK_CASES["walk_iou"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t = ref.Tracker(model=None, **case["tracker"])
    rows = cases.run_track_sequence(t, cases.make_track_sequence(case))
    assert np.array_equal(rows, _gold("walk_iou")["rows"])


def test_kalman_filter_and_track_life_cycle():
    from centernet_lightning_b200.tracker import KalmanFilter, Track, TrackState, Tracker, match_with_threshold
    kf = KalmanFilter(2, 1)
    kf.F = np.array([[1.0, 1.0], [0.0, 1.0]])
    kf.H = np.array([[1.0, 0.0]])
    kf.P = np.eye(2)
    for z in (1.0, 2.0, 3.0, 4.0):                                      # unit-velocity target, exact measurements
        kf.predict(Q=np.zeros((2, 2)))
        kf.update(np.array([z]), R=np.array([[1e-6]]))
    assert abs(kf.x[0] - 4.0) < 1e-3 and abs(kf.x[1] - 1.0) < 1e-2
    assert np.allclose(kf.P, kf.P.T) and np.all(np.linalg.eigvalsh(kf.P) > -1e-12)
    tr = Track(0, np.array([0.1, 0.1, 0.3, 0.3]), 0, np.ones(4), min_birth_age=2, max_inactive_age=1, use_kalman=True)
    assert abs(np.linalg.norm(tr.embedding) - 1) < 1e-12 and tr.state == TrackState.UNCONFIRMED
    tr.update_matched(np.array([0.1, 0.1, 0.3, 0.3]), np.ones(4)); assert not tr.active
    tr.update_matched(np.array([0.11, 0.1, 0.31, 0.3]), np.ones(4)); assert tr.active
    tr.kalman_predict()
    tr.update_unmatched(); assert tr.state == TrackState.INACTIVE
    tr.update_unmatched(); assert tr.to_delete
    m, ur, uc = match_with_threshold(np.array([[0.1, 0.9], [0.8, 0.6], [0.3, 0.2]]), 0.5)
    assert m == [(0, 0), (2, 1)] and ur == [1] and uc == []
    assert match_with_threshold(np.zeros((0, 3)), 0.5) == ([], [], [0, 1, 2])
    with pytest.raises(NotImplementedError):
        Tracker(reid_cost="euclidean")
    if not torch.cuda.is_available():                                   # string costs are CUDA-only: no silent host fallback
        t = Tracker(model=None)
        b = np.array([[0.1, 0.1, 0.2, 0.2]], np.float32)
        t.update(b, np.zeros(1, np.int64), np.ones(1, np.float32), np.ones((1, 8), np.float32))
        with pytest.raises(Exception):
            t.update(b, np.zeros(1, np.int64), np.ones(1, np.float32), np.ones((1, 8), np.float32))


# ---------------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_cuda_cost_matrices_equal_oracle(cuda):
    from centernet_lightning_b200.tracker import CostMatrices
    eng = CostMatrices(cuda)
    rng = np.random.default_rng(5)
    for na, nb, e in [(1, 1, 64), (300, 100, 64), (37, 211, 16), (5, 3, 1)]:
        a = rng.standard_normal((na, e)).astype(np.float32)              # detections arrive as float32
        b = rng.standard_normal((nb, e))
        b /= np.linalg.norm(b, axis=1, keepdims=True)
        b1, b2 = _rand_boxes(rng, na).astype(np.float32), _rand_boxes(rng, nb)
        for giou in (False, True):
            reid, box = eng(a, b, b1, b2, giou=giou)
            assert reid.shape == box.shape == (na, nb) and reid.dtype == np.float64
            np.testing.assert_allclose(reid, tracker_np.cosine_distance_matrix(a, b), rtol=0, atol=1e-14)
            ref_box = (tracker_np.box_giou_distance_matrix if giou else tracker_np.box_iou_distance_matrix)(b1, b2)
            assert np.array_equal(box, ref_box)
    a = rng.standard_normal((4, 8))
    reid, box = eng(a, 3.0 * a, None, None)                              # parallel vectors: clipped cosine, distance 0 on the diagonal
    assert box is None and np.abs(reid.diagonal()).max() < 1e-15
    reid, box = eng(None, None, _rand_boxes(rng, 3), _rand_boxes(rng, 2))
    assert reid is None and box.shape == (3, 2)
    assert eng(np.zeros((0, 8)), np.zeros((2, 8)), None, None)[0].shape == (0, 2)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(cases.TRACK_CASES))
def test_cuda_tracker_reproduces_reference_goldens(cuda, name):
    from centernet_lightning_b200.tracker import Tracker
    case = cases.TRACK_CASES[name]
    t = Tracker(model=None, device=cuda, **case["tracker"])
    rows = cases.run_track_sequence(t, cases.make_track_sequence(case))
    gold = _gold(name)
    assert rows.shape == gold["rows"].shape
    assert np.array_equal(rows[:, :3], gold["rows"][:, :3])
    assert np.array_equal(rows[:, 3:], gold["rows"][:, 3:])


@pytest.mark.gpu
def test_step_batch_end_to_end(cuda):
    """reference tracker.py:83-129 on the sm_100a model: forward -> gather_tracking2d -> update, frame by frame."""
    from centernet_lightning_b200.model import CenterNet
    from centernet_lightning_b200.tracker import Tracker, build_tracker
    net = CenterNet(1, reid_dim=64, box_multiplier=16.0).init_synthetic_(3).to(cuda)
    trk = build_tracker(dict(num_detections=50, detection_threshold=0.0, min_birth_age=1), model=net)
    g = torch.Generator().manual_seed(0)
    base = torch.rand((1, 3, 128, 128), generator=g)
    imgs = torch.cat([base + 0.01 * i for i in range(4)])                # nearly identical frames: tracks must persist
    out = trk.step_batch(imgs)
    assert len(out["bboxes"]) == len(out["track_ids"]) == 4 and trk.frame == 4
    assert len(out["track_ids"][0]) == 0                                 # nothing is confirmed on the first frame
    assert len(out["track_ids"][-1]) > 10
    assert set(out["track_ids"][-1]) & set(out["track_ids"][-2])         # identities carried across frames
    # the same detections pushed through update() by hand give the same tracks
    trk2 = Tracker(model=None, device=cuda, num_detections=50, detection_threshold=0.0, min_birth_age=1)
    heat, box, reid = net(imgs.to(cuda))
    det = {k: v.cpu().numpy() for k, v in net.gather_tracking2d(heat, box, reid, num_detections=50, normalize_bbox=True).items()}
    for i in range(4):
        trk2.update(det["bboxes"][i], det["labels"][i], det["scores"][i], det["embeddings"][i])
    assert [t.track_id for t in trk2.tracks if t.active] == out["track_ids"][-1]
    single = trk.step_single(imgs[0])
    assert set(single) == {"bboxes", "track_ids"}
