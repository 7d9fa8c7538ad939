"""Tracker oracle (TEST INFRASTRUCTURE ONLY).

* ``cosine_distance_matrix`` / ``box_iou_distance_matrix`` / ``box_giou_distance_matrix``: the host computations the
  reference's tracker performs per frame - ``scipy.spatial.distance.cdist(A, B, "cosine")``
  (centernet_lightning/models/tracker.py:61,157) and the numpy box formulas of centernet_lightning/utils/box.py:49-92 -
  restated in float64.  They check csrc/cnl_track.cu.
* ``import_reference_tracker()``: the reference's UNMODIFIED ``centernet_lightning/models/tracker.py`` imported from
  where it lies.  The snapshot's ``centernet_lightning/utils/__init__.py`` has its re-exports commented out (mid-refactor)
  and ``filterpy`` is not installed, so the module is loaded under two stand-ins: a ``centernet_lightning.utils`` package
  re-exporting the reference's own ``utils/box.py`` functions, and a ``filterpy.kalman`` whose ``KalmanFilter`` is a
  placeholder (Kalman parity is therefore UNPINNED; goldens use ``use_kalman=False``).  Used by
  tests/golden/gen_golden.py to produce tests/golden/tracker_*.npz and by tests/test_oracle.py when /root/reference exists.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

from .ref_import import REFERENCE_ROOT, reference_available


def cosine_distance_matrix(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """scipy cdist 'cosine': 1 - u.v / (|u| |v|), cosine clipped to [-1, 1]; float64, sequential dot products."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    out = np.empty((len(a), len(b)), np.float64)
    na = [np.sqrt(_dot(u, u)) for u in a]
    nb = [np.sqrt(_dot(v, v)) for v in b]
    for i, u in enumerate(a):
        for j, v in enumerate(b):
            c = _dot(u, v) / (na[i] * nb[j])
            if abs(c) > 1.0:
                c = np.copysign(1.0, c)
            out[i, j] = 1.0 - c
    return out


def _dot(u, v) -> float:
    s = 0.0
    for x, y in zip(u.tolist(), v.tolist()):
        s += x * y
    return s


def _inter_union(b1: np.ndarray, b2: np.ndarray):
    b1 = np.asarray(b1, dtype=np.float64)
    b2 = np.asarray(b2, dtype=np.float64)
    area1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    area2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    lt = np.maximum(b1[:, None, :2], b2[None, :, :2])
    rb = np.minimum(b1[:, None, 2:], b2[None, :, 2:])
    wh = np.clip(rb - lt, 0, None)
    inter = wh[..., 0] * wh[..., 1]
    return inter, area1[:, None] + area2[None, :] - inter


def box_iou_distance_matrix(b1: np.ndarray, b2: np.ndarray) -> np.ndarray:
    inter, union = _inter_union(b1, b2)                       # utils/box.py:49-67, 83-86
    return 1 - inter / union


def box_giou_distance_matrix(b1: np.ndarray, b2: np.ndarray) -> np.ndarray:
    inter, union = _inter_union(b1, b2)                       # utils/box.py:70-80, 89-92
    b1 = np.asarray(b1, dtype=np.float64)
    b2 = np.asarray(b2, dtype=np.float64)
    lt = np.minimum(b1[:, None, :2], b2[None, :, :2])
    rb = np.maximum(b1[:, None, 2:], b2[None, :, 2:])
    wh = np.clip(rb - lt, 0, None)
    hull = wh[..., 0] * wh[..., 1]
    return 1 - (inter / union - (hull - union) / hull)


def import_reference_tracker():
    """The reference's own tracker module (see the module docstring for the two stand-ins it is loaded under)."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    name = "_cnl_ref_tracker_pkg"
    if name + ".models.tracker" in sys.modules:
        return sys.modules[name + ".models.tracker"]

    def load(mod_name, path):
        spec = importlib.util.spec_from_file_location(mod_name, path)
        m = importlib.util.module_from_spec(spec)
        sys.modules[mod_name] = m
        spec.loader.exec_module(m)
        return m

    root = os.path.join(REFERENCE_ROOT, "centernet_lightning")
    pkg = types.ModuleType(name)
    pkg.__path__ = []
    sys.modules[name] = pkg
    box = load(name + ".utils.box", os.path.join(root, "utils", "box.py"))
    utils = types.ModuleType(name + ".utils")
    utils.__path__ = []
    for k in ("box_iou_distance_matrix", "box_giou_distance_matrix", "box_iou_matrix", "box_giou_matrix"):
        setattr(utils, k, getattr(box, k))
    utils.load_config = lambda *a, **k: {}
    sys.modules[name + ".utils"] = utils
    models = types.ModuleType(name + ".models")
    models.__path__ = []
    sys.modules[name + ".models"] = models
    if "filterpy" not in sys.modules:
        fp, fk = types.ModuleType("filterpy"), types.ModuleType("filterpy.kalman")
        fk.KalmanFilter = object
        fp.kalman = fk
        sys.modules["filterpy"], sys.modules["filterpy.kalman"] = fp, fk
    return load(name + ".models.tracker", os.path.join(root, "models", "tracker.py"))
