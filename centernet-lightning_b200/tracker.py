"""Multi-object tracker on top of the tracking head - the consumer of ``gather_tracking2d``.

Same surface as the reference's ``Tracker`` (centernet_lightning/models/tracker.py:44-201): constructor parameters and
defaults (:51), ``step_batch(images)`` (:83-120), ``step_single(img)`` (:122-129), ``update(bboxes, labels, scores,
embeddings)`` (:131-201), ``reset()`` (:77-80), ``.tracks`` of objects with ``track_id / bbox / label / embedding /
active``; ``build_tracker(config, model)`` (:353-357).

What runs where:
  * forward + decode + embedding gather: the sm_100a path (``model(images)`` and ``model.gather_tracking2d``);
  * the two association cost matrices of a frame - cosine distance of the re-id embeddings and 1-IoU / 1-GIoU of the
    boxes - are computed by ONE CUDA launch for all (detection, track) pairs (csrc/cnl_track.cu, fp64, equal to the
    scipy / numpy results of the reference); there is no host implementation of them here;
  * the Hungarian assignment (scipy.optimize.linear_sum_assignment, reference :28) and the per-track bookkeeping stay on
    the host, as in the reference.
``reid_cost`` / ``box_cost`` may also be callables ``f(A, B) -> cost matrix`` exactly as the reference allows (:61-64)."""
from __future__ import annotations

import ctypes as C
from enum import Enum, auto
from typing import Callable, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib


class TrackState(Enum):
    UNCONFIRMED = auto()
    ACTIVE = auto()
    INACTIVE = auto()
    TO_DELETE = auto()


# ----------------------------------------------------------------------------------------------------------------
# association costs on the GPU
# ----------------------------------------------------------------------------------------------------------------
class CostMatrices:
    """Device staging + launch of cnl_track_cost_matrices for one device (buffers grow, never shrink)."""

    def __init__(self, device: Union[str, torch.device] = "cuda:0"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("tracker association costs run on CUDA (sm_100a) only - there is no CPU fallback")
        self.lib = _lib.load()
        self._cap = 0
        self._buf: Optional[torch.Tensor] = None

    def _stage(self, n_doubles: int) -> torch.Tensor:
        if n_doubles > self._cap:
            self._cap = max(2 * self._cap, n_doubles, 1 << 14)
            self._buf = torch.empty(self._cap, dtype=torch.float64, device=self.device)
        return self._buf

    def __call__(self, det_emb: Optional[np.ndarray], trk_emb: Optional[np.ndarray], det_box: Optional[np.ndarray],
                 trk_box: Optional[np.ndarray], giou: bool = False) -> Tuple[Optional[np.ndarray], Optional[np.ndarray]]:
        """(reid_cost, box_cost), each (n_det, n_trk) float64 or None when its inputs are None."""
        want_reid, want_box = det_emb is not None, det_box is not None
        na = len(det_emb) if want_reid else len(det_box)
        nb = len(trk_emb) if want_reid else len(trk_box)
        if na == 0 or nb == 0:
            empty = np.zeros((na, nb), np.float64)
            return (empty if want_reid else None), (empty.copy() if want_box else None)
        parts = []
        dim = 0
        if want_reid:
            det_emb = np.ascontiguousarray(det_emb, dtype=np.float64)
            trk_emb = np.ascontiguousarray(trk_emb, dtype=np.float64)
            dim = det_emb.shape[1]
            if trk_emb.shape != (nb, dim):
                raise ValueError(f"embedding shapes {det_emb.shape} / {trk_emb.shape}")
            parts += [det_emb.reshape(-1), trk_emb.reshape(-1)]
        if want_box:
            det_box = np.ascontiguousarray(det_box, dtype=np.float64)
            trk_box = np.ascontiguousarray(trk_box, dtype=np.float64)
            if det_box.shape != (na, 4) or trk_box.shape != (nb, 4):
                raise ValueError(f"box shapes {det_box.shape} / {trk_box.shape}")
            parts += [det_box.reshape(-1), trk_box.reshape(-1)]
        host = torch.from_numpy(np.concatenate(parts))
        n_in = host.numel()
        n_out = na * nb * (int(want_reid) + int(want_box))
        ws_bytes = self.lib.cnl_track_workspace_bytes(na, nb)
        buf = self._stage(n_in + n_out + ws_bytes // 8 + 64)
        with torch.cuda.device(self.device):
            buf[:n_in].copy_(host, non_blocking=False)
            base, off = buf.data_ptr(), 0
            ptr = {}
            for name, size in (("de", na * dim), ("te", nb * dim)) if want_reid else ():
                ptr[name] = base + 8 * off
                off += size
            for name, size in (("db", na * 4), ("tb", nb * 4)) if want_box else ():
                ptr[name] = base + 8 * off
                off += size
            out_off = off
            reid_ptr = base + 8 * off if want_reid else None
            off += na * nb if want_reid else 0
            box_ptr = base + 8 * off if want_box else None
            off += na * nb if want_box else 0
            ws_off = (off + 31) // 32 * 32                       # 256-byte aligned workspace inside the staging buffer
            st = self.lib.cnl_track_cost_matrices(ptr.get("de"), ptr.get("te"), dim, ptr.get("db"), ptr.get("tb"), na, nb,
                                                  int(bool(giou)), reid_ptr, box_ptr, base + 8 * ws_off, ws_bytes,
                                                  torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(st, "cnl_track_cost_matrices")
            out = buf[out_off:out_off + n_out].cpu().numpy()
        reid = out[:na * nb].reshape(na, nb) if want_reid else None
        box = out[-na * nb:].reshape(na, nb) if want_box else None
        return reid, box


def match_with_threshold(cost_matrix: np.ndarray, threshold: float):
    """Hungarian assignment, keeping only pairs cheaper than ``threshold`` (reference tracker.py:27-43)."""
    from scipy.optimize import linear_sum_assignment
    n_rows, n_cols = cost_matrix.shape
    rows, cols = linear_sum_assignment(cost_matrix) if n_rows and n_cols else ((), ())
    pairs = [(int(r), int(c)) for r, c in zip(rows, cols) if cost_matrix[r, c] < threshold]
    used_r = {r for r, _ in pairs}
    used_c = {c for _, c in pairs}
    return pairs, [r for r in range(n_rows) if r not in used_r], [c for c in range(n_cols) if c not in used_c]


# ----------------------------------------------------------------------------------------------------------------
# constant-velocity Kalman filter (the subset of filterpy.kalman.KalmanFilter the reference drives, tracker.py:233-304)
# ----------------------------------------------------------------------------------------------------------------
class KalmanFilter:
    """x' = F x, P' = F P F^T + Q;  K = P H^T (H P H^T + R)^-1, x += K (z - H x), P = (I-KH) P (I-KH)^T + K R K^T."""

    def __init__(self, dim_x: int, dim_z: int):
        self.x = np.zeros(dim_x)
        self.P = np.eye(dim_x)
        self.F = np.eye(dim_x)
        self.H = np.zeros((dim_z, dim_x))
        self._I = np.eye(dim_x)

    def predict(self, Q: np.ndarray) -> None:
        self.x = self.F @ self.x
        self.P = self.F @ self.P @ self.F.T + Q

    def update(self, z: np.ndarray, R: np.ndarray) -> None:
        PHT = self.P @ self.H.T
        S = self.H @ PHT + R
        K = PHT @ np.linalg.inv(S)
        self.x = self.x + K @ (np.asarray(z, dtype=np.float64) - self.H @ self.x)
        IKH = self._I - K @ self.H
        self.P = IKH @ self.P @ IKH.T + K @ R @ K.T


class Track:
    """One tracked object (reference tracker.py:217-351): life cycle UNCONFIRMED -> ACTIVE <-> INACTIVE -> TO_DELETE,
    unit-norm appearance embedding with exponential smoothing, optional constant-velocity Kalman filter on the box."""

    def __init__(self, track_id: int, bbox: np.ndarray, label, embedding: np.ndarray, min_birth_age: int = 2,
                 max_inactive_age: int = 30, smoothing_factor: float = 0.9, use_kalman: bool = False):
        self.track_id = track_id
        self.state = TrackState.UNCONFIRMED
        self.birth_age = 0
        self.inactive_age = 0
        self.bbox = bbox
        self.label = label
        self.embedding = embedding / np.linalg.norm(embedding)
        self.min_birth_age = min_birth_age
        self.max_inactive_age = max_inactive_age
        self.smoothing_factor = smoothing_factor
        self.kf: Optional[KalmanFilter] = None
        if use_kalman:
            kf = KalmanFilter(dim_x=8, dim_z=4)
            kf.x[:4] = bbox                                      # state = box corners + their velocities
            kf.F[:4, 4:] = np.eye(4)                             # constant velocity
            kf.H = np.eye(4, 8)                                  # only the corners are observed
            wh = np.asarray(bbox[2:] - bbox[:2], dtype=np.float64)
            std = np.tile(wh, 4)
            std[:4] /= 10                                        # DeepSORT-style initial uncertainty (reference :247-252)
            std[4:] /= 16
            kf.P = np.diag(std ** 2)
            self.kf = kf

    active = property(lambda self: self.state == TrackState.ACTIVE)
    confirmed = property(lambda self: self.state != TrackState.UNCONFIRMED)
    to_delete = property(lambda self: self.state == TrackState.TO_DELETE)

    def kalman_predict(self) -> None:
        if self.kf is None:
            return
        wh = self.kf.x[2:4] - self.kf.x[:2]
        std = np.tile(wh, 4)
        std[:4] /= 20                                            # process noise (reference :275-281)
        std[4:] /= 160
        self.kf.predict(Q=np.diag(np.square(std)))

    def update_matched(self, bbox: np.ndarray, embedding: np.ndarray) -> None:
        if self.state == TrackState.UNCONFIRMED:
            self.birth_age += 1
            if self.birth_age >= self.min_birth_age:
                self.state = TrackState.ACTIVE
        elif self.state == TrackState.INACTIVE:
            self.state = TrackState.ACTIVE
            self.inactive_age = 0
        if self.kf is None:
            self.bbox = bbox
        else:
            wh = self.kf.x[2:4] - self.kf.x[:2]
            std = np.tile(wh, 2) / 20                            # measurement noise (reference :309-313)
            self.kf.update(bbox, R=np.diag(std ** 2))
            self.bbox = self.kf.x[:4]
        unit = embedding / np.linalg.norm(embedding)
        self.embedding = (1 - self.smoothing_factor) * self.embedding + self.smoothing_factor * unit

    def update_unmatched(self) -> None:
        if self.state == TrackState.UNCONFIRMED:
            self.state = TrackState.TO_DELETE
        elif self.state == TrackState.ACTIVE:
            self.state = TrackState.INACTIVE
            self.inactive_age = 0
        elif self.state == TrackState.INACTIVE:
            self.inactive_age += 1
            if self.inactive_age >= self.max_inactive_age:
                self.state = TrackState.TO_DELETE

    def __repr__(self) -> str:
        return f"track id: {self.track_id}, bbox: {self.bbox}, label: {self.label}, embedding: {len(self.embedding)} dim"


class Tracker:
    def __init__(self, model=None, nms_kernel: int = 3, num_detections: int = 300, detection_threshold: float = 0.3,
                 reid_cost: Union[str, Callable] = "cosine", reid_threshold: float = 0.2,
                 box_cost: Union[str, Callable, None] = "iou", box_threshold: float = 0.5, smoothing_factor: float = 0.5,
                 use_kalman: bool = False, max_inactive_age: int = 30, min_birth_age: int = 2,
                 device: Union[str, torch.device, None] = None):
        self.model = model
        self.nms_kernel = nms_kernel
        self.num_detections = num_detections
        self.detection_threshold = detection_threshold
        if isinstance(reid_cost, str) and reid_cost != "cosine":
            raise NotImplementedError(f"reid_cost {reid_cost!r}: the CUDA cost kernel implements 'cosine' (pass a callable otherwise)")
        if isinstance(box_cost, str) and box_cost not in ("iou", "giou"):
            raise ValueError(f"box_cost {box_cost!r}: 'iou', 'giou', a callable or None")
        self.reid_cost = reid_cost
        self.reid_threshold = reid_threshold
        self.box_cost = box_cost
        self.box_threshold = box_threshold
        self.smoothing_factor = smoothing_factor
        self.use_kalman = use_kalman
        self.max_inactive_age = max_inactive_age
        self.min_birth_age = min_birth_age
        self._device = device
        self._costs: Optional[CostMatrices] = None
        self.reset()

    def reset(self) -> None:
        self.frame = 0
        self.next_track_id = 0
        self.tracks: List[Track] = []

    # ---- device ------------------------------------------------------------------------------------------------
    def _cost_engine(self) -> CostMatrices:
        if self._costs is None:
            dev = self._device
            if dev is None and self.model is not None:
                dev = next(self.model.parameters()).device
            self._costs = CostMatrices(dev if dev is not None else "cuda:0")
        return self._costs

    # ---- inference + association (reference :83-129) -------------------------------------------------------------
    @torch.no_grad()
    def step_batch(self, images: torch.Tensor, **kwargs) -> Dict[str, list]:
        if self.model is None:
            raise RuntimeError("step_batch needs a model; only update() works without one")
        nms_kernel = kwargs.get("nms_kernel", self.nms_kernel)
        num_detections = kwargs.get("num_detections", self.num_detections)
        self.model.eval()
        dev = next(self.model.parameters()).device
        heatmap, box_2d, reid = self.model(images.to(dev))
        det = self.model.gather_tracking2d(heatmap, box_2d, reid, nms_kernel=nms_kernel, num_detections=num_detections,
                                           normalize_bbox=True)
        det = {k: v.cpu().numpy() for k, v in det.items()}
        out = {"bboxes": [], "track_ids": []}
        for bboxes, labels, scores, emb in zip(det["bboxes"], det["labels"], det["scores"], det["embeddings"]):
            self.update(bboxes, labels, scores, emb, **kwargs)
            self.frame += 1
            out["bboxes"].append([t.bbox for t in self.tracks if t.active])
            out["track_ids"].append([t.track_id for t in self.tracks if t.active])
        return out

    @torch.no_grad()
    def step_single(self, img: torch.Tensor, **kwargs) -> Dict[str, list]:
        out = self.step_batch(img.unsqueeze(0), **kwargs)
        return {k: v[0] for k, v in out.items()}

    # ---- one frame (reference :131-201) ---------------------------------------------------------------------------
    def _cost_matrices(self, det_emb, trk_emb, det_box, trk_box):
        """(reid_cost, box_cost or None): strings -> the CUDA kernel (one launch for both), callables -> as given."""
        gpu_reid = isinstance(self.reid_cost, str)
        gpu_box = isinstance(self.box_cost, str)
        reid = box = None
        if gpu_reid or gpu_box:
            reid, box = self._cost_engine()(det_emb if gpu_reid else None, trk_emb if gpu_reid else None,
                                            det_box if gpu_box else None, trk_box if gpu_box else None,
                                            giou=self.box_cost == "giou")
        if not gpu_reid:
            reid = np.asarray(self.reid_cost(det_emb, trk_emb))
        if self.box_cost is not None and not gpu_box:
            box = np.asarray(self.box_cost(det_box, trk_box))
        return reid, box

    def update(self, bboxes: np.ndarray, labels: np.ndarray, scores: np.ndarray, embeddings: np.ndarray, **kwargs) -> None:
        detection_threshold = kwargs.get("detection_threshold", self.detection_threshold)
        reid_threshold = kwargs.get("reid_threshold", self.reid_threshold)
        box_threshold = kwargs.get("box_threshold", self.box_threshold)
        keep = scores >= detection_threshold
        det_bboxes, det_labels, det_emb = bboxes[keep], labels[keep], embeddings[keep]

        if not self.tracks:
            new_dets: Sequence[int] = range(len(det_bboxes))
        else:
            trk_emb = np.stack([t.embedding for t in self.tracks], axis=0)
            trk_box = np.stack([t.bbox for t in self.tracks], axis=0)
            reid_cost, box_cost = self._cost_matrices(det_emb, trk_emb, det_bboxes, trk_box)
            matches, free_dets, free_trks = match_with_threshold(reid_cost, reid_threshold)
            if self.box_cost is not None:
                # second chance for what appearance did not match: box overlap on the remaining detections x tracks
                sub = box_cost[np.ix_(free_dets, free_trks)] if (free_dets and free_trks) else np.zeros((len(free_dets), len(free_trks)))
                more, d_left, t_left = match_with_threshold(sub, box_threshold)
                matches += [(free_dets[d], free_trks[t]) for d, t in more]
                free_dets, free_trks = [free_dets[d] for d in d_left], [free_trks[t] for t in t_left]
            # NOTE the reference indexes the UNFILTERED arrays with indices of the filtered ones here (:186); detections
            # arrive sorted by score, so the kept ones are a prefix and both index spaces coincide
            for d, t in matches:
                self.tracks[t].update_matched(bboxes[d], embeddings[d])
            for t in free_trks:
                self.tracks[t].update_unmatched()
            new_dets = free_dets
        for d in new_dets:
            self.tracks.append(Track(self.next_track_id, det_bboxes[d], det_labels[d], det_emb[d],
                                     min_birth_age=self.min_birth_age, max_inactive_age=self.max_inactive_age,
                                     smoothing_factor=self.smoothing_factor, use_kalman=self.use_kalman))
            self.next_track_id += 1
        self.tracks = [t for t in self.tracks if not t.to_delete]
        for t in self.tracks:
            t.kalman_predict()


def build_tracker(config, model=None) -> Tracker:
    """reference tracker.py:353-357: ``config`` is the tracker section (dict) or a YAML path with a ``tracker`` key."""
    if isinstance(config, str):
        import yaml
        with open(config) as f:
            config = yaml.safe_load(f)["tracker"]
    return Tracker(model=model, **config)
