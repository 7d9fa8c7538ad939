"""Tracker association costs: the CUDA path (staging + cnl_track_cost_matrices + D2H, what Tracker.update pays per frame)
against the reference's host path (scipy cdist "cosine" + its utils/box.py IoU/GIoU arithmetic, restated in
oracle/tracker_np.py) on this box's host cores.  python tools/bench_tracker_costs.py  -> one line per shape.

VERDICT r1 item 9: keep the GPU path only where it wins."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from centernet_lightning_b200.tracker import CostMatrices  # noqa: E402
from oracle import tracker_np  # noqa: E402


def main():
    from scipy.spatial.distance import cdist
    eng = CostMatrices("cuda:0")
    rng = np.random.default_rng(0)
    for n_det, n_trk, dim in ((300, 50, 64), (300, 300, 64), (100, 30, 64), (300, 300, 128)):
        de, te = rng.standard_normal((n_det, dim)), rng.standard_normal((n_trk, dim))
        c = rng.uniform(0.1, 0.9, (n_det, 2)); s = rng.uniform(0.02, 0.2, (n_det, 2))
        db = np.concatenate([c - s / 2, c + s / 2], 1)
        c = rng.uniform(0.1, 0.9, (n_trk, 2)); s = rng.uniform(0.02, 0.2, (n_trk, 2))
        tb = np.concatenate([c - s / 2, c + s / 2], 1)
        for _ in range(3):
            eng(de, te, db, tb, giou=False)
        reps = 30
        t0 = time.perf_counter()
        for _ in range(reps):
            r_gpu, b_gpu = eng(de, te, db, tb, giou=False)
        gpu_ms = (time.perf_counter() - t0) / reps * 1e3
        t0 = time.perf_counter()
        for _ in range(reps):
            r_cpu = cdist(de, te, "cosine")
            b_cpu = tracker_np.box_iou_distance_matrix(db, tb)
        cpu_ms = (time.perf_counter() - t0) / reps * 1e3
        t0 = time.perf_counter()
        for _ in range(reps):
            cdist(de, te, "cosine")
        cdist_ms = (time.perf_counter() - t0) / reps * 1e3
        same = bool(np.allclose(r_gpu, r_cpu, rtol=0, atol=1e-13) and np.array_equal(b_gpu, b_cpu))
        print(f"n_det {n_det} n_trk {n_trk} dim {dim}: cuda path {gpu_ms:.3f} ms/frame (H2D + kernel + D2H, synchronous) | "
              f"host scipy cdist {cdist_ms:.3f} ms + box costs = {cpu_ms:.3f} ms | equal {same}", flush=True)


if __name__ == "__main__":
    main()
