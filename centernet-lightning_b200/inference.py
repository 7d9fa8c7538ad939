"""Folder inference - the caller on the input side of the hot path.

Mirrors the reference's ``CenterNet.inference_detection(img_dir, num_detections=100)`` contract (README.md:49-65;
its source is missing from the snapshot, the closest template is models/fairmot.py:155-216 + datasets/inference.py:7-42):
files are auto-discovered and sorted, read with cv2 and converted BGR->RGB (datasets/inference.py:28-29), resized to
img_size x img_size (README.md:85 ``A.Resize``: cv2.INTER_LINEAR on the uint8 image), normalised with ImageNet statistics
(README.md:86 ``A.Normalize``), run in batches, and the result is a dict of numpy arrays ``bboxes (n_img,k,4) x1y1x2y2``,
``labels (n_img,k)``, ``scores (n_img,k)`` with boxes in resized-image pixels.

B200-first layout of the loader:
  * JPEG/PNG decode + resize run on a pool of host threads (cv2 releases the GIL) straight into PINNED uint8 HWC batches;
  * the batch crosses PCIe as uint8 (0.79 MB/image instead of 3.1 MB as fp32) and is normalised + transposed to fp32 NCHW
    by one CUDA kernel (csrc/cnl_io.cu, bit-exact with the numpy arithmetic of A.Normalize);
  * batches are double-buffered: the threads decode batch i+1 (and i+2) while the GPU runs forward+decode of batch i."""
from __future__ import annotations

import os
from collections import deque
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import preprocess


def load_resized_u8(path: str, size: int, out: Optional[np.ndarray] = None) -> np.ndarray:
    """BGR file -> RGB uint8 (size,size,3), bilinear (cv2.INTER_LINEAR).  Writes into ``out`` when given."""
    import cv2
    img = cv2.imread(path)
    if img is None:
        raise FileNotFoundError(f"cannot read image {path}")
    img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
    if out is not None:
        cv2.resize(img, (size, size), dst=out, interpolation=cv2.INTER_LINEAR)
        return out
    return cv2.resize(img, (size, size), interpolation=cv2.INTER_LINEAR)


def discover(img_dir: str, img_names: Optional[Sequence[str]] = None,
             exts=(".jpg", ".jpeg", ".png", ".bmp")) -> list:
    if not os.path.isdir(img_dir):
        raise FileNotFoundError(img_dir)                        # reference datasets/inference.py:12 asserts the same
    if img_names is None:
        img_names = sorted(f for f in os.listdir(img_dir) if f.lower().endswith(exts))
    return list(img_names)


def run_folder(net, img_dir: str, img_names, batch_size: int, num_detections: Optional[int], img_size: int,
               device: torch.device, workers: Optional[int] = None, lookahead: int = 2) -> Dict[str, np.ndarray]:
    names = discover(img_dir, img_names)
    k = num_detections or net.hparams.num_detections
    if not names:
        return {"bboxes": np.zeros((0, k, 4), np.float32), "labels": np.zeros((0, k), np.int64), "scores": np.zeros((0, k), np.float32)}
    old_k = net.hparams.num_detections
    net.hparams.num_detections = k
    net.to(device)
    workers = workers or min(32, os.cpu_count() or 4)
    n_slots = lookahead + 1
    pinned = [torch.empty((batch_size, img_size, img_size, 3), dtype=torch.uint8).pin_memory() for _ in range(n_slots)]
    views = [p.numpy() for p in pinned]
    dev_u8 = [torch.empty((batch_size, img_size, img_size, 3), dtype=torch.uint8, device=device) for _ in range(2)]
    x = torch.empty((batch_size, 3, img_size, img_size), dtype=torch.float32, device=device)
    chunks = [names[s:s + batch_size] for s in range(0, len(names), batch_size)]
    out = {"bboxes": [], "labels": [], "scores": []}
    pool = ThreadPoolExecutor(max_workers=workers)

    def submit(b):
        slot = b % n_slots
        futs = [pool.submit(load_resized_u8, os.path.join(img_dir, n), img_size, views[slot][i]) for i, n in enumerate(chunks[b])]
        if len(chunks[b]) < batch_size:
            views[slot][len(chunks[b]):] = 0                     # fixed batch shape -> one engine / CUDA graph
        return futs

    try:
        pending = deque(submit(b) for b in range(min(lookahead, len(chunks))))
        for b, chunk in enumerate(chunks):
            for f in pending.popleft():
                f.result()                                       # re-raises loader errors
            if b + lookahead < len(chunks):
                pending.append(submit(b + lookahead))            # slot (b+lookahead) % n_slots was consumed at step b-1
            d8 = dev_u8[b & 1]
            d8.copy_(pinned[b % n_slots], non_blocking=True)
            preprocess.normalize_u8(d8, out=x)
            det = net.detect(x)
            out["bboxes"].append(det["boxes"][:len(chunk)].cpu().numpy())      # the D2H reads also order the reuse of the pinned slot
            out["labels"].append(det["labels"][:len(chunk)].cpu().numpy())
            out["scores"].append(det["scores"][:len(chunk)].cpu().numpy())
    finally:
        pool.shutdown(wait=True, cancel_futures=True)
        net.hparams.num_detections = old_k
    return {key: np.concatenate(v, axis=0) for key, v in out.items()}
