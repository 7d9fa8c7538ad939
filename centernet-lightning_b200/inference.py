"""Folder inference - the caller on the input side of the hot path.

Mirrors the reference's ``CenterNet.inference_detection(img_dir, num_detections=100)`` contract (README.md:49-65;
its source is missing from the snapshot, the closest template is models/fairmot.py:155-216 + datasets/inference.py:7-42):
files are auto-discovered and sorted, resized to img_size x img_size, normalised with ImageNet statistics
(README.md:80-86 A.Resize + A.Normalize), run in batches, and the result is a dict of numpy arrays
``bboxes (n_img,k,4) x1y1x2y2``, ``labels (n_img,k)``, ``scores (n_img,k)`` with boxes in resized-image pixels."""
from __future__ import annotations

import os
from typing import Dict, Optional, Sequence

import numpy as np
import torch

_MEAN = np.array([0.485, 0.456, 0.406], np.float32)
_STD = np.array([0.229, 0.224, 0.225], np.float32)


def load_image(path: str, size: int) -> np.ndarray:
    """RGB, bilinear resize to size x size, ImageNet normalisation -> (3,size,size) float32."""
    from PIL import Image
    with Image.open(path) as im:
        im = im.convert("RGB").resize((size, size), Image.BILINEAR)
        a = np.asarray(im, dtype=np.float32) / 255.0
    a = (a - _MEAN) / _STD
    return np.ascontiguousarray(a.transpose(2, 0, 1))


def discover(img_dir: str, img_names: Optional[Sequence[str]] = None,
             exts=(".jpg", ".jpeg", ".png", ".bmp")) -> list:
    if not os.path.isdir(img_dir):
        raise FileNotFoundError(img_dir)                        # reference datasets/inference.py:12 asserts the same
    if img_names is None:
        img_names = sorted(f for f in os.listdir(img_dir) if f.lower().endswith(exts))
    return list(img_names)


def run_folder(net, img_dir: str, img_names, batch_size: int, num_detections: Optional[int], img_size: int,
               device: torch.device) -> Dict[str, np.ndarray]:
    names = discover(img_dir, img_names)
    k = num_detections or net.hparams.num_detections
    old_k = net.hparams.num_detections
    net.hparams.num_detections = k
    net.to(device)
    out = {"bboxes": [], "labels": [], "scores": []}
    pinned = torch.empty((batch_size, 3, img_size, img_size), dtype=torch.float32).pin_memory()
    try:
        for s in range(0, len(names), batch_size):
            chunk = names[s:s + batch_size]
            for i, n in enumerate(chunk):
                pinned[i].copy_(torch.from_numpy(load_image(os.path.join(img_dir, n), img_size)))
            if len(chunk) < batch_size:
                pinned[len(chunk):].zero_()                      # fixed batch shape -> one engine / CUDA graph
            det = net.detect(pinned.to(device, non_blocking=True))
            out["bboxes"].append(det["boxes"][:len(chunk)].cpu().numpy())
            out["labels"].append(det["labels"][:len(chunk)].cpu().numpy())
            out["scores"].append(det["scores"][:len(chunk)].cpu().numpy())
    finally:
        net.hparams.num_detections = old_k
    if not names:
        return {"bboxes": np.zeros((0, k, 4), np.float32), "labels": np.zeros((0, k), np.int64), "scores": np.zeros((0, k), np.float32)}
    return {key: np.concatenate(v, axis=0) for key, v in out.items()}
