"""Decode oracle, ATen flavour: the same op sequence the reference executes, on CPU tensors.

TEST INFRASTRUCTURE ONLY (see oracle/spec_model.py header).  This is the "port" that
``bench.py`` times as the CPU baseline for the decode half of the path: it issues the
same ATen calls as /root/reference/centernet_lightning/models/centernet.py:243-304
(max_pool2d, eq, mul, max(dim=1), topk, gather, remainder, floor-div, clamp_min, stack),
so its cost and its bits are those of the reference's CPU PyTorch path.
tests/test_oracle.py checks it against the golden outputs of the imported reference.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F


def topk_from_heatmap(heatmap: torch.Tensor, num_detections: int = 100, nms_kernel: int = 3):
    """centernet.py:243-261."""
    n = heatmap.shape[0]
    pad = (nms_kernel - 1) // 2
    pooled = F.max_pool2d(heatmap, kernel_size=nms_kernel, stride=1, padding=pad)      # :252
    heatmap = heatmap * (pooled == heatmap)                                            # :252-253
    best, labels = torch.max(heatmap, dim=1)                                           # :254
    scores, indices = torch.topk(best.view(n, -1), num_detections)                     # :257-259
    labels = torch.gather(labels.view(n, -1), dim=-1, index=indices)                   # :260
    return scores, indices, labels


def gather_and_decode_boxes(box_offsets: torch.Tensor, indices: torch.Tensor, normalize_boxes: bool = False,
                            box_log: bool = False, box_multiplier: float = 1.0, stride: int = 4) -> torch.Tensor:
    """centernet.py:263-304 (whole-map exp/mul/clamp first, then four gathers, like the reference)."""
    h, w = box_offsets.shape[-2:]
    cx = torch.remainder(indices, w) + 0.5                                             # :278
    cy = torch.div(indices, w, rounding_mode="floor") + 0.5                            # :279
    off = box_offsets.flatten(start_dim=-2)                                            # :282
    if box_log:
        off = torch.exp(off)                                                           # :284
    off = (off * box_multiplier).clamp_min(0)                                          # :285-286
    sides = [torch.gather(off[..., c, :], dim=-1, index=indices) for c in range(4)]    # :293-296
    boxes = torch.stack((cx - sides[0], cy - sides[1], cx + sides[2], cy + sides[3]), dim=-1)
    if normalize_boxes:
        boxes[..., [0, 2]] /= w                                                        # :300
        boxes[..., [1, 3]] /= h                                                        # :301
    else:
        boxes *= stride                                                                # :303
    return boxes


def gather_embeddings(reid: torch.Tensor, indices: torch.Tensor) -> torch.Tensor:
    """fairmot.py:63-73."""
    n, e = reid.shape[:2]
    idx = indices.unsqueeze(1).expand(n, e, -1)
    return torch.gather(reid.view(n, e, -1), dim=-1, index=idx).swapaxes(1, 2)


def decode_detections(heatmap: torch.Tensor, box_offsets: torch.Tensor, *, num_detections: int = 100,
                      nms_kernel: int = 3, normalize_boxes: bool = False, box_log: bool = False,
                      box_multiplier: float = 1.0, stride: int = 4,
                      reid: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """centernet.py:229-241 (+ fairmot.py:138-151)."""
    scores, indices, labels = topk_from_heatmap(heatmap, num_detections, nms_kernel)
    boxes = gather_and_decode_boxes(box_offsets, indices, normalize_boxes, box_log, box_multiplier, stride)
    out = {"boxes": boxes, "scores": scores, "labels": labels, "indices": indices}
    if reid is not None:
        out["embeddings"] = gather_embeddings(reid, indices)
    return out
