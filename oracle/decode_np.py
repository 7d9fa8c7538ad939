"""Decode oracle: numpy restatement of the reference's heatmap -> detections path.

TEST INFRASTRUCTURE ONLY (see oracle/spec_model.py header for who may import it).

Follows, line by line (paths relative to /root/reference):

* ``centernet_lightning/models/centernet.py:243-261``  get_topk_from_heatmap
* ``centernet_lightning/models/centernet.py:263-304``  gather_and_decode_boxes
* ``centernet_lightning/models/centernet.py:229-241``  decode_detections
* ``centernet_lightning/models/fairmot.py:63-73``      EmbeddingHead.gather_at_indices

PINNING: the reference has no test or golden vector for decode, so
this file is pinned against the reference's *own code run in the authoring
container*: tests/golden/gen_golden.py imports the unmodified
``CenterNet.decode_detections`` under stub modules and stores its outputs in
tests/golden/*.npz; tests/test_oracle.py checks this restatement against them
bit for bit.

Tie order.  ``torch.topk`` returns equal scores in an unspecified order, so the
reference's (index, label) order is only defined when the top-(k+1) scores are
distinct.  This oracle (and the CUDA kernel) use the canonical order
"score descending, then flat index ascending"; ``same_detections`` compares two
results exactly on scores and as sets within each equal-score group.
"""
from __future__ import annotations

import numpy as np


def _maxpool_same(h: np.ndarray, k: int) -> np.ndarray:
    """F.max_pool2d(h, k, stride=1, padding=(k-1)//2) with -inf padding (centernet.py:251-252)."""
    if k % 2 != 1:
        raise ValueError("nms_kernel must be odd (an even kernel changes the map size in the reference)")
    p = (k - 1) // 2
    if p == 0:
        return h.copy()
    H, W = h.shape[-2:]
    pad = np.full(h.shape[:-2] + (H + 2 * p, W + 2 * p), -np.inf, dtype=h.dtype)
    pad[..., p:p + H, p:p + W] = h
    out = np.full_like(h, -np.inf)
    for dy in range(k):
        for dx in range(k):
            np.maximum(out, pad[..., dy:dy + H, dx:dx + W], out=out)
    return out


def candidate_map(heatmap: np.ndarray, nms_kernel: int = 3, pseudo_nms: bool = True):
    """centernet.py:250-254: per-pixel (best score, label) after pseudo-NMS and the max over classes."""
    h = np.asarray(heatmap, dtype=np.float32)
    if pseudo_nms:
        h = h * (_maxpool_same(h, nms_kernel) == h).astype(np.float32)
    labels = np.argmax(h, axis=1)
    best = np.take_along_axis(h, labels[:, None], axis=1)[:, 0]
    return best, labels


def topk_from_heatmap(heatmap: np.ndarray, num_detections: int = 100, nms_kernel: int = 3,
                      pseudo_nms: bool = True):
    """centernet.py:243-261.  heatmap (N,C,H,W) float32 probabilities.

    Returns scores (N,k) f32 descending, indices (N,k) i64 (flat y*W+x), labels (N,k) i64."""
    h = np.asarray(heatmap, dtype=np.float32)
    N, C, H, W = h.shape
    if num_detections > H * W:
        raise ValueError("num_detections must be <= H*W (torch.topk would raise)")
    if pseudo_nms:
        keep = _maxpool_same(h, nms_kernel) == h            # :252
        h = h * keep.astype(np.float32)                     # :253 (non-peaks -> exactly 0)
    labels = np.argmax(h, axis=1)                           # :254 first maximal class
    best = np.take_along_axis(h, labels[:, None], axis=1)[:, 0]
    best = best.reshape(N, -1)                              # :257
    labels = labels.reshape(N, -1)
    # canonical order: score desc, index asc  (stable sort on -score)
    order = np.argsort(-best, axis=1, kind="stable")[:, :num_detections]
    scores = np.take_along_axis(best, order, axis=1)
    lab = np.take_along_axis(labels, order, axis=1)         # :260
    return scores.astype(np.float32), order.astype(np.int64), lab.astype(np.int64)


def gather_and_decode_boxes(box_offsets: np.ndarray, indices: np.ndarray, normalize_boxes: bool = False,
                            box_log: bool = False, box_multiplier: float = 1.0, stride: int = 4) -> np.ndarray:
    """centernet.py:263-304.  box_offsets (N,4,H,W) f32, indices (N,k) i64 -> (N,k,4) f32.

    Every step is a separate float32 rounding (no FMA), as in ATen's op-by-op evaluation."""
    b = np.asarray(box_offsets, dtype=np.float32)
    idx = np.asarray(indices, dtype=np.int64)
    H, W = b.shape[-2:]
    cx = (idx % W).astype(np.float32) + np.float32(0.5)     # :278
    cy = (idx // W).astype(np.float32) + np.float32(0.5)    # :279
    flat = b.reshape(b.shape[:-2] + (H * W,))               # :282
    g = np.take_along_axis(flat, idx[..., None, :].repeat(4, axis=-2), axis=-1)  # gather commutes with the elementwise ops
    if box_log:
        g = np.exp(g, dtype=np.float32)                     # :283-284
    g = g * np.float32(box_multiplier)                      # :285
    g = np.maximum(g, np.float32(0))                        # :286
    x1 = cx - g[..., 0, :]                                  # :293
    y1 = cy - g[..., 1, :]
    x2 = cx + g[..., 2, :]
    y2 = cy + g[..., 3, :]
    boxes = np.stack((x1, y1, x2, y2), axis=-1).astype(np.float32)
    if normalize_boxes:                                     # :299-301
        boxes[..., [0, 2]] /= np.float32(W)
        boxes[..., [1, 3]] /= np.float32(H)
    else:
        boxes *= np.float32(stride)                         # :303
    return boxes


def gather_embeddings(reid: np.ndarray, indices: np.ndarray) -> np.ndarray:
    """fairmot.py:63-73.  reid (N,E,H,W), indices (N,k) -> (N,k,E); no normalisation."""
    r = np.asarray(reid, dtype=np.float32)
    N, E = r.shape[:2]
    flat = r.reshape(N, E, -1)
    g = np.take_along_axis(flat, np.asarray(indices)[:, None, :].repeat(E, axis=1), axis=-1)
    return np.ascontiguousarray(np.swapaxes(g, 1, 2))


def decode_detections(heatmap: np.ndarray, box_offsets: np.ndarray, *, num_detections: int = 100,
                      nms_kernel: int = 3, normalize_boxes: bool = False, box_log: bool = False,
                      box_multiplier: float = 1.0, stride: int = 4, reid: np.ndarray | None = None):
    """centernet.py:229-241 (+ fairmot.py:138-151 when reid is given)."""
    scores, indices, labels = topk_from_heatmap(heatmap, num_detections, nms_kernel)
    boxes = gather_and_decode_boxes(box_offsets, indices, normalize_boxes, box_log, box_multiplier, stride)
    out = {"boxes": boxes, "scores": scores, "labels": labels, "indices": indices}
    if reid is not None:
        out["embeddings"] = gather_embeddings(reid, indices)
    return out


def sigmoid_f32(x: np.ndarray) -> np.ndarray:
    """The logistic function the fused-from-logits kernel specifies: 1/(1+exp(-x)) in float32."""
    x = np.asarray(x, dtype=np.float32)
    return (np.float32(1) / (np.float32(1) + np.exp(-x, dtype=np.float32))).astype(np.float32)


def same_detections(a: dict, b: dict, *, score_tol: float = 0.0, box_tol: float = 0.0) -> tuple[bool, str]:
    """Compare two decode results.  Scores must match (exactly when score_tol == 0);
    (index,label,box) rows must match as sets inside every run of equal scores - the only
    freedom the reference's torch.topk leaves."""
    sa, sb = np.asarray(a["scores"]), np.asarray(b["scores"])
    if sa.shape != sb.shape:
        return False, f"shape {sa.shape} vs {sb.shape}"
    if score_tol == 0.0:
        if not np.array_equal(sa, sb):
            bad = np.argwhere(sa != sb)[0]
            return False, f"scores differ at {tuple(bad)}: {sa[tuple(bad)]!r} vs {sb[tuple(bad)]!r}"
    elif not np.allclose(sa, sb, rtol=0, atol=score_tol):
        return False, f"scores differ by {np.abs(sa - sb).max()}"
    N, k = sa.shape
    for n in range(N):
        j = 0
        while j < k:
            e = j + 1
            while e < k and sa[n, e] == sa[n, j]:
                e += 1
            last_group_open = (e == k)          # the group may continue past k: only subset relation is defined
            ra = _rows(a, n, j, e)
            rb = _rows(b, n, j, e)
            if e - j == 1 or not last_group_open:
                if not _rows_equal_as_sets(ra, rb, box_tol):
                    return False, f"image {n} ranks [{j},{e}) differ:\n{ra}\nvs\n{rb}"
            # an equal-score group cut by k: both are valid selections; nothing further to check
            j = e
    return True, "ok"


def _rows(d: dict, n: int, j: int, e: int) -> np.ndarray:
    cols = []
    if "indices" in d:
        cols.append(np.asarray(d["indices"])[n, j:e, None].astype(np.float64))
    cols.append(np.asarray(d["labels"])[n, j:e, None].astype(np.float64))
    cols.append(np.asarray(d["boxes"])[n, j:e].astype(np.float64))
    return np.concatenate(cols, axis=1)


def _rows_equal_as_sets(ra: np.ndarray, rb: np.ndarray, tol: float) -> bool:
    if ra.shape != rb.shape:
        return False
    ka = np.lexsort(ra.T[::-1])
    kb = np.lexsort(rb.T[::-1])
    return np.allclose(ra[ka], rb[kb], rtol=0, atol=tol)
