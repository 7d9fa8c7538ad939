"""Host side of the fused decode: torch tensors in/out, libcnl_b200.so underneath.

Mirrors the reference's decode interface (names, argument meaning, error behaviour):

* ``decode_detections``       <- CenterNet.decode_detections        (reference models/centernet.py:229-241)
* ``get_topk_from_heatmap``   <- CenterNet.get_topk_from_heatmap    (:243-261)
* ``gather_and_decode_boxes`` <- CenterNet.gather_and_decode_boxes  (:263-304, staticmethod)
* ``reid=`` argument          <- EmbeddingHead.gather_at_indices    (reference models/fairmot.py:63-73)

torch is used for device memory and the current stream only; every arithmetic step runs in
csrc/cnl_decode.cu.  There is no fallback: CPU tensors raise.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import _lib

_workspaces: Dict[Tuple[int, int], torch.Tensor] = {}


def _ws(device: torch.device, nbytes: int) -> torch.Tensor:
    """Scratch for the allocating entry points, one per (device, stream): two decodes running concurrently on different
    streams must not share the histogram / candidate map."""
    index = device.index if device.index is not None else torch.cuda.current_device()
    key = (index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20) + 256, dtype=torch.uint8, device=device)
        ws = ws[(-ws.data_ptr()) % 256:]
        _workspaces[key] = ws
    return ws


def _check_map(name: str, t: torch.Tensor, ndim: int = 4) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} is on {t.device}: the cnl_b200 decode runs on CUDA only (no CPU fallback)")
    if t.dtype != torch.float32:
        raise ValueError(f"{name} must be float32 (got {t.dtype})")
    if t.dim() != ndim:
        raise ValueError(f"{name} must have {ndim} dims (got shape {tuple(t.shape)})")
    return t.contiguous()


class DecodeBuffers:
    """Pre-allocated outputs + workspace for one decode shape (used inside CUDA-graph capture: no allocation per call)."""

    def __init__(self, n: int, h: int, w: int, k: int, reid_dim: int, device: torch.device,
                 packed: Optional[torch.Tensor] = None):
        lib = _lib.load()
        self.n, self.h, self.w, self.k, self.e = n, h, w, k, reid_dim
        with torch.cuda.device(device):
            self.scores = torch.empty((n, k), dtype=torch.float32, device=device)
            self.labels = torch.empty((n, k), dtype=torch.int64, device=device)
            self.indices = torch.empty((n, k), dtype=torch.int64, device=device)
            self.boxes = torch.empty((n, k, 4), dtype=torch.float32, device=device)
            self.emb = torch.empty((n, k, reid_dim), dtype=torch.float32, device=device) if reid_dim else None
            self.ws = torch.empty(lib.cnl_decode_workspace_bytes_k(n, h, w, k) + 256, dtype=torch.uint8, device=device)
            off = (-self.ws.data_ptr()) % 256
            self.ws = self.ws[off:]
        # optional packed rows (n, k, 8 + reid_dim) written by the select kernel itself: the send buffer of the cross-rank
        # all_gather (distributed.DetectionGather), so that no torch kernel has to pack the detections
        self.packed = packed
        if packed is not None:
            if tuple(packed.shape) != (n, k, 8 + reid_dim) or packed.dtype != torch.float32 or not packed.is_contiguous():
                raise ValueError(f"packed must be a contiguous float32 tensor of shape {(n, k, 8 + reid_dim)}")
        self.clean = False          # True once a decode has run on this workspace: its histogram is left zeroed (no memset needed)

    def as_dict(self) -> Dict[str, torch.Tensor]:
        out = {"boxes": self.boxes, "scores": self.scores, "labels": self.labels}
        if self.emb is not None:
            out["embeddings"] = self.emb
        return out


def decode_into(bufs: DecodeBuffers, heatmap: torch.Tensor, box_offsets: torch.Tensor, reid: Optional[torch.Tensor], *,
                num_detections: int, nms_kernel: int, normalize_boxes: bool, box_log: bool, box_multiplier: float,
                stride: int, from_logits: bool, _peaks_only: bool = False) -> int:
    """Allocation-free decode on the current stream (CUDA-graph capturable).  Returns the number of launches."""
    lib = _lib.load()
    n, c, h, w = heatmap.shape
    if (n, h, w, num_detections) != (bufs.n, bufs.h, bufs.w, bufs.k):
        raise ValueError("DecodeBuffers were built for another shape")
    flags = int(bool(from_logits)) | (2 if bufs.clean else 0)           # CNL_DECODE_WORKSPACE_CLEAN
    if _peaks_only:
        flags = int(bool(from_logits)) | 4                              # CNL_DECODE_PEAKS_ONLY (profiling: memset + peaks kernel)
    bufs.clean = False
    st = lib.cnl_decode_detections_packed(
        heatmap.data_ptr(), box_offsets.data_ptr(), reid.data_ptr() if reid is not None else None,
        n, c, h, w, bufs.e if reid is not None else 0, flags, int(nms_kernel), int(num_detections),
        int(bool(normalize_boxes)), int(bool(box_log)), float(box_multiplier), int(stride),
        bufs.boxes.data_ptr(), bufs.scores.data_ptr(), bufs.labels.data_ptr(), bufs.indices.data_ptr(),
        bufs.emb.data_ptr() if (bufs.emb is not None and reid is not None) else None,
        bufs.packed.data_ptr() if bufs.packed is not None else None, bufs.packed.shape[-1] if bufs.packed is not None else 0,
        bufs.ws.data_ptr(), bufs.ws.numel(), torch.cuda.current_stream(heatmap.device).cuda_stream)
    _lib.check(st, "cnl_decode_detections")
    bufs.clean = not _peaks_only
    return 2 if flags & 2 else 3            # [memset +] peaks + select


def boxes_xyxy_to_xywh(boxes: torch.Tensor) -> torch.Tensor:
    """(...,4) xyxy -> xywh through the library (reference models/centernet.py:207, torchvision box_convert)."""
    lib = _lib.load()
    if not boxes.is_cuda or boxes.dtype != torch.float32 or boxes.shape[-1] != 4:
        raise ValueError("boxes must be a CUDA float32 tensor (...,4)")
    boxes = boxes.contiguous()
    out = torch.empty_like(boxes)
    st = lib.cnl_boxes_xyxy_to_xywh(boxes.data_ptr(), out.data_ptr(), boxes.numel() // 4,
                                    torch.cuda.current_stream(boxes.device).cuda_stream)
    _lib.check(st, "cnl_boxes_xyxy_to_xywh")
    return out


def sigmoid(x: torch.Tensor) -> torch.Tensor:
    """fp32 logistic through the library (reference models/centernet.py:205)."""
    lib = _lib.load()
    x = _check_map("heatmap", x, x.dim())
    out = torch.empty_like(x)
    st = lib.cnl_sigmoid(x.data_ptr(), out.data_ptr(), x.numel(), torch.cuda.current_stream(x.device).cuda_stream)
    _lib.check(st, "cnl_sigmoid")
    return out


def decode_detections(heatmap: torch.Tensor, box_offsets: Optional[torch.Tensor], *, num_detections: int = 100,
                      nms_kernel: int = 3, normalize_boxes: bool = False, box_log: bool = False,
                      box_multiplier: float = 1.0, stride: int = 4, reid: Optional[torch.Tensor] = None,
                      from_logits: bool = False, _force_generic: bool = False) -> Dict[str, torch.Tensor]:
    """Returns {"boxes" (N,k,4) f32, "scores" (N,k) f32, "labels" (N,k) i64, "indices" (N,k) i64[, "embeddings" (N,k,E)]}.

    ``heatmap`` holds probabilities (reference semantics) unless ``from_logits`` is set, in which case the
    kernel applies the logistic itself (the fused form of ``outputs['heatmap'].sigmoid()``, reference :205)."""
    lib = _lib.load()
    heatmap = _check_map("heatmap", heatmap)
    n, c, h, w = heatmap.shape
    dev = heatmap.device
    if box_offsets is not None:
        box_offsets = _check_map("box_offsets", box_offsets)
        if tuple(box_offsets.shape) != (n, 4, h, w):
            raise ValueError(f"box_offsets must be {(n, 4, h, w)}, got {tuple(box_offsets.shape)}")
    e = 0
    if reid is not None:
        reid = _check_map("reid", reid)
        if reid.shape[0] != n or tuple(reid.shape[2:]) != (h, w):
            raise ValueError(f"reid must be (N,E,{h},{w}), got {tuple(reid.shape)}")
        e = reid.shape[1]
    k = int(num_detections)
    if not 1 <= k <= h * w:
        raise ValueError(f"num_detections={k} must be in 1..H*W={h * w} (torch.topk raises in the reference)")
    with torch.cuda.device(dev):
        scores = torch.empty((n, k), dtype=torch.float32, device=dev)
        labels = torch.empty((n, k), dtype=torch.int64, device=dev)
        indices = torch.empty((n, k), dtype=torch.int64, device=dev)
        boxes = torch.empty((n, k, 4), dtype=torch.float32, device=dev) if box_offsets is not None else None
        emb = torch.empty((n, k, e), dtype=torch.float32, device=dev) if reid is not None else None
        nbytes = lib.cnl_decode_workspace_bytes_k(n, h, w, k)
        ws = _ws(dev, nbytes)
        st = lib.cnl_decode_detections(
            heatmap.data_ptr(), box_offsets.data_ptr() if box_offsets is not None else None,
            reid.data_ptr() if reid is not None else None,
            n, c, h, w, e, int(bool(from_logits)), -int(nms_kernel) if _force_generic else int(nms_kernel), k,
            int(bool(normalize_boxes)), int(bool(box_log)), float(box_multiplier), int(stride),
            boxes.data_ptr() if boxes is not None else None, scores.data_ptr(), labels.data_ptr(), indices.data_ptr(),
            emb.data_ptr() if emb is not None else None,
            ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(st, "cnl_decode_detections")
    out = {"scores": scores, "labels": labels, "indices": indices}
    if boxes is not None:
        out["boxes"] = boxes
    if emb is not None:
        out["embeddings"] = emb
    return out


def get_topk_from_heatmap(heatmap: torch.Tensor, num_detections: int = 100, nms_kernel: int = 3,
                          pseudo_nms: bool = True, from_logits: bool = False):
    """(scores, indices, labels), each (N,k) - reference models/centernet.py:243-261."""
    out = decode_detections(heatmap, None, num_detections=num_detections,
                            nms_kernel=nms_kernel if pseudo_nms else 1, from_logits=from_logits)
    return out["scores"], out["indices"], out["labels"]


def gather_and_decode_boxes(box_offsets: torch.Tensor, indices: torch.Tensor, normalize_boxes: bool = False,
                            box_log: bool = False, box_multiplier: float = 1.0, stride: int = 4) -> torch.Tensor:
    """reference models/centernet.py:263-304.  Batch dim optional, as in the reference ((4,H,W) + (k,))."""
    lib = _lib.load()
    squeeze = box_offsets.dim() == 3
    if squeeze:
        box_offsets, indices = box_offsets.unsqueeze(0), indices.unsqueeze(0)
    box_offsets = _check_map("box_offsets", box_offsets)
    n, four, h, w = box_offsets.shape
    if four != 4:
        raise ValueError("box_offsets must have 4 channels (left, top, right, bottom)")
    if indices.dtype != torch.int64 or indices.dim() != 2 or indices.shape[0] != n or not indices.is_cuda:
        raise ValueError("indices must be a CUDA int64 tensor of shape (N, k)")
    indices = indices.contiguous()
    k = indices.shape[1]
    dev = box_offsets.device
    with torch.cuda.device(dev):
        boxes = torch.empty((n, k, 4), dtype=torch.float32, device=dev)
        st = lib.cnl_gather_boxes(box_offsets.data_ptr(), indices.data_ptr(), n, h, w, k, int(bool(normalize_boxes)),
                                  int(bool(box_log)), float(box_multiplier), int(stride), boxes.data_ptr(),
                                  torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(st, "cnl_gather_boxes")
    return boxes[0] if squeeze else boxes
