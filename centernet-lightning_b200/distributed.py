"""Multi-GPU: images shard across ranks (one process per GPU, weights replicated); the only exchange on the path is
one fixed-shape NCCL all_gather of the final detections.  The reference's analogue is the
``dist.all_gather_object`` of per-image prediction dicts in eval/coco.py:10-18 - a pickled Python list; here it is a
single packed float32 tensor (B_local, k, 6[+E]) per rank: 76.8 KB at B_local=32, k=100 - latency-bound on NVSwitch."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous batch split; the first (total % world) ranks take one extra image."""
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


PACK_FIXED = 8        # lanes before the embedding: x1 y1 x2 y2 score bits(index i32) bits(label lo) bits(label hi)


def pack_detections(det: Dict[str, torch.Tensor], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """boxes (B,k,4) f32 | scores (B,k) f32 | labels (B,k) i64 [| indices (B,k) i64] [| embeddings (B,k,E)] -> (B,k,8+E) f32
    rows in the layout the select kernel writes itself (cnl_decode_detections_packed); integer lanes travel bit-exactly.
    Host-side / CPU form (gloo tests, callers without a DetectionGather)."""
    boxes, scores, labels = det["boxes"], det["scores"], det["labels"]
    b, k = scores.shape
    emb = det.get("embeddings")
    width = PACK_FIXED + (emb.shape[-1] if emb is not None else 0)
    if out is None:
        out = torch.empty((b, k, width), dtype=torch.float32, device=scores.device)
    out[..., 0:4] = boxes
    out[..., 4] = scores
    idx = det.get("indices")
    out[..., 5] = (idx if idx is not None else torch.zeros_like(labels)).to(torch.int32).view(torch.float32)
    out[..., 6:8] = labels.to(torch.int64).contiguous().view(torch.int32).view(b, k, 2).view(torch.float32)
    if emb is not None:
        out[..., PACK_FIXED:] = emb
    return out


def unpack_detections(packed: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Zero-copy views into packed rows: boxes (B,k,4) / scores (B,k) strided float32, labels (B,k) strided int64,
    indices (B,k) strided int32 [, embeddings (B,k,E)].  No kernel is launched."""
    b, k, width = packed.shape
    out = {"boxes": packed[..., 0:4], "scores": packed[..., 4],
           "labels": packed.view(torch.int64)[..., 3],          # lanes 6,7 of every 8+E (even) float row
           "indices": packed.view(torch.int32)[..., 5]}
    if width > PACK_FIXED:
        out["embeddings"] = packed[..., PACK_FIXED:]
    return out


class DetectionGather:
    """Pre-allocated all_gather of packed detections (equal shard sizes).  ``local`` is handed to the decode as its packed
    output (``CenterNet.detect(..., packed_out=gather.local)``), so the step launches no torch kernel around the one
    NCCL all_gather; the result is a dict of views into ``full``."""

    def __init__(self, batch_local: int, k: int, emb_dim: int, device: torch.device, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.local = torch.empty((batch_local, k, PACK_FIXED + emb_dim), dtype=torch.float32, device=device)
        self.full = torch.empty((self.world * batch_local, k, PACK_FIXED + emb_dim), dtype=torch.float32, device=device)
        self.views = unpack_detections(self.full)

    def __call__(self, det: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        if det is not None:                         # detections that were not written into self.local by the kernel
            pack_detections(det, self.local)
        dist.all_gather_into_tensor(self.full, self.local, group=self.group)
        return self.views


def gather_detections(det: Dict[str, torch.Tensor], group=None) -> Dict[str, torch.Tensor]:
    """Convenience (allocating) form; works on CPU tensors with the gloo backend as well (used by the CPU tests)."""
    world = dist.get_world_size(group)
    local = pack_detections(det)
    parts: List[torch.Tensor] = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(parts, local, group=group)
    return {k: v.contiguous() for k, v in unpack_detections(torch.cat(parts, dim=0)).items()}


def detect_sharded(net, images: torch.Tensor, group=None) -> Dict[str, torch.Tensor]:
    """Every rank holds (or receives) the full batch, runs its contiguous shard and gathers all detections."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if images.shape[0] % world:
        raise ValueError("batch must divide evenly across ranks for the fixed-shape all_gather")
    s, e = shard_range(images.shape[0], rank, world)
    det = net.detect(images[s:e])
    return gather_detections(det, group)
