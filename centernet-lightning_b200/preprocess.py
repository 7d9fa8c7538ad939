"""GPU side of the image loader: uint8 HWC batches -> normalised fp32 NCHW (csrc/cnl_io.cu).

Mirrors ``A.Normalize()`` + ``ToTensorV2()`` of the reference's inference transform (README.md:84-87).  The kernel
reproduces, rounding for rounding, the float32 arithmetic of albumentations 1.x ``functional.normalize`` as restated in
oracle/preprocess_np.py (float32 ``mean*255`` / ``std*255``, float32 reciprocal, float32 subtract then multiply), so a
uint8 batch copied to the GPU (1 byte per sample over PCIe) yields the tensor that restatement yields.  albumentations
itself is not installed here and not pinned by the reference (requirements.txt), so agreement with the real library is
UNPINNED: expected to the last bit for 1.x, not verified."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def normalize_constants(mean: Sequence[float] = IMAGENET_MEAN, std: Sequence[float] = IMAGENET_STD, max_pixel_value: float = 255.0):
    """(mean*max, 1/(std*max)) as albumentations 1.x builds them: float32 arrays multiplied in float32, float32 reciprocal.
    The mean is handed to the C ABI as float64 (exact widening): float(double(x) - mean) is the float32 subtraction."""
    mean255 = np.array(mean, dtype=np.float32)
    mean255 *= np.float32(max_pixel_value)
    std255 = np.array(std, dtype=np.float32)
    std255 *= np.float32(max_pixel_value)
    return mean255.astype(np.float64), np.reciprocal(std255, dtype=np.float32)


def normalize_u8(images_hwc: torch.Tensor, out: Optional[torch.Tensor] = None, mean: Sequence[float] = IMAGENET_MEAN,
                 std: Sequence[float] = IMAGENET_STD) -> torch.Tensor:
    """images_hwc: (N,H,W,3) uint8 CUDA tensor, RGB.  Returns (N,3,H,W) float32."""
    if images_hwc.dim() != 4 or images_hwc.shape[-1] != 3 or images_hwc.dtype != torch.uint8:
        raise ValueError(f"images must be (N,H,W,3) uint8, got {tuple(images_hwc.shape)} {images_hwc.dtype}")
    if not images_hwc.is_cuda:
        raise RuntimeError("images are on the CPU: copy the uint8 batch to the GPU first (no CPU fallback)")
    images_hwc = images_hwc.contiguous()
    n, h, w, _ = images_hwc.shape
    if out is None:
        out = torch.empty((n, 3, h, w), dtype=torch.float32, device=images_hwc.device)
    elif tuple(out.shape) != (n, 3, h, w) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != images_hwc.device:
        raise ValueError("out must be a contiguous (N,3,H,W) float32 tensor on the images' device")
    mean255, inv = normalize_constants(mean, std)
    m = (C.c_double * 3)(*mean255.tolist())
    s = (C.c_float * 3)(*inv.tolist())
    st = _lib.load().cnl_normalize_images_u8(images_hwc.data_ptr(), out.data_ptr(), n, h, w, m, s,
                                             torch.cuda.current_stream(images_hwc.device).cuda_stream)
    _lib.check(st, "cnl_normalize_images_u8")
    return out
