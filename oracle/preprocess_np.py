"""Preprocessing oracle (TEST INFRASTRUCTURE ONLY): numpy restatement of the inference transform the reference
applies before the model - README.md:84-87 ``A.Compose([A.Resize(512, 512), A.Normalize(), ToTensorV2()])`` on the
RGB uint8 image read by ``InferenceDataset.__getitem__`` (centernet_lightning/datasets/inference.py:28-33).

albumentations is a third-party dependency that is absent from /root/reference and not installed here
(requirements.txt: ``albumentations``, no pinned version).  Its published algorithm, restated:
  * ``A.Resize(h, w)``: ``cv2.resize(img, (w, h), interpolation=cv2.INTER_LINEAR)`` on the uint8 image;
  * ``A.Normalize(mean=(0.485,0.456,0.406), std=(0.229,0.224,0.225), max_pixel_value=255)``:
        mean = np.array(mean, dtype=np.float32); mean *= max_pixel_value      (float32, albumentations 1.x
        std  = np.array(std,  dtype=np.float32); std  *= max_pixel_value       functional.normalize)
        denominator = np.reciprocal(std, dtype=np.float32)
        img = img.astype(np.float32); img -= mean; img *= denominator          (cv2.subtract / cv2.multiply for 3-channel
                                                                               images in 1.x: the same float32 results)
  * ``ToTensorV2``: HWC -> CHW.
PARITY STATUS: "parity unpinned" (no reference test or golden vector covers the transform, and the library is not
installed here to validate the restatement against; written from its published source)."""
from __future__ import annotations

import numpy as np

MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


def normalize(img_u8_hwc: np.ndarray, mean=MEAN, std=STD, max_pixel_value: float = 255.0) -> np.ndarray:
    mean = np.array(mean, dtype=np.float32)
    mean *= np.float32(max_pixel_value)
    std = np.array(std, dtype=np.float32)
    std *= np.float32(max_pixel_value)
    denominator = np.reciprocal(std, dtype=np.float32)
    img = img_u8_hwc.astype(np.float32)
    img -= mean
    img *= denominator
    return img


def to_chw(img_hwc: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(np.moveaxis(img_hwc, -1, -3))


def load_resized_u8(path: str, size: int) -> np.ndarray:
    """datasets/inference.py:28-29 + A.Resize: BGR file -> RGB uint8 (size,size,3)."""
    import cv2
    img = cv2.imread(path)
    if img is None:
        raise FileNotFoundError(path)
    img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
    return cv2.resize(img, (size, size), interpolation=cv2.INTER_LINEAR)
