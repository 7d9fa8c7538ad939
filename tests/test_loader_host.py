"""CPU: host side of the folder-inference loader (file discovery, cv2 decode + resize into a caller buffer, the constants
handed to the normalisation kernel) against the preprocessing oracle.  The kernel itself is checked in test_forward_gpu.py."""
import numpy as np
import pytest

from centernet_lightning_b200 import inference, preprocess
from oracle import preprocess_np


def test_discover_sorts_and_filters(tmp_path):
    for n in ("b.png", "a.JPG", "c.txt", "d.jpeg"):
        (tmp_path / n).write_bytes(b"x")
    assert inference.discover(str(tmp_path)) == ["a.JPG", "b.png", "d.jpeg"]
    assert inference.discover(str(tmp_path), ["z.png"]) == ["z.png"]          # explicit names are taken as given
    with pytest.raises(FileNotFoundError):
        inference.discover(str(tmp_path / "missing"))


def test_load_resized_u8_equals_oracle_and_fills_caller_buffer(tmp_path):
    import cv2
    rng = np.random.default_rng(0)
    img = rng.integers(0, 255, (37, 53, 3), dtype=np.uint8)
    path = str(tmp_path / "x.png")
    cv2.imwrite(path, img)
    ref = preprocess_np.load_resized_u8(path, 64)
    got = inference.load_resized_u8(path, 64)
    assert got.dtype == np.uint8 and got.shape == (64, 64, 3) and np.array_equal(got, ref)
    assert np.array_equal(cv2.cvtColor(cv2.imread(path), cv2.COLOR_BGR2RGB), img[..., ::-1])   # imwrite took BGR: the loader returns RGB
    buf = np.zeros((2, 64, 64, 3), np.uint8)
    out = inference.load_resized_u8(path, 64, buf[1])
    assert out is buf[1] or np.shares_memory(out, buf) and np.array_equal(buf[1], ref) and not buf[0].any()
    with pytest.raises(FileNotFoundError):
        inference.load_resized_u8(str(tmp_path / "nope.png"), 64)


def test_normalize_constants_are_albumentations_arrays():
    mean255, inv = preprocess.normalize_constants()
    assert mean255.dtype == np.float64 and inv.dtype == np.float32
    # albumentations 1.x functional.normalize: float32 arrays scaled in float32 (the mean travels widened to float64)
    np.testing.assert_array_equal(mean255, (np.array(preprocess_np.MEAN, dtype=np.float32) * np.float32(255.0)).astype(np.float64))
    np.testing.assert_array_equal(inv, np.reciprocal(np.array(preprocess_np.STD, dtype=np.float32) * np.float32(255.0), dtype=np.float32))
    # the kernel's formula with these constants == the oracle, evaluated in numpy with the same roundings
    x = np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(3, axis=2)
    mine = (x.astype(np.float64) - mean255).astype(np.float32) * inv
    assert np.array_equal(mine, preprocess_np.normalize(x))


def test_cpu_tensors_are_rejected_without_fallback():
    import torch
    with pytest.raises(RuntimeError):
        preprocess.normalize_u8(torch.zeros((1, 4, 4, 3), dtype=torch.uint8))
    with pytest.raises(ValueError):
        preprocess.normalize_u8(torch.zeros((1, 4, 4, 3), dtype=torch.float32))
