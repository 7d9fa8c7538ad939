"""Folder-inference benchmark: images/s of CenterNet.inference_detection on a folder of synthetic
JPEGs, next to (a) the same loader work done serially on one host core with float32 batches (what the reference's
InferenceDataset + albumentations pipeline does per item) and (b) the model-only rate of the same batches.

    python tools/bench_inference.py [--n 512] [--batch 32] [--src 640x480] [--workers W]
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centernet_lightning_b200.model import CenterNet  # noqa: E402
from centernet_lightning_b200 import inference  # noqa: E402


def main():
    import cv2
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--src", default="640x480")
    ap.add_argument("--workers", type=int, default=0)
    a = ap.parse_args()
    sw, sh = (int(v) for v in a.src.split("x"))
    rng = np.random.default_rng(0)
    d = tempfile.mkdtemp(prefix="cnl_imgs_")
    base = cv2.GaussianBlur(rng.integers(0, 255, (sh, sw, 3), dtype=np.uint8), (0, 0), 3)       # photo-like JPEG entropy
    for i in range(a.n):
        cv2.imwrite(os.path.join(d, f"{i:05d}.jpg"), np.roll(base, 7 * i, axis=1), [cv2.IMWRITE_JPEG_QUALITY, 90])
    dev = torch.device("cuda:0")
    net = CenterNet(80, box_multiplier=16.0).init_synthetic_(0).to(dev)
    workers = a.workers or min(32, os.cpu_count() or 4)
    net.inference_detection(d, img_names=sorted(os.listdir(d))[:2 * a.batch], batch_size=a.batch, workers=workers)      # warm-up: engine + graph
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = net.inference_detection(d, batch_size=a.batch, workers=workers)
    dt = time.perf_counter() - t0
    assert out["bboxes"].shape == (a.n, 100, 4)
    # (a) serial single-core loader of the reference's shape: decode + resize + float normalise per item
    from oracle import preprocess_np
    names = sorted(os.listdir(d))[:64]
    t1 = time.perf_counter()
    for n in names:
        preprocess_np.to_chw(preprocess_np.normalize(preprocess_np.load_resized_u8(os.path.join(d, n), 512)))
    serial = len(names) / (time.perf_counter() - t1)
    # (b) model only on resident batches
    x = torch.rand((a.batch, 3, 512, 512), device=dev)
    for _ in range(3):
        net.detect(x)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    for _ in range(10):
        net.detect(x)
    torch.cuda.synchronize()
    model_only = 10 * a.batch / (time.perf_counter() - t2)
    print(json.dumps({"folder_images_per_s": a.n / dt, "n_images": a.n, "batch": a.batch, "source": a.src + " JPEG q90",
                      "loader_threads": workers, "host_cores": os.cpu_count(),
                      "serial_cpu_loader_images_per_s_one_core": serial, "model_only_images_per_s": model_only,
                      "h2d_bytes_per_image": 512 * 512 * 3}), flush=True)
    for f in os.listdir(d):
        os.remove(os.path.join(d, f))
    os.rmdir(d)


if __name__ == "__main__":
    main()
