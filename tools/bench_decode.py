"""Decode-only micro-benchmark.  Four input maps (4 x 168 MB > 126 MB L2) rotate; the four decodes are captured in one
CUDA graph so the CPU launch path is not what is measured; CUDA events bracket `iters` graph replays."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centernet_lightning_b200 import decode  # noqa: E402


def run(n, c, h, w, k, iters, logits, nbuf=4):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    bufs = [torch.randn((n, c, h, w), generator=g, device=dev) * 1.5 - 2.19 for _ in range(nbuf)]
    if not logits:
        bufs = [b.sigmoid_() for b in bufs]
    box = torch.randn((n, 4, h, w), generator=g, device=dev)
    out = decode.DecodeBuffers(n, h, w, k, 0, dev)
    kw = dict(num_detections=k, nms_kernel=3, normalize_boxes=False, box_log=False, box_multiplier=16.0, stride=4, from_logits=logits)

    def body():
        for b in bufs:
            decode.decode_into(out, b, box, None, **kw)
    body()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        body()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        body()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        graph.replay()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / (iters * nbuf)
    alg = n * (4 * c * h * w + 16 * k + 28 * k)
    return dict(shape=[n, c, h, w], k=k, from_logits=logits, us=ms * 1e3, alg_GBs=alg / ms / 1e6, img_per_s=n / ms * 1e3)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=50)
    a = ap.parse_args()
    for shape in [(32, 80, 128, 128), (8, 80, 256, 256), (16, 2, 128, 128), (1, 80, 128, 128)]:
        for logits in (True, False):
            print(json.dumps(run(*shape, 100, a.iters, logits)), flush=True)
