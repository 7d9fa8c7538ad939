"""Decode-only micro-benchmark.  Four input maps (4 x 168 MB > 126 MB L2) rotate; the four decodes are captured in one
CUDA graph so the CPU launch path is not what is measured; CUDA events bracket `iters` graph replays."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centernet_lightning_b200 import decode  # noqa: E402


def run(n, c, h, w, k, iters, logits, nbuf=4, smooth=False, source=None, peaks_only=False):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    bufs = [torch.randn((n, c, h, w), generator=g, device=dev) * 1.5 - 2.19 for _ in range(nbuf)]
    if smooth:          # spatially correlated maps (what a network emits): far fewer 3x3 peaks than white noise
        bufs = [torch.nn.functional.avg_pool2d(b, 7, 1, 3) * 5.0 + 6.5 for b in bufs]
    if not logits:
        bufs = [b.sigmoid_() for b in bufs]
    if source is not None:      # head maps of the synthetic network itself (what bench.py's decode_roofline times)
        bufs = [source.clone() for _ in range(nbuf)]
    box = torch.randn((n, 4, h, w), generator=g, device=dev)
    out = decode.DecodeBuffers(n, h, w, k, 0, dev)
    kw = dict(num_detections=k, nms_kernel=3, normalize_boxes=False, box_log=False, box_multiplier=16.0, stride=4, from_logits=logits)

    def body():
        for b in bufs:
            decode.decode_into(out, b, box, None, _peaks_only=peaks_only, **kw)
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
    return dict(shape=[n, c, h, w], k=k, from_logits=logits, smooth=smooth, peaks_only=peaks_only, us=ms * 1e3, alg_GBs=alg / ms / 1e6, img_per_s=n / ms * 1e3)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--net", action="store_true", help="also time the decode on the synthetic network's own heatmap")
    a = ap.parse_args()
    if a.net:
        from centernet_lightning_b200.model import CenterNet
        net = CenterNet(80, box_multiplier=16.0).init_synthetic_(0).to("cuda:0")
        heat = net.model(torch.rand((32, 3, 512, 512), device="cuda:0"))["heatmap"].clone()
        print("net heatmap: mean %.3f std %.3f max %.3f" % (heat.mean().item(), heat.std().item(), heat.max().item()), flush=True)
        r = run(32, 80, 128, 128, 100, a.iters, True, source=heat)
        r["source"] = "network"
        print(json.dumps(r), flush=True)
    for shape in [(32, 80, 128, 128), (8, 80, 256, 256), (16, 2, 128, 128), (1, 80, 128, 128)]:
        for logits in (True, False):
            print(json.dumps(run(*shape, 100, a.iters, logits)), flush=True)
            if shape[0] > 1:
                print(json.dumps(run(*shape, 100, a.iters, logits, peaks_only=True)), flush=True)
    print(json.dumps(run(32, 80, 128, 128, 100, a.iters, True, smooth=True)), flush=True)
