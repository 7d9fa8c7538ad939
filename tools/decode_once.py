"""One decode of the headline shape per mode (for ncu): python tools/decode_once.py [logits|probs]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centernet_lightning_b200 import decode  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "logits"
n, c, h, w, k = (int(x) for x in (sys.argv[2:7] if len(sys.argv) > 6 else (32, 80, 128, 128, 100)))
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
bufs = [torch.randn((n, c, h, w), generator=g, device=dev) * 1.5 - 2.19 for _ in range(4)]
if mode == "probs":
    bufs = [b.sigmoid_() for b in bufs]
box = torch.randn((n, 4, h, w), generator=g, device=dev)
out = decode.DecodeBuffers(n, h, w, k, 0, dev)
kw = dict(num_detections=k, nms_kernel=3, normalize_boxes=False, box_log=False, box_multiplier=16.0, stride=4, from_logits=(mode == "logits"))
for _ in range(3):
    for b in bufs:
        decode.decode_into(out, b, box, None, **kw)
torch.cuda.synchronize()
