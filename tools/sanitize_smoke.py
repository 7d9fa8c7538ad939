"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): tiny model forward + decode + generic decode."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centernet_lightning_b200.model import CenterNet  # noqa: E402
from centernet_lightning_b200 import decode  # noqa: E402

dev = torch.device("cuda:0")
net = CenterNet(8, backbone="resnet18", neck_config={"out_channels": 64}, head_config={"width": 64, "depth": 1},
                box_multiplier=16.0, num_detections=20).init_synthetic_(0).to(dev)
x = torch.rand((2, 3, 96, 160), device=dev)
out = net.detect(x, use_graph=False)
torch.cuda.synchronize()
print({k: tuple(v.shape) for k, v in out.items()})
h = torch.randn((2, 5, 17, 23), device=dev)
b = torch.randn((2, 4, 17, 23), device=dev)
print(decode.decode_detections(h, b, num_detections=7, from_logits=True)["scores"][0, :3])
h = torch.randn((2, 8, 128, 128), device=dev)
b = torch.randn((2, 4, 128, 128), device=dev)
print(decode.decode_detections(h, b, num_detections=50, from_logits=True)["scores"][0, :3])
torch.cuda.synchronize()
print("sanitize smoke done")
