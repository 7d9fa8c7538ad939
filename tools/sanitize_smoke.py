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
# row-rolling conv form (layer1 rows wider than 64 pixels) + 256-column decode tiles
x = torch.rand((1, 3, 32, 288), device=dev)
out = net.detect(x, use_graph=False)
torch.cuda.synchronize()
h = torch.randn((1, 6, 8, 256), device=dev)
b = torch.randn((1, 4, 8, 256), device=dev)
print(decode.decode_detections(h, b, num_detections=9, from_logits=True)["scores"][0, :3])
# simple neck: upsampling stores and transposed-conv phases
net2 = CenterNet(4, backbone="resnet18", neck="simple", neck_config={"upsample_channels": (64, 64, 64), "upsample_type": "conv_transpose", "deconv_kernel": 4},
                 head_config={"width": 64, "depth": 1}, num_detections=10).init_synthetic_(1).to(dev)
print({k: tuple(v.shape) for k, v in net2.detect(torch.rand((1, 3, 64, 96), device=dev), use_graph=False).items()})
net3 = CenterNet(4, backbone="resnet18", neck="simple", neck_config={"upsample_channels": (64, 64, 64)},
                 head_config={"width": 64, "depth": 1}, num_detections=10).init_synthetic_(1).to(dev)
print({k: tuple(v.shape) for k, v in net3.detect(torch.rand((1, 3, 64, 96), device=dev), use_graph=False).items()})
# MobileNetV2 trunk (3x3/2 stem, depthwise 3x3, ReLU6, Cout tiles of 192) + BiFPN / IDA necks (fuse kernel, separable convs)
net4 = CenterNet(4, backbone="mobilenet_v2", neck="bifpn", neck_config={"conv_type": "separable", "num_layers": 1},
                 head_config={"width": 64, "depth": 1}, num_detections=10).init_synthetic_(2).to(dev)
print({k: tuple(v.shape) for k, v in net4.detect(torch.rand((1, 3, 64, 96), device=dev), use_graph=False).items()})
net5 = CenterNet(4, backbone="resnet18", neck="ida", neck_config={"weighted_fusion": True},
                 head_config={"width": 64, "depth": 1}, num_detections=10).init_synthetic_(3).to(dev)
print({k: tuple(v.shape) for k, v in net5.detect(torch.rand((1, 3, 64, 64), device=dev), use_graph=False).items()})
# loader + tracker kernels
from centernet_lightning_b200 import preprocess  # noqa: E402
from centernet_lightning_b200.tracker import CostMatrices  # noqa: E402
import numpy as np  # noqa: E402
print(preprocess.normalize_u8(torch.randint(0, 255, (2, 17, 23, 3), dtype=torch.uint8, device=dev)).shape,
      preprocess.normalize_u8(torch.randint(0, 255, (2, 16, 32, 3), dtype=torch.uint8, device=dev)).shape)
rng = np.random.default_rng(0)
bx = rng.random((7, 4)); bx[:, 2:] += bx[:, :2]
r, c = CostMatrices(dev)(rng.standard_normal((7, 16)), rng.standard_normal((5, 16)), bx, bx[:5], giou=True)
print(r.shape, c.shape)
torch.cuda.synchronize()
print("sanitize smoke done")
