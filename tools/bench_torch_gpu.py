"""The "library" bar of SURVEY 8d: the spec model (oracle/spec_model.py = the reference's graph) run by eager PyTorch /
cuDNN on the same B200, batch 32 @ 3x512x512, followed by the reference's own op sequence for the decode
(oracle/decode_torch.py on the GPU).  NOT the product path and not a parity check - context for the headline number.

    python tools/bench_torch_gpu.py
Modes: fp32 with TF32 disabled (what the 1e-3 bar is defined against), fp32 with TF32 tensor cores (PyTorch's conv
default), bf16 autocast; NCHW and channels_last."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import decode_torch, spec_model  # noqa: E402


def run(mode, channels_last, iters=10):
    dev = torch.device("cuda:0")
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = mode == "tf32"
    torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"
    m = spec_model.synth_init(spec_model.build_spec_model(80), seed=0).to(dev).eval()
    x = torch.rand((32, 3, 512, 512), device=dev)
    if channels_last:
        m = m.to(memory_format=torch.channels_last)
        x = x.contiguous(memory_format=torch.channels_last)

    def step():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
            out = m(x)
        return decode_torch.decode_detections(out["heatmap"].float().sigmoid(), out["box_2d"].float(), num_detections=100,
                                              box_multiplier=16.0, stride=4)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        step()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    return {"mode": mode, "channels_last": channels_last, "ms_per_step": ms, "images_per_s": 32 / ms * 1e3}


if __name__ == "__main__":
    for mode in ("fp32", "tf32", "bf16"):
        for cl in (False, True):
            print(json.dumps(run(mode, cl)), flush=True)
