"""Every GPU kernel of one multi-GPU bench step, by name (torch.profiler / CUPTI activity records on rank 0):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/step_kernels.py

VERDICT r1 item 6: the step must launch only this package's kernels plus ONE ncclDevKernel_AllGather - no ATen kernel
around the collective (the select kernel writes the packed rows into the all_gather's send buffer itself and the gathered
buffer is consumed through views)."""
import collections
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centernet_lightning_b200 import distributed as cdist  # noqa: E402
from centernet_lightning_b200.model import CenterNet  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, K = 32, 100
    net = CenterNet(80, "resnet34", box_multiplier=16.0, num_detections=K).init_synthetic_(0).to(dev)
    inputs = [torch.rand((B, 3, 512, 512), device=dev) for _ in range(2)]
    gather = cdist.DetectionGather(B, K, 0, dev) if world > 1 else None

    def step(i):
        det = net.detect(inputs[i & 1], static_input=True, packed_out=gather.local if gather is not None else None)
        return gather() if gather is not None else det

    for i in range(4):
        step(i)
    torch.cuda.synchronize(dev)
    steps = 3
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for i in range(steps):
            step(i)
        torch.cuda.synchronize(dev)
    if rank == 0:
        names = collections.Counter()
        for ev in prof.events():
            if ev.device_type == torch.autograd.DeviceType.CUDA:
                names[ev.name] += 1
        print(f"# kernels / memcpys of {steps} bench steps on rank 0 of {world} (batch {B}/GPU, 512x512, k={K}); count per step in brackets")
        for n, c in sorted(names.items(), key=lambda kv: -kv[1]):
            print(f"{c:5d} [{c / steps:5.1f}]  {n[:160]}")
        foreign = [n for n in names if not (n.startswith("void cnl::") or n.startswith("cnl::") or "ncclDevKernel" in n)]
        print("# kernels that are neither cnl:: nor NCCL:", foreign if foreign else "none")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
