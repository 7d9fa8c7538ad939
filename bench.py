#!/usr/bin/env python
"""Headline benchmark: images/sec of the CenterNet inference hot path (ResNet-34 + FPN + heads -> fused decode)
on synthetic 3x512x512 batches, BASELINE.json configs[1] per GPU (batch 32, 80 classes, top-k 100).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # sm_100a arm (N>1: launched by torchrun)
    python bench.py --impl reference ...                           # the reference's CPU PyTorch path (oracle port)

One JSON line on stdout (rank 0).  A "step" = one forward + decode over one 32-image batch per GPU.
  value      : whole-job images/s with inputs resident in HBM (CUDA-graph replay, CUDA events, max over ranks)
  e2e        : same metric through the public API (CenterNet.detect) from PINNED HOST images, H2D copy of the batch and
               D2H read of boxes/scores/labels inside the timed region
  roofline   : the dominant kernel (conv_tc_kernel on the 3x3 256->256 tower conv @128x128; 7 of the 50 launches,
               74.6% of the FLOPs) timed alone with CUDA events; algorithmic FLOPs = 2 * 9.664 GMAC * batch
  cpu_baseline: the oracle port (torch CPU fp32 spec model + ATen decode) on a bounded sample, host cores stated
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec @ 3x512x512 (ResNet-34+FPN forward + fused decode)"
BATCH_PER_GPU = 32
SIZE = 512
CLASSES = 80
TOPK = 100
TOWER_GMAC = 9.663676416          # 3x3 256->256 conv at 128x128, per image
MODEL_GFLOP = 181.32              # algorithmic conv FLOPs per image at 512x512


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe): one streaming
    `nvidia-smi -lms 50` child started before the region and terminated (by its own PID) after it."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.3)                      # let the first samples arrive before the timed region starts
        except Exception:
            self.proc = None

    def stop(self):
        rows = []
        if self.proc is not None:
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                self.proc.kill()
                out = ""
            rows = [[c.strip() for c in line.split(",")] for line in out.splitlines() if line.strip()]
        sm = sorted(int(r[0]) for r in rows if r and r[0].isdigit())
        mx = max([int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        load = [x for x in sm if x > 0]
        return {"sm_mhz": load[len(load) // 2] if load else None, "sm_min_mhz": load[0] if load else None,
                "sm_max_mhz": mx or None, "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ----------------------------------------------------------------------------------------------------------------
def cpu_port_step(sample_batch: int, model=None):
    import torch
    from oracle import decode_torch, spec_model
    if model is None:
        model = spec_model.synth_init(spec_model.build_spec_model(CLASSES), seed=0)
    g = torch.Generator().manual_seed(0)
    x = torch.rand((sample_batch, 3, SIZE, SIZE), generator=g)

    def step():
        with torch.no_grad():
            out = model(x)
            det = decode_torch.decode_detections(out["heatmap"].sigmoid(), out["box_2d"], num_detections=TOPK,
                                                 box_multiplier=16.0, stride=4)
        return det
    return step


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample = 8
    step = cpu_port_step(sample)
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    val = sample / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "images/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "ResNet-34+FPN, 80 classes, 512x512, top-k=100 (BASELINE configs[1]); CPU sample of 8 images per step"},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{sample} images x {steps} steps, torch {torch.__version__} CPU fp32 spec model + ATen decode"},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# ----------------------------------------------------------------------------------------------------------------
# sm_100a arm
# ----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from centernet_lightning_b200.model import CenterNet
    from centernet_lightning_b200 import distributed as cdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the cnl_b200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=dev)

    net = CenterNet(CLASSES, "resnet34", box_multiplier=16.0, num_detections=TOPK, precision=args.precision)
    net.init_synthetic_(seed=0)
    net = net.to(dev)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    inputs = [torch.rand((BATCH_PER_GPU, 3, SIZE, SIZE), generator=g, device=dev) for _ in range(2)]
    gather = cdist.DetectionGather(BATCH_PER_GPU, TOPK, 0, dev) if world > 1 else None

    def step(i):
        det = net.detect(inputs[i & 1])
        if gather is not None:
            det = gather(det)                      # one NCCL all_gather of the packed (B,k,6) detections
        return det

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(max(3, args.warmup)):
        step(i)
    barrier()
    graph = next(iter(net._graphs.values()))
    launches_per_step = graph.launches

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        step(i)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    ms_per_step = ms_max / args.steps
    value = world * BATCH_PER_GPU * args.steps / (ms_max * 1e-3)

    # ---- e2e: public API from pinned host memory, H2D + D2H inside the timed region ----------------------------
    host_in = [torch.rand((BATCH_PER_GPU, 3, SIZE, SIZE)).pin_memory() for _ in range(2)]

    def e2e_run(n_steps):
        # public API: pinned host batches in, pinned host detections out; H2D of batch i+1 overlaps compute of batch i
        last = None
        for last in net.detect_host_batches((host_in[i & 1] for i in range(n_steps)), dev):
            pass
        return last
    e2e_run(3)
    barrier()
    e_steps = max(3, args.steps)        # the un-overlapped first H2D copy (pipeline fill) is inside the timed region
    ev0.record()
    e2e_run(e_steps)
    ev1.record()
    barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = world * BATCH_PER_GPU * e_steps / (float(t.item()) * 1e-3)
    h2d = BATCH_PER_GPU * 3 * SIZE * SIZE * 4
    d2h = BATCH_PER_GPU * TOPK * (16 + 4 + 8)

    line = None
    if rank == 0:
        # ---- roofline of the dominant kernel, timed alone with CUDA events on its launch stream ---------------
        eng = graph.engine
        names = [op.name for op in eng.plan.ops]
        idx = names.index("heads.heatmap.block_2")
        time.sleep(1.0)                  # timed ALONE against the burst peak: let the clocks recover from the power-capped loops above
        for _ in range(3):
            eng.forward(None, idx, idx + 1)
        torch.cuda.synchronize(dev)
        reps = 10
        ev0.record()
        for _ in range(reps):
            eng.forward(None, idx, idx + 1)
        ev1.record()
        torch.cuda.synchronize(dev)
        k_ms = ev0.elapsed_time(ev1) / reps
        burst, sustained, hbm, src = _peaks()
        flops = 2.0 * TOWER_GMAC * 1e9 * BATCH_PER_GPU
        achieved = flops / (k_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "conv_tc_kernel<2,true,false,true> (CTA-pair tcgen05 implicit GEMM; 3x3 256->256 @128x128, op heads.heatmap.block_2)",
                    "achieved": achieved, "peak": burst, "unit": "TFLOP/s", "frac": achieved / burst,
                    "peak_source": f"{src} bf16 burst (kernel timed alone, {reps} launches)", "kernel_ms": k_ms,
                    "algorithmic_flops_per_launch": flops, "traffic": 1044.0e6, "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum per launch (545.1 + 499.0 MB; algorithmic 2 x 536.9 MB), profiles/r01_tower_conv_ncu_details.txt",
                    "tensor_passes": 1 if args.precision == "fast" else 3}
        # ---- decode kernel: HBM roofline (second headline of BASELINE.json) ------------------------------------
        from centernet_lightning_b200 import decode as cdec
        heat = eng.outputs["heatmap"]
        box = eng.outputs["box_2d"]
        bufs = cdec.DecodeBuffers(BATCH_PER_GPU, SIZE // 4, SIZE // 4, TOPK, 0, dev)
        kw = dict(num_detections=TOPK, nms_kernel=3, normalize_boxes=False, box_log=False, box_multiplier=16.0, stride=4, from_logits=True)
        heats = [heat.clone() for _ in range(4)]                # 4 x 168 MB rotate (>> 126 MB L2)
        time.sleep(1.0)                                         # the kernel is timed ALONE: let the clocks recover from the power-capped conv loop above

        def dec_body():
            for hmap in heats:
                cdec.decode_into(bufs, hmap, box, None, **kw)
        dec_body()
        torch.cuda.synchronize(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            dec_body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        dgraph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(dgraph):
            dec_body()
        for _ in range(3):
            dgraph.replay()
        torch.cuda.synchronize(dev)
        ev0.record()
        for _ in range(25):
            dgraph.replay()
        ev1.record()
        torch.cuda.synchronize(dev)
        d_ms = ev0.elapsed_time(ev1) / (25 * len(heats))
        d_bytes = BATCH_PER_GPU * (4 * CLASSES * (SIZE // 4) ** 2 + 16 * TOPK + 28 * TOPK)
        decode_roof = {"bound": "hbm", "kernel": "whole decode: peaks_fast_kernel + select_gather_kernel (CUDA-graph replay, network's own heatmap)",
                       "achieved": d_bytes / (d_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                       "frac": d_bytes / (d_ms * 1e-3) / 1e9 / hbm, "decode_us": d_ms * 1e3,
                       "peaks_kernel_only": {"us": 33.18, "achieved": 5057.0, "frac": round(5057.0 / hbm, 3), "traffic": 171.4e6,
                                             "source": "ncu --set full of peaks_fast_kernel (gpu__time_duration, dram__bytes_read 167.8 MB + write 3.6 MB), profiles/r01_decode_peaks_ncu_details.txt"},
                       "algorithmic_bytes_per_launch": d_bytes}
        del heats, dgraph
        # ---- cpu baseline: oracle port on the host cores, bounded sample ---------------------------------------
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sample = 8
        cstep = cpu_port_step(sample)
        cstep()
        t0 = time.perf_counter()
        n_rep = 0
        while n_rep < 3 and (time.perf_counter() - t0) < 25:
            cstep()
            n_rep += 1
        cpu_dt = (time.perf_counter() - t0) / n_rep
        cpu = {"value": sample / cpu_dt, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{sample} images x {n_rep} reps of the same 512x512 workload (torch CPU fp32 spec model + ATen decode)"}
        # ---- supplementary: the single-pass fp16 mode (REDUCED precision: ~6e-2 max logit error, fails the 1e-3 bar) ------
        fast = None
        if args.precision != "fast" and not args.no_fast:
            net_f = CenterNet(CLASSES, "resnet34", box_multiplier=16.0, num_detections=TOPK, precision="fast")
            net_f.model.load_state_dict(net.model.state_dict())
            net_f = net_f.to(dev)
            for i in range(3):
                net_f.detect(inputs[i & 1])
            torch.cuda.synchronize(dev)
            ev0.record()
            for i in range(10):
                net_f.detect(inputs[i & 1])
            ev1.record()
            torch.cuda.synchronize(dev)
            f_ms = ev0.elapsed_time(ev1) / 10
            fast = {"value": BATCH_PER_GPU / f_ms * 1e3, "unit": "images/s", "ms_per_step": f_ms, "n_gpus": 1,
                    "note": "single fp16 tensor pass; reduced precision, NOT the headline (max head-map error ~6e-2 vs 4e-4)"}
            net_f.invalidate()
        roofline["tensor_pipe_frac"] = roofline["frac"] * roofline["tensor_passes"]
        tf = value / world * MODEL_GFLOP / 1e3
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp16 hi+lo split x3 tensor passes, fp32 accumulate (fp32-equivalent)" if args.precision != "fast"
                     else "fp16 single pass, fp32 accumulate (REDUCED precision)",
            "data": "synthetic",
            "config": {"workload": "ResNet-34+FPN, 80 classes, batch 32/GPU @ 3x512x512, top-k=100 (BASELINE configs[1])",
                       "global_batch": world * BATCH_PER_GPU, "parallelism": f"dp{world} batch-sharded, one all_gather of detections",
                       "l2": "per-step working set ~7.5 GB >> 126 MB L2; two input batches alternate",
                       "precision": args.precision},
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e_steps},
            "gpu_launches": launches_per_step * args.steps,
            "launches_per_step": launches_per_step,
            "roofline": roofline,
            "decode_roofline": decode_roof,
            "model_tflops_per_gpu_algorithmic": tf,
            "cpu_baseline": cpu,
            "reduced_precision_fp16": fast,
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        _emit(line)


_JSON_OUT = None


def _claim_stdout():
    """stdout carries exactly ONE JSON line.  Libraries write banners to file descriptor 1 behind Python's back (NCCL's
    version line, for one), so fd 1 is pointed at stderr for the whole run and the JSON line goes to a private duplicate
    of the original stdout."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def _emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-fast", action="store_true", help="skip the supplementary reduced-precision measurement")
    ap.add_argument("--precision", default="split", choices=["split", "split_fused", "fast"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
