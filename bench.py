#!/usr/bin/env python
"""Benchmark of the B200 hot path: bf16 training of the heatmap CNN (BASELINE.json configs[1]:
"same CNN, bf16 training batch 64 on 1xB200, random init"), data-parallel over N GPUs.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                      (CPU arm: the oracle port on the host cores)

One "step" = forward + MSE loss (Gaussian targets rendered on device from keypoints) + backward
+ bucketed gradient all-reduce (N>1) + fused Adam, batch 64 PER GPU (weak scaling).  Prints ONE JSON
line on rank 0.  See DESIGN.md "measurement" for how every field is produced.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

JOINTS = 36
BATCH_PER_GPU = 64
IMG = 192
CFG = {"model type": "MODEL_18_POINTS_PER_WING", "number of base filters": 64, "convolution kernel size": 3,
       "dilation rate": 2, "dropout ratio": 0.5, "precision": "bf16"}
TRAIN_GFLOP_PER_SAMPLE = 80.09   # BASELINE.md section 2 (3 x fwd - conv1 dgrad), C=36
FWD_GFLOP_PER_SAMPLE = 26.754
VIT_CFG = dict(CFG, **{"model type": "MODEL_18_POINTS_PER_WING_VIT", "optimizer": "adam", "patch size": 16,
                       "projection dim": 256, "num heads": 12, "transformer layers": 8, "dim head": -1})
VIT_TRAIN_GFLOP_PER_SAMPLE = 46.92
VIT_FWD_GFLOP_PER_SAMPLE = 15.666
# FourCamerasBaseLine (pytorch/CNNs.py:189-237), 4 views x (192,192,4), 72 heatmaps: dense-equivalent MACs per 4-view
# sample = 4 x 9 598 M (shared encoder) + 2 416 M (1x1 mixing conv) + 4 x (16 987 + 2 x 33 974 + 955) M (decoder on
# 1280 channels) = 384.4 GMAC forward; training = 3 x forward - the first layer's input gradient
FOURCAM_CFG = dict(CFG, **{"model type": "ALL_CAMS_18_POINTS"})
FOURCAM_FWD_GFLOP_PER_SAMPLE = 768.74
FOURCAM_TRAIN_GFLOP_PER_SAMPLE = 3 * 768.74 - 4 * 2 * 0.0849


def _peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        d["_source"] = "measured"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "_source": "fallback"}


class ClockSampler(threading.Thread):
    """samples SM clock / throttle reasons through NVML every 20 ms while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
                 "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80)}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop_evt.wait(0.02)

    def stop(self) -> dict:
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------ CPU arms
def cpu_train_step_seconds(batch: int, steps: int, warmup: int, model: str = "cnn"):
    """the oracle port (oracle/pose_oracle.py: the reference's modules restated on torch CPU ops),
    forward + MSE + backward + Adam on the host cores; returns (seconds per step list, cores)."""
    from oracle import pose_oracle as po  # checker / baseline only
    torch.set_num_threads(os.cpu_count() or 1)
    if model == "vit":
        sd, forward, cin = po.vit_state_dict(JOINTS, seed=0), po.vit_forward, 4
    elif model == "fourcam":
        sd, forward, cin = po.four_cameras_state_dict(JOINTS, seed=0), po.four_cameras_baseline_forward, 16
    else:
        sd, forward, cin = po.basicnet_state_dict(JOINTS, seed=0), po.basicnet_forward, 4
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point()}
    opt = torch.optim.Adam([v for k, v in params.items() if not k.endswith("cls_token")], lr=1e-3)
    x = po.synthetic_crops(batch, seed=1, cin=cin)
    tgt = torch.from_numpy(po.gaussian_targets(po.synthetic_points(batch, JOINTS, seed=2)))
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss = po.mse_loss(forward(params, x), tgt)
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times, torch.get_num_threads()


def cpu_inference_frames_per_sec(batch: int, reps: int, model: str = "cnn") -> float:
    """forward + per-joint argmax peaks of the oracle port on the host cores (pytorch/train_pytorch.py:155-170,
    199-213 without the plotting): frames per second, best of `reps` after one warm-up."""
    from oracle import pose_oracle as po  # checker / baseline only
    torch.set_num_threads(os.cpu_count() or 1)
    if model == "vit":
        sd, forward, cin = po.vit_state_dict(JOINTS, seed=0), po.vit_forward, 4
    elif model == "fourcam":
        sd, forward, cin = po.four_cameras_state_dict(JOINTS, seed=0), po.four_cameras_baseline_forward, 16
    else:
        sd, forward, cin = po.basicnet_state_dict(JOINTS, seed=0), po.basicnet_forward, 4
    x = po.synthetic_crops(batch, seed=3, cin=cin)
    best = float("inf")
    with torch.no_grad():
        for i in range(reps + 1):
            t0 = time.perf_counter()
            out = forward(sd, x)
            po.find_peaks_argmax(out.permute(0, 2, 3, 1).contiguous())
            dt = time.perf_counter() - t0
            if i > 0:
                best = min(best, dt)
    return batch / best


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 2 if args.model == "fourcam" else 8      # a four-view sample is ~29 BasicNet samples of arithmetic
    times, cores = cpu_train_step_seconds(sample, args.steps, args.warmup, args.model)
    ms = 1e3 * float(np.mean(times))
    value = sample / (ms / 1e3)
    line = {
        "impl": "reference", "metric": "train_samples_per_sec", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": {"cnn": "BasicNet", "vit": "VIT_encoder_CNN_decoder", "fourcam": "FourCamerasBaseLine"}[
                       args.model] + f" C={JOINTS} training step (fwd + MSE + bwd + Adam), 192x192x"
                       f"{16 if args.model == 'fourcam' else 4} crops",
                   "batch_per_step": sample, "note": "reference's own CPU PyTorch path restated in oracle/ "
                   "(the Python reference cannot travel to the GPU box); each step is a bounded 8-sample "
                   "slice of the workload"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps of batch {sample} after {args.warmup} warm-up"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if not args.no_inference:
        line["inference"] = {"metric": "inference_frames_per_sec", "unit": "frames/s",
                             "value": cpu_inference_frames_per_sec(sample, 2, args.model),
                             "sample": f"forward + argmax peaks of {sample} frames, best of 2 after 1 warm-up"}
    print(json.dumps(line), flush=True)



# ------------------------------------------------------------------------------------------ HBM-bound kernels
def bandwidth_kernels(dev, hbm_gbs: float, iters: int = 5) -> list:
    """CUDA-event time of every heatmap / loss / peak kernel at the bench workload size, as ALGORITHMIC bytes
     / time against the measured HBM copy bandwidth.  Every operand set is larger than the
    126 MB L2, so no iteration finds its input cached."""
    from pose_estimation_amitai_b200 import ops
    B, C, H, W, CP = BATCH_PER_GPU, JOINTS, IMG, IMG, 48
    E = B * C * H * W
    out = torch.rand(B, C, H, W, device=dev) - 0.3
    pts = torch.randint(8, IMG - 8, (B, C, 2), device=dev).float()
    tgt = ops.gaussian_heatmaps(pts)
    IB = 256
    hm = torch.rand(IB, C, H, W, device=dev)
    hm_bf = hm.to(torch.bfloat16)
    x64 = (torch.rand(B, H, W, 64, device=dev) - 0.5).to(torch.bfloat16)
    gy64 = (torch.rand(B, H // 2, W // 2, 64, device=dev) - 0.5).to(torch.bfloat16)
    mask64 = torch.randint(-2 ** 31, 2 ** 31 - 1, (B * H * W, 2), device=dev, dtype=torch.int64).to(torch.int32)
    rs = np.random.RandomState(0)
    th = np.array([[math.cos(a), math.sin(a), tx, -math.sin(a), math.cos(a), ty] for a, tx, ty in
                   zip(np.radians(rs.uniform(-30, 30, IB)), rs.uniform(-10, 10, IB), rs.uniform(-10, 10, IB))], np.float32)
    aff_theta = torch.from_numpy(th).to(dev)
    aff_flips = torch.from_numpy(rs.randint(0, 4, IB).astype(np.int32)).to(dev)
    aff_src = torch.from_numpy(rs.permutation(IB).astype(np.int32)).to(dev)
    box_u8 = torch.randint(0, 256, (IB, 4, H, W), device=dev, dtype=torch.uint8)
    aff_out, aff_out4 = torch.empty(B, C, H, W, device=dev), torch.empty(IB, 4, H, W, device=dev)
    cases = [
        ("affine_nearest_kernel: batch gather + rotate/shift/flip of 64 x 36 confidence maps (fp32 -> fp32)", 2 * E * 4,
         lambda: ops.affine_nearest(hm, aff_theta[:B], aff_flips[:B], src_index=aff_src[:B], out=aff_out)),
        ("affine_nearest_kernel: batch gather + ToTensor + rotate/shift/flip of 256 x 4 uint8 crops (u8 -> fp32)",
         IB * 4 * H * W * 5, lambda: ops.affine_nearest(box_u8, aff_theta, aff_flips, src_index=aff_src, out=aff_out4)),
        ("mse_nhwc_bf16_kernel: MSE + grad (bf16 NHWC), Gaussian target fused", E * 4 + B * H * W * CP * 2 + 8 * B * C,
         lambda: ops.mse_loss_fwd_bwd(out, None, points=pts, grad_nhwc_dtype=torch.bfloat16, cpad=CP)),
        ("mse_nhwc_bf16_kernel: MSE + grad (bf16 NHWC), fp32 target read", 2 * E * 4 + B * H * W * CP * 2,
         lambda: ops.mse_loss_fwd_bwd(out, tgt, grad_nhwc_dtype=torch.bfloat16, cpad=CP)),
        ("mse_kernel: MSE + fp32 NCHW grad (autograd path)", 3 * E * 4,
         lambda: ops.mse_loss_fwd_bwd(out, tgt, want_grad_nchw=True)),
        ("gaussian_kernel: sigma=3 targets from keypoints", E * 4 + 8 * B * C, lambda: ops.gaussian_heatmaps(pts)),
        ("argmax_planar_kernel: peaks of 256 frames, fp32 NCHW", IB * C * H * W * 4 + 8 * IB * C,
         lambda: ops.peaks_argmax(hm)),
        ("argmax_planar_kernel: peaks of 256 frames, bf16 NCHW", IB * C * H * W * 2 + 8 * IB * C,
         lambda: ops.peaks_argmax(hm_bf)),
        ("softargmax: 256 frames, fp32 NCHW", IB * C * H * W * 4 + 8 * IB * C, lambda: ops.peaks_softargmax(hm)),
        ("pool_fwd_vec_kernel: 2x2 maxpool + LeakyReLU, 192^2 x 64", x64.numel() * 2 * 5 // 4,
         lambda: ops.maxpool_lrelu_fwd(x64)),
        ("pool_bwd_vec_kernel: maxpool backward (g and masked g), 192^2 x 64",
         x64.numel() * 2 * 3 + gy64.numel() * 2 + mask64.numel() * 4, lambda: ops.maxpool_lrelu_bwd(x64, gy64, mask64)),
    ]
    res = []
    for name, nbytes, fn in cases:
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / iters * 1e3
        gbs = nbytes / (us * 1e-6) / 1e9
        res.append({"kernel": name, "bytes": int(nbytes), "us": round(us, 1), "gbs": round(gbs, 1),
                    "frac_of_hbm_peak": round(gbs / hbm_gbs, 3)})
    return res

# ------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args) -> None:
    import torch.distributed as dist
    from pose_estimation_amitai_b200 import CNNs, ops, parallel

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert world == args.gpus or world == 1 and args.gpus == 1, "--gpus must equal WORLD_SIZE"
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)

    torch.manual_seed(0)  # same random init on every rank
    global TRAIN_GFLOP_PER_SAMPLE, FWD_GFLOP_PER_SAMPLE
    if args.model == "vit":   # BASELINE.json configs[3]: parity-test configuration, measured on request
        from pose_estimation_amitai_b200 import VITs
        model = VITs.VIT_encoder_CNN_decoder(dict(VIT_CFG), np.array((IMG, IMG, 4)), JOINTS).to(dev)
        TRAIN_GFLOP_PER_SAMPLE, FWD_GFLOP_PER_SAMPLE = VIT_TRAIN_GFLOP_PER_SAMPLE, VIT_FWD_GFLOP_PER_SAMPLE
    elif args.model == "fourcam":   # the multi-camera baseline, measured on request
        model = CNNs.FourCamerasBaseLine(dict(FOURCAM_CFG), np.array((IMG, IMG, 16)), JOINTS).to(dev)
        TRAIN_GFLOP_PER_SAMPLE, FWD_GFLOP_PER_SAMPLE = FOURCAM_TRAIN_GFLOP_PER_SAMPLE, FOURCAM_FWD_GFLOP_PER_SAMPLE
    else:
        model = CNNs.BasicNet(dict(CFG), np.array((IMG, IMG, 4)), JOINTS).to(dev)
    dp = parallel.DataParallelStep(model, lr=1e-3)
    CIN = 16 if args.model == "fourcam" else 4

    B = args.batch_per_gpu if args.batch_per_gpu > 0 else (16 if args.model == "fourcam" else BATCH_PER_GPU)
    g = torch.Generator().manual_seed(1 + rank)
    x_host = torch.rand(B, CIN, IMG, IMG, generator=g).pin_memory()
    pts_host = torch.randint(8, IMG - 8, (B, JOINTS, 2), generator=torch.Generator().manual_seed(2 + rank)
                             ).float().pin_memory()
    x_dev, pts_dev = x_host.to(dev), pts_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    # ---- resident-input step ---------------------------------------------------------------
    def step_resident(_i):
        dp.step(x_dev, points=pts_dev)

    for i in range(max(args.warmup, 3)):
        step_resident(i)
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ops.launch_count()
    ms_total = timed(step_resident, args.steps)
    launches = ops.launch_count() - launches0
    clocks = sampler.stop()
    ms_step = ms_total / args.steps
    value = world * B / (ms_step / 1e3)

    # ---- end-to-end step: host (pinned) inputs -> device, result scalar -> host --------------
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [(torch.empty_like(x_dev), torch.empty_like(pts_dev)) for _ in range(2)]
    loss_host = torch.zeros(1).pin_memory()
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    probe = os.environ.get("POSEB200_E2E_PROBE", "")   # timing experiments only: "noh2d", "nod2h"

    def issue_copy(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            if probe != "noh2d":
                bufs[slot][0].copy_(x_host, non_blocking=True)
                bufs[slot][1].copy_(pts_host, non_blocking=True)
            ready[slot].record(copy_stream)

    def step_e2e(i):
        slot = i & 1
        if i == 0:
            issue_copy(0)
        issue_copy(slot ^ 1)  # prefetch the next step's inputs while this step computes
        torch.cuda.current_stream().wait_event(ready[slot])
        loss = dp.step(bufs[slot][0], points=bufs[slot][1])
        consumed[slot].record(torch.cuda.current_stream())
        if probe != "nod2h":
            loss_host.copy_(loss, non_blocking=True)

    for s in range(2):
        consumed[s].record(torch.cuda.current_stream())
    for i in range(3):
        step_e2e(i)
    barrier()
    for s in range(2):
        consumed[s].record(torch.cuda.current_stream())
    ms_e2e = timed(step_e2e, args.steps) / args.steps
    e2e_value = world * B / (ms_e2e / 1e3)
    h2d = x_host.numel() * 4 + pts_host.numel() * 4

    line = {
        "metric": "train_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong" if args.batch_per_gpu > 0 and args.batch_per_gpu * world == BATCH_PER_GPU else "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": {"cnn": "BasicNet (pytorch/CNNs.py)", "vit": "VIT_encoder_CNN_decoder (pytorch/VITs.py)",
                                "fourcam": "FourCamerasBaseLine (pytorch/CNNs.py:189-237, 4 views per sample)"}[args.model] +
                               f" C={JOINTS} bf16 training step: fwd + MSE(Gaussian sigma=3 "
                               "targets rendered on device from keypoints) + bwd + grad all-reduce + fused Adam",
                   "batch_per_gpu": B, "global_batch": B * world, "image": [IMG, IMG, CIN], "joints": JOINTS,
                   "parallelism": f"dp{world}", "l2": "per-step working set (~5 GB of activations) >> 126 MB L2",
                   "grad_buckets_bytes": dp.buckets.bucket_sizes_bytes()},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "samples/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4,
                "note": "pinned host crops + keypoints copied every step (double-buffered on a copy stream), "
                        "loss scalar read back every step"},
        "gpu_launches": int(launches),
    }

    # ---- inference sweep point (BASELINE.json configs[4]): frame-sharded forward + on-device argmax peaks,
    #      no collective; frames/s over all ranks.  Resident inputs, then host-pinned inputs -> peaks on the host.
    inf = None
    if not args.no_inference:
        IB = args.infer_batch
        xi_host = torch.rand(IB, CIN, IMG, IMG, generator=torch.Generator().manual_seed(11 + rank)).pin_memory()
        xi_dev = xi_host.to(dev)
        peaks_host = torch.empty(IB, JOINTS, 2).pin_memory()

        def infer_resident(_i):
            model.predict_peaks(xi_dev)

        xbuf = [torch.empty_like(xi_dev) for _ in range(2)]
        iready = [torch.cuda.Event() for _ in range(2)]
        idone = [torch.cuda.Event() for _ in range(2)]

        def infer_copy(slot):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(idone[slot])
                xbuf[slot].copy_(xi_host, non_blocking=True)
                iready[slot].record(copy_stream)

        def infer_e2e(i):
            slot = i & 1
            if i == 0:
                infer_copy(0)
            infer_copy(slot ^ 1)
            torch.cuda.current_stream().wait_event(iready[slot])
            pk = model.predict_peaks(xbuf[slot])
            idone[slot].record(torch.cuda.current_stream())
            peaks_host.copy_(pk, non_blocking=True)

        isteps = max(3, args.steps // 2)
        for i in range(3):
            infer_resident(i)
        ms_inf = timed(infer_resident, isteps) / isteps
        for sl in range(2):
            idone[sl].record(torch.cuda.current_stream())
        for i in range(2):
            infer_e2e(i)
        barrier()
        for sl in range(2):
            idone[sl].record(torch.cuda.current_stream())
        ms_inf_e2e = timed(infer_e2e, isteps) / isteps
        inf = {"metric": "inference_frames_per_sec", "value": world * IB / (ms_inf / 1e3), "unit": "frames/s",
               "frames_per_gpu_per_step": IB, "ms_per_step": ms_inf, "steps": isteps,
               "e2e": {"value": world * IB / (ms_inf_e2e / 1e3), "unit": "frames/s", "ms_per_step": ms_inf_e2e,
                       "h2d_bytes_per_step": xi_host.numel() * 4, "d2h_bytes_per_step": peaks_host.numel() * 4},
               "fwd_tflops": world * IB / (ms_inf / 1e3) * FWD_GFLOP_PER_SAMPLE / 1e3,
               "workload": f"{type(model).__name__} C={JOINTS} bf16 forward + per-joint argmax peaks on device, "
                           "frame-sharded, no collective"}
        line["inference"] = inf
        del xi_dev, xbuf
        torch.cuda.empty_cache()

    # ---- roofline of the dominant kernel family (tcgen05 contractions), timed live: every rank runs the step
    #      (it contains the gradient all-reduce), rank 0 reports
    ops.profile_begin()
    dp.step(x_dev, points=pts_dev)
    rec = ops.profile_end()
    barrier()
    if rank == 0:
        peaks = _peaks()
        by = {}
        for name, flops, ms in rec:
            a = by.setdefault(name, [0.0, 0.0, 0])
            a[0] += flops; a[1] += ms; a[2] += 1
        dom = max(by.items(), key=lambda kv: kv[1][1])
        dname, (dflops, dms, dcount) = dom
        achieved = dflops / (dms * 1e-3) / 1e12
        peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
        line["roofline"] = {
            "kernel": {"pb_conv_tc": "tc_conv2_kernel (+tc_conv_kernel)", "pb_wgrad_tc": "tc_wgrad2_kernel (+tc_wgrad_kernel)",
                       "pb_conv_simt": "conv_simt_kernel", "pb_wgrad_simt": "wgrad_simt_kernel"}.get(dname, dname),
            "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            # dram__bytes_read.sum + dram__bytes_write.sum of this kernel's heaviest launch (conv2/conv3 forward, the
            # 64->64 layer at 192^2; algorithmic: input 302 MB + residual 302 MB + output 302 MB + sign mask 19 MB)
            # from the ncu --set full capture profiles/r1g_conv2fwd_pair_ncu_summary.txt; the other captures are listed
            "traffic": 894.0e6 if args.model == "cnn" else None,
            "traffic_detail": {"unit": "bytes per launch, ncu --set full, batch 64",
                               "conv2 fwd (64->64 @192^2)": 894.0e6, "conv2 dgrad (two outputs)": 1182.2e6,
                               "conv5 fwd (128->128 @96^2)": 431.3e6, "conv8 fwd (256->256 @48^2)": 201.4e6,
                               "conv5 wgrad (tc_wgrad2_kernel)": 537.8e6,
                               "tensor_pipe_active_pct": {"conv8 fwd": 84.5, "conv5 fwd": 69.8, "conv5 wgrad": 72.5,
                                                          "conv2 fwd": 41.2, "conv2 dgrad": 31.8},
                               "source": "profiles/r1g_*_ncu_summary.txt"},
            "launches_per_step": dcount, "avg_launch_ms": dms / dcount,
            "share_of_step": dms / ms_step,
            "peak_source": f"{peaks['_source']} bf16_tflops_sustained (kernel timed inside a long step)",
            "per_family": {k: {"tflops": v[0] / (v[1] * 1e-3) / 1e12, "ms": v[1], "launches": v[2]}
                           for k, v in by.items()},
            "step_tflops": value / world * TRAIN_GFLOP_PER_SAMPLE / 1e3,
            "step_frac_of_peak": value / world * TRAIN_GFLOP_PER_SAMPLE / 1e3 / peak,
        }
        if not args.no_bandwidth:
            del x_dev
            torch.cuda.empty_cache()
            line["bandwidth_kernels"] = {"peak_gbs": peaks["hbm_gbs"], "peak_source": peaks["_source"] + " hbm_gbs",
                                         "note": "timings include the launch of small helper kernels each op needs "
                                                 "(loss zeroing, key finalize)",
                                         "kernels": bandwidth_kernels(dev, peaks["hbm_gbs"])}
        # ---- CPU baseline on this box's host cores (bounded sample) -----------------------------
        if world == 1 and not args.no_cpu_baseline:
            cb = 2 if args.model == "fourcam" else 8
            times, cores = cpu_train_step_seconds(cb, 3, 1, args.model)
            line["cpu_baseline"] = {"value": cb / float(np.mean(times)), "unit": "samples/s", "cores": cores,
                                    "kind": "port", "sample": f"3 steps of batch {cb} (fwd+MSE+bwd+Adam) after 1 "
                                    "warm-up, oracle/pose_oracle.py on torch CPU fp32"}
            if not args.no_inference:
                line["cpu_baseline"]["inference_frames_per_sec"] = cpu_inference_frames_per_sec(cb, 2, args.model)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-inference", action="store_true")
    ap.add_argument("--no-bandwidth", action="store_true")
    ap.add_argument("--model", default="cnn", choices=["cnn", "vit", "fourcam"])
    ap.add_argument("--infer-batch", type=int, default=256)
    ap.add_argument("--joints", type=int, default=JOINTS,
                    help="output heatmaps: 36 (named config) or 18 (the reference's own per-wing data, SURVEY.md 8)")
    ap.add_argument("--batch-per-gpu", type=int, default=0,
                    help="override the 64 samples per GPU of the named config (64 / N gives the strong-scaling point)")
    args = ap.parse_args()
    if args.joints != JOINTS:
        # dense-equivalent FLOPs of the head scale with C (26.372 / 26.754 GFLOP forward at C = 18 / 36)
        globals()["FWD_GFLOP_PER_SAMPLE"] = 26.372 + (26.754 - 26.372) * (args.joints - 18) / 18.0
        globals()["TRAIN_GFLOP_PER_SAMPLE"] = 3 * globals()["FWD_GFLOP_PER_SAMPLE"] - 2 * 0.0849
        globals()["JOINTS"] = args.joints
    if args.model == "fourcam":
        globals()["JOINTS"] = 72 if args.joints == 36 else args.joints
        if args.infer_batch == 256:
            args.infer_batch = 32
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
