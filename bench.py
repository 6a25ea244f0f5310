#!/usr/bin/env python
"""Benchmark of the B200 hot path: bf16 training of the heatmap CNN (BASELINE.json configs[1]:
"same CNN, bf16 training batch 64 on 1xB200, random init"), data-parallel over N GPUs.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                      (CPU arm: the oracle port on the host cores)

One "step" = forward + MSE loss (Gaussian targets rendered on device from keypoints) + backward
+ bucketed gradient all-reduce (N>1) + fused Adam, batch 64 PER GPU (weak scaling).  Prints ONE JSON
line on rank 0.  See DESIGN.md "measurement" for how every field is produced.

Besides the headline (BASELINE.json configs[1] / [2]) the same line carries, each measured in this run:
  fp16_forward   the same step in the "fp16" precision (the mode that meets the strict heatmap gate)
  inference      configs[4]: frame-sharded forward + arg-max peaks at 256 / 1024 / 4096 frames per GPU
  vit            configs[3]: the ViT-encoder model's training step (value, e2e, fraction of the tensor peak)
  dp_check       N > 1: every rank holds bit-identical parameters after the timed steps, and the N-rank gradient
                 equals the one-rank gradient of the concatenated batch (cosine, norm ratio)
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
MODEL_CIN = {"cnn": 4, "vit": 4, "fourcam": 16}


def workload_text(model_name: str, joints: int, precision: str = "bf16") -> str:
    return (MODEL_NAMES[model_name] + f" C={joints} {precision} training step: fwd + MSE(Gaussian sigma=3 "
            "targets rendered on device from keypoints) + bwd + grad all-reduce + fused Adam")


def workload_config(model_name: str, batch_per_gpu: int, world: int, joints: int) -> dict:
    """the `config` object of the JSON line -- ONE function for both arms: `--impl reference` times the reference's
    CPU path on exactly this workload (a bounded sample of it per step, stated in its `cpu_baseline.sample`)."""
    return {"workload": workload_text(model_name, joints), "batch_per_gpu": batch_per_gpu,
            "global_batch": batch_per_gpu * world, "image": [IMG, IMG, MODEL_CIN[model_name]], "joints": joints,
            "parallelism": f"dp{world}", "l2": "per-step working set (~5 GB of activations) >> 126 MB L2"}


def _ncu_traffic() -> dict:
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the heaviest contraction launches and their
    tensor-pipe activity, from the committed `ncu --set full` captures (profiles/ncu_traffic.json names each source
    file); bench.py measures time live and never runs under a profiler, so these are read, not measured, here."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh)
    return {}


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


def reference_sample_batch(model_name: str, batch_per_gpu: int, steps: int, warmup: int, budget_s: float = 150.0) -> int:
    """samples per step of the CPU arm: the named per-GPU batch when (steps + warmup) passes over it fit the budget at
    the port's ~13 samples/s on 16 host threads (0.45 four-view samples/s for the four-camera model), else the largest
    power-of-two slice of it that does -- the run must end within a few minutes whatever K the driver asks for."""
    rate = 0.45 if model_name == "fourcam" else 13.0
    b = batch_per_gpu
    while b > 1 and (steps + warmup) * b / rate > budget_s:
        b //= 2
    return max(1, b)


def run_reference(args) -> None:
    """`--impl reference`: the reference's own CPU PyTorch path for this step (restated in oracle/, pinned to the real
    modules by tests/golden -- the Python reference itself cannot travel to the GPU box) on every host thread, on the
    GPU arm's config / metric / unit.  Rank 0 alone runs it; each step is a bounded sample of one GPU's shard."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    B = args.batch_per_gpu if args.batch_per_gpu > 0 else (16 if args.model == "fourcam" else BATCH_PER_GPU)
    sample = reference_sample_batch(args.model, B, args.steps, args.warmup)
    while sample > 1:    # one probe step on THIS box's cores: halve the sample while the run would exceed ~4 minutes
        probe, _ = cpu_train_step_seconds(sample, 1, 0, args.model)
        if (args.steps + args.warmup) * probe[0] <= 240.0:
            break
        sample //= 2
    times, cores = cpu_train_step_seconds(sample, args.steps, args.warmup, args.model)
    ms = 1e3 * float(np.mean(times))
    value = sample / (ms / 1e3)
    what = (f"{args.steps} steps of {sample} samples" + (" (one GPU's whole batch)" if sample == B else
            f" (a slice of one GPU's batch of {B})") + f" after {args.warmup} warm-up: fwd + MSE + bwd + Adam in fp32, "
            "oracle/pose_oracle.py on torch CPU")
    line = {
        "impl": "reference", "metric": "train_samples_per_sec", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.model, B, world, JOINTS),
        "note": "CPU arm: the reference's own PyTorch modules restated in oracle/ (the Python reference cannot travel "
                "to the GPU box), fp32, all host threads, rank 0 only; `ms_per_step` is per bounded sample "
                "(`samples_per_step`), `value` = samples_per_step / that",
        "samples_per_step": sample,
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": what},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if not args.no_inference:
        fb = min(sample, 8)
        line["inference"] = {"metric": "inference_frames_per_sec", "unit": "frames/s",
                             "value": cpu_inference_frames_per_sec(fb, 2, args.model),
                             "sample": f"forward + argmax peaks of {fb} frames, best of 2 after 1 warm-up"}
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
        ("gaussian_sep_kernel: sigma=3 targets from keypoints", E * 4 + 8 * B * C, lambda: ops.gaussian_heatmaps(pts)),
        ("argmax_planar_vec_kernel: peaks of 256 frames, fp32 NCHW", IB * C * H * W * 4 + 8 * IB * C,
         lambda: ops.peaks_argmax(hm)),
        ("argmax_planar_vec_kernel: peaks of 256 frames, bf16 NCHW", IB * C * H * W * 2 + 8 * IB * C,
         lambda: ops.peaks_argmax(hm_bf)),
        ("softargmax_kernel (table form): 256 frames, fp32 NCHW", IB * C * H * W * 4 + 8 * IB * C, lambda: ops.peaks_softargmax(hm)),
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
class _Ctx:
    """rank / device plumbing shared by every leg of one bench run."""

    def __init__(self, args):
        import torch.distributed as dist
        self.dist = dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        assert self.world == args.gpus or self.world == 1 and args.gpus == 1, "--gpus must equal WORLD_SIZE"
        self.dev = torch.device("cuda", self.local_rank)
        torch.cuda.set_device(self.dev)
        self.copy_stream = torch.cuda.Stream(device=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps) -> float:
        """milliseconds of `steps` calls: barrier + synchronize on both sides, CUDA events, max over ranks."""
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return ms.item()


MODEL_NAMES = {"cnn": "BasicNet (pytorch/CNNs.py)", "vit": "VIT_encoder_CNN_decoder (pytorch/VITs.py)",
               "fourcam": "FourCamerasBaseLine (pytorch/CNNs.py:189-237, 4 views per sample)"}


class TrainLeg:
    """one model + its data-parallel step + resident and pinned-host inputs; measures `value` and `e2e`."""

    def __init__(self, ctx: _Ctx, model_name: str, precision: str, batch: int, joints: int, graph: bool = True):
        from pose_estimation_amitai_b200 import CNNs, parallel
        self.ctx, self.model_name, self.precision, self.B, self.joints = ctx, model_name, precision, batch, joints
        dev = ctx.dev
        torch.manual_seed(0)  # same random init on every rank (FlatBuckets broadcasts rank 0's anyway)
        if model_name == "vit":       # BASELINE.json configs[3]
            from pose_estimation_amitai_b200 import VITs
            self.model = VITs.VIT_encoder_CNN_decoder(dict(VIT_CFG, precision=precision), np.array((IMG, IMG, 4)), joints).to(dev)
            self.train_gflop, self.fwd_gflop = VIT_TRAIN_GFLOP_PER_SAMPLE, VIT_FWD_GFLOP_PER_SAMPLE
        elif model_name == "fourcam":
            self.model = CNNs.FourCamerasBaseLine(dict(FOURCAM_CFG, precision=precision), np.array((IMG, IMG, 16)), joints).to(dev)
            self.train_gflop, self.fwd_gflop = FOURCAM_TRAIN_GFLOP_PER_SAMPLE, FOURCAM_FWD_GFLOP_PER_SAMPLE
        else:
            self.model = CNNs.BasicNet(dict(CFG, precision=precision), np.array((IMG, IMG, 4)), joints).to(dev)
            self.train_gflop, self.fwd_gflop = TRAIN_GFLOP_PER_SAMPLE, FWD_GFLOP_PER_SAMPLE
        self.dp = parallel.DataParallelStep(self.model, lr=1e-3)
        self.graph = bool(graph)
        if self.graph:
            self.dp.enable_graph()   # the step is replayed from one CUDA graph (captured on its third call)
        self.cin = 16 if model_name == "fourcam" else 4
        self.x_host = torch.rand(batch, self.cin, IMG, IMG, generator=torch.Generator().manual_seed(1 + ctx.rank)).pin_memory()
        self.pts_host = torch.randint(8, IMG - 8, (batch, joints, 2), generator=torch.Generator().manual_seed(2 + ctx.rank)
                                      ).float().pin_memory()
        self.x_dev, self.pts_dev = self.x_host.to(dev), self.pts_host.to(dev)

    def step_resident(self, _i):
        return self.dp.step(self.x_dev, points=self.pts_dev)

    def measure(self, steps: int, warmup: int, clocks: bool = False) -> dict:
        from pose_estimation_amitai_b200 import ops
        ctx = self.ctx
        for i in range(max(warmup, 3) + (2 if self.graph else 0)):   # graph mode: two eager steps precede the capture
            self.step_resident(i)
        sampler = ClockSampler(ctx.local_rank) if clocks else None
        if sampler is not None:
            sampler.start()
        launches0 = ops.launch_count()
        ms_step = ctx.timed(self.step_resident, steps) / steps
        res = {"ms_per_step": ms_step, "value": ctx.world * self.B / (ms_step / 1e3),
               "gpu_launches": int(ops.launch_count() - launches0)}
        if sampler is not None:
            res["clocks"] = sampler.stop()
        return res

    def measure_e2e(self, steps: int) -> dict:
        """the same step fed from pinned HOST memory every step (double-buffered on a copy stream) with the loss
        scalar read back every step."""
        ctx = self.ctx
        bufs = [(torch.empty_like(self.x_dev), torch.empty_like(self.pts_dev)) for _ in range(2)]
        loss_host = torch.zeros(1).pin_memory()
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        probe = os.environ.get("POSEB200_E2E_PROBE", "")   # timing experiments only: "noh2d", "nod2h"

        def issue_copy(slot):
            with torch.cuda.stream(ctx.copy_stream):
                ctx.copy_stream.wait_event(consumed[slot])
                if probe != "noh2d":
                    bufs[slot][0].copy_(self.x_host, non_blocking=True)
                    bufs[slot][1].copy_(self.pts_host, non_blocking=True)
                ready[slot].record(ctx.copy_stream)

        def step_e2e(i):
            slot = i & 1
            if i == 0:
                issue_copy(0)
            issue_copy(slot ^ 1)  # prefetch the next step's inputs while this step computes
            torch.cuda.current_stream().wait_event(ready[slot])
            loss = self.dp.step(bufs[slot][0], points=bufs[slot][1])
            consumed[slot].record(torch.cuda.current_stream())
            if probe != "nod2h":
                loss_host.copy_(loss, non_blocking=True)

        for s in range(2):
            consumed[s].record(torch.cuda.current_stream())
        for i in range(3):
            step_e2e(i)
        ctx.barrier()
        for s in range(2):
            consumed[s].record(torch.cuda.current_stream())
        ms = ctx.timed(step_e2e, steps) / steps
        return {"value": ctx.world * self.B / (ms / 1e3), "unit": "samples/s", "ms_per_step": ms,
                "h2d_bytes_per_step": self.x_host.numel() * 4 + self.pts_host.numel() * 4, "d2h_bytes_per_step": 4,
                "note": "pinned host crops + keypoints copied every step (double-buffered on a copy stream), "
                        "loss scalar read back every step"}

    def workload(self) -> str:
        return workload_text(self.model_name, self.joints, self.precision)

    def close(self):
        torch.cuda.synchronize()
        if self.dp is not None:
            self.dp._graphs.clear()      # captured graphs (they hold NCCL work at N > 1) go before anything else
        self.model = self.dp = self.x_dev = self.pts_dev = self.x_host = self.pts_host = None
        torch.cuda.empty_cache()


def inference_sweep(ctx: _Ctx, model, cin: int, joints: int, fwd_gflop: float, batches, steps: int) -> list:
    """BASELINE.json configs[4]: frame-sharded forward + on-device arg-max peaks (fused into the last layer), no
    collective; frames/s over all ranks at every per-GPU batch in `batches`.  `value`: frames resident in HBM;
    `e2e`: pinned host frames -> device every step (double-buffered), peaks -> pinned host every step."""
    dev, world = ctx.dev, ctx.world
    top = max(batches)
    base = torch.rand(min(256, top), cin, IMG, IMG, generator=torch.Generator().manual_seed(11 + ctx.rank))
    xi_host = torch.empty(top, cin, IMG, IMG).pin_memory()
    for o in range(0, top, base.shape[0]):      # distinct draws are not needed for timing: tile one 256-frame block
        xi_host[o:o + base.shape[0]].copy_(base[:min(base.shape[0], top - o)])
    peaks_host = torch.empty(top, joints, 2).pin_memory()
    out = []
    for IB in batches:
        xh, ph = xi_host[:IB], peaks_host[:IB]
        xi_dev = xh.to(dev)
        xbuf = [torch.empty_like(xi_dev) for _ in range(2)]
        iready = [torch.cuda.Event() for _ in range(2)]
        idone = [torch.cuda.Event() for _ in range(2)]

        def infer_resident(_i):
            model.predict_peaks(xi_dev)

        def infer_copy(slot):
            with torch.cuda.stream(ctx.copy_stream):
                ctx.copy_stream.wait_event(idone[slot])
                xbuf[slot].copy_(xh, non_blocking=True)
                iready[slot].record(ctx.copy_stream)

        def infer_e2e(i):
            slot = i & 1
            if i == 0:
                infer_copy(0)
            infer_copy(slot ^ 1)
            torch.cuda.current_stream().wait_event(iready[slot])
            pk = model.predict_peaks(xbuf[slot])
            idone[slot].record(torch.cuda.current_stream())
            ph.copy_(pk, non_blocking=True)

        isteps = max(3, steps if IB <= 256 else steps // 2)
        for i in range(3 if IB <= 256 else 2):
            infer_resident(i)
        ms_inf = ctx.timed(infer_resident, isteps) / isteps
        for sl in range(2):
            idone[sl].record(torch.cuda.current_stream())
        for i in range(2):
            infer_e2e(i)
        ctx.barrier()
        for sl in range(2):
            idone[sl].record(torch.cuda.current_stream())
        ms_e2e = ctx.timed(infer_e2e, isteps) / isteps
        out.append({"metric": "inference_frames_per_sec", "value": world * IB / (ms_inf / 1e3), "unit": "frames/s",
                    "frames_per_gpu_per_step": IB, "ms_per_step": ms_inf, "steps": isteps,
                    "e2e": {"value": world * IB / (ms_e2e / 1e3), "unit": "frames/s", "ms_per_step": ms_e2e,
                            "h2d_bytes_per_step": xh.numel() * 4, "d2h_bytes_per_step": ph.numel() * 4},
                    "fwd_tflops": world * IB / (ms_inf / 1e3) * fwd_gflop / 1e3})
        del xi_dev, xbuf
        torch.cuda.empty_cache()
    return out


def dp_check(ctx: _Ctx, leg: TrainLeg, per_rank: int = 8, grad_equivalence: bool = True) -> dict:
    """N > 1 only.  (a) after the timed steps every rank must hold BIT-identical parameters (checksums over the
    flat buffer's words, gathered and compared).  (b) SURVEY.md section 4 "DP gradient equivalence": the gradient the
    N ranks obtain from their shards through the bucketed all-reduce (sum / N) against the gradient ONE rank obtains
    from the concatenated N x `per_rank` batch -- cosine and norm ratio over the whole flat gradient."""
    dist, world, dev = ctx.dist, ctx.world, ctx.dev
    b = leg.dp.buckets
    torch.cuda.synchronize()
    mine = b.params_checksum()
    allsums = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allsums, mine)
    identical = all(bool((s == allsums[0]).all()) for s in allsums)
    if not grad_equivalence:
        # the ViT decoder normalises by the min / max of the batch a process holds (pytorch/VITs.py:55-58): N shards
        # are N reference processes, not one process on the concatenated batch, so only (a) applies
        return {"status": "ok" if identical else "FAILED", "params_bit_identical_across_ranks": identical}
    # (b) shard gradients through the production path: hooks fire the bucket all-reduces, no optimiser step
    x, pts = leg.x_dev[:per_rank].contiguous(), leg.pts_dev[:per_rank].contiguous()
    b.reset()
    leg.model.train_step(x, points=pts)
    b.flush()
    b.wait()
    torch.cuda.synchronize()
    g_dp = (b.flat_grad / world).clone()
    xs = [torch.empty_like(x) for _ in range(world)]
    ps = [torch.empty_like(pts) for _ in range(world)]
    dist.all_gather(xs, x)
    dist.all_gather(ps, pts)
    leg.model.set_grad_ready_hook(None)
    try:
        leg.model.train_step(torch.cat(xs), points=torch.cat(ps))
    finally:
        leg.model.set_grad_ready_hook(b.grad_ready)
    torch.cuda.synchronize()
    g_one = b.flat_grad.double()
    cos = (torch.dot(g_dp.double(), g_one) / (g_dp.double().norm() * g_one.norm() + 1e-300)).item()
    ratio = (g_dp.double().norm() / (g_one.norm() + 1e-300)).item()
    stats = torch.tensor([cos, ratio], device=dev, dtype=torch.float64)
    lo, hi = stats.clone(), stats.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    ok = identical and lo[0].item() >= 0.9999 and abs(lo[1].item() - 1) <= 1e-2 and abs(hi[1].item() - 1) <= 1e-2
    return {"status": "ok" if ok else "FAILED", "params_bit_identical_across_ranks": identical,
            "grad_cosine_nrank_vs_1rank_concat": lo[0].item(), "grad_norm_ratio": [lo[1].item(), hi[1].item()],
            "samples_per_rank": per_rank, "gate": "cosine >= 0.9999, |norm ratio - 1| <= 1e-2 (bf16 operands; the two "
            "sides sum the same per-sample terms in different orders)"}


def roofline_record(ctx: _Ctx, leg: TrainLeg, ms_step: float, value: float, reps: int = 3) -> dict:
    """per-launch CUDA-event times of the contraction kernels inside full training steps.  The stream is held behind a
    ~25 ms device-side spin while a step is enqueued, so the GPU never waits for the host between two events and an
    event pair brackets exactly one kernel; per launch the median over `reps` steps is kept."""
    from pose_estimation_amitai_b200 import ops
    runs = []
    for _ in range(reps):
        torch.cuda.synchronize()
        torch.cuda._sleep(50_000_000)
        ops.profile_begin()
        leg.step_resident(0)
        runs.append(ops.profile_end())
    ctx.barrier()
    rec = [(runs[0][i][0], runs[0][i][1], float(np.median([r[i][2] for r in runs]))) for i in range(len(runs[0]))]
    peaks = _peaks()
    by = {}
    for name, flops, ms in rec:
        a = by.setdefault(name, [0.0, 0.0, 0])
        a[0] += flops; a[1] += ms; a[2] += 1
    dname, (dflops, dms, dcount) = max(by.items(), key=lambda kv: kv[1][1])
    achieved = dflops / (dms * 1e-3) / 1e12
    peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    burst = peaks["bf16_tflops"]
    traffic = _ncu_traffic() if leg.model_name == "cnn" else {}
    step_tflops = value / ctx.world * leg.train_gflop / 1e3
    return {
        "kernel": {"pb_conv_tc": "tc_conv2_kernel (+tc_conv_kernel)", "pb_wgrad_tc": "tc_wgrad2_kernel (+tc_wgrad_kernel)",
                   "pb_conv_simt": "conv_simt_kernel", "pb_wgrad_simt": "wgrad_simt_kernel"}.get(dname, dname),
        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        "frac_of_burst_peak": achieved / burst, "burst_peak": burst,
        # dram__bytes_read.sum + dram__bytes_write.sum of this family's heaviest launch, from the committed ncu --set full
        # capture named in profiles/ncu_traffic.json (bench.py never runs under a profiler)
        "traffic": traffic.get("traffic"), "traffic_detail": traffic.get("detail"),
        "launches_per_step": dcount, "avg_launch_ms": dms / dcount, "share_of_step": dms / ms_step,
        "contractions_ms_per_step": sum(v[1] for v in by.values()),
        "peak_source": f"{peaks['_source']} bf16_tflops_sustained (kernel timed inside a long step)",
        "per_family": {k: {"tflops": v[0] / (v[1] * 1e-3) / 1e12, "ms": v[1], "launches": v[2]} for k, v in by.items()},
        "step_tflops": step_tflops, "step_frac_of_peak": step_tflops / peak, "step_frac_of_burst_peak": step_tflops / burst,
    }


def run_gpu(args) -> None:
    ctx = _Ctx(args)
    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    B = args.batch_per_gpu if args.batch_per_gpu > 0 else (16 if args.model == "fourcam" else BATCH_PER_GPU)
    use_graph = not args.no_graph and args.model in ("cnn", "vit")   # the four-camera step is launched eagerly (not captured yet)
    leg = TrainLeg(ctx, args.model, "bf16", B, JOINTS, graph=use_graph)

    m = leg.measure(args.steps, args.warmup, clocks=True)
    ms_step, value = m["ms_per_step"], m["value"]
    e2e = leg.measure_e2e(args.steps)
    line = {
        "metric": "train_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong" if args.batch_per_gpu > 0 and args.batch_per_gpu * world == BATCH_PER_GPU else "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args.model, B, world, JOINTS),
        "launch": "one CUDA-graph replay per step (parallel.DataParallelStep.enable_graph)" if use_graph
                  else "eager launches from Python",
        "grad_buckets_bytes": leg.dp.buckets.bucket_sizes_bytes(),
        "clocks": m["clocks"], "e2e": e2e, "gpu_launches": m["gpu_launches"],
    }

    # ---- N > 1: parameters identical on every rank, N-rank gradient == 1-rank gradient of the concatenated batch
    if world > 1:
        line["dp_check"] = dp_check(ctx, leg)
        if rank == 0:
            print(f"dp_check: {line['dp_check']['status']} {json.dumps(line['dp_check'])}", file=sys.stderr, flush=True)

    # ---- inference sweep (BASELINE.json configs[4])
    if not args.no_inference:
        batches = [args.infer_batch] if args.no_extras or args.infer_batch != 256 else [256, 1024, 4096]
        sweep = inference_sweep(ctx, leg.model, leg.cin, JOINTS, leg.fwd_gflop, batches, max(3, args.steps // 2))
        line["inference"] = dict(sweep[0], workload=f"{type(leg.model).__name__} C={JOINTS} bf16 forward + per-joint "
                                 "argmax peaks on device (fused into the last layer's epilogue), frame-sharded, no collective",
                                 sweep=[{k: v for k, v in s.items() if k != "metric"} for s in sweep])

    # ---- roofline of the dominant kernel family (tcgen05 contractions), timed live: every rank runs the steps
    #      (they contain the gradient all-reduce), rank 0 reports
    roof = roofline_record(ctx, leg, ms_step, value)
    peaks = _peaks()
    leg.close()

    # ---- the same step with IEEE-half forward operands (the precision that meets the strict 2e-2 heatmap gate)
    if args.model == "cnn" and not args.no_extras:
        sub_steps = max(5, args.steps // 2)
        l16 = TrainLeg(ctx, "cnn", "fp16", B, JOINTS, graph=use_graph)
        m16 = l16.measure(sub_steps, 3)
        line["fp16_forward"] = {"metric": "train_samples_per_sec", "value": m16["value"], "unit": "samples/s",
                                "ms_per_step": m16["ms_per_step"], "steps": sub_steps, "dtype": "fp16 forward operands, "
                                "bf16 gradient operands, fp32 accumulation / master weights",
                                "parity": "heatmaps max |err| / (|ref| + 0.1 max|ref|) <= 2e-2 vs the fp32 reference "
                                          "(tests/test_gpu_network.py, smoke())",
                                "e2e": l16.measure_e2e(sub_steps)}
        l16.close()
        # ---- BASELINE.json configs[3]: the ViT-encoder heatmap model, same step definition
        lv = TrainLeg(ctx, "vit", "bf16", B, JOINTS, graph=use_graph)
        mv = lv.measure(sub_steps, 3)
        vt = mv["value"] / world * lv.train_gflop / 1e3
        pk = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
        line["vit"] = {"metric": "train_samples_per_sec", "value": mv["value"], "unit": "samples/s",
                       "ms_per_step": mv["ms_per_step"], "steps": sub_steps, "gpu_launches": mv["gpu_launches"],
                       "config": {"workload": lv.workload(), "batch_per_gpu": B, "global_batch": B * world,
                                  "patch": 16, "dim": 256, "heads": 12, "dim_head": 256, "depth": 8},
                       "e2e": lv.measure_e2e(sub_steps),
                       "roofline": {"bound": "tensor", "achieved": vt, "peak": pk, "unit": "TFLOP/s", "frac": vt / pk,
                                    "note": "whole-step arithmetic rate (46.92 GFLOP per sample)"}}
        if world > 1:
            line["vit"]["dp_check"] = dp_check(ctx, lv, grad_equivalence=False)
        lv.close()

    if rank == 0:
        line["roofline"] = roof
        if not args.no_bandwidth:
            line["bandwidth_kernels"] = {"peak_gbs": peaks["hbm_gbs"], "peak_source": peaks["_source"] + " hbm_gbs",
                                         "note": "timings include the launch of small helper kernels each op needs "
                                                 "(loss zeroing, key finalize); the peak is the driver-measured COPY "
                                                 "rate (read + write), which a read-only stream such as arg-max can "
                                                 "exceed by a few per cent",
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
        # teardown must never outlive the measurement: with collectives captured in CUDA graphs,
        # destroy_process_group() was seen to block (profiles/r2k notes), so the ranks meet at a barrier and leave
        # without it -- and a watchdog ends the process should even that stall
        sys.stdout.flush()
        sys.stderr.flush()
        threading.Timer(30.0, lambda: os._exit(0)).start()
        ctx.barrier()
        os._exit(0)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-inference", action="store_true")
    ap.add_argument("--no-bandwidth", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel of the step from Python instead of "
                    "replaying the step's CUDA graph")
    ap.add_argument("--no-extras", action="store_true",
                    help="headline only: skip the fp16_forward / vit sub-records and the 1024 / 4096 inference points")
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
