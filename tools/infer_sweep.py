#!/usr/bin/env python
"""Inference sweep (BASELINE.json configs[4]): frame-sharded forward + on-device argmax peaks, frames per GPU-shard
in {256, 512, 1024, 2048, 4096}; one process per GPU (torchrun for N > 1), no collective on the data path.
Prints one JSON line per batch size on rank 0 (frames/s over all ranks; max-over-ranks device time)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from bench import CFG, FWD_GFLOP_PER_SAMPLE, IMG, JOINTS


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", type=int, nargs="*", default=[256, 512, 1024, 2048, 4096])
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    import torch.distributed as dist
    from pose_estimation_amitai_b200 import CNNs
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    model = CNNs.BasicNet(dict(CFG), np.array((IMG, IMG, 4)), JOINTS).to(dev)
    for nb in args.batches:
        x = torch.rand(nb, 4, IMG, IMG, device=dev)
        for _ in range(2):
            model.predict_peaks(x)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            pk = model.predict_peaks(x)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            fps = world * nb / (ms.item() / 1e3)
            print(json.dumps({"metric": "inference_frames_per_sec", "frames_per_gpu": nb, "n_gpus": world,
                              "ms_per_step": ms.item(), "value": fps, "fwd_tflops_per_gpu": fps / world *
                              FWD_GFLOP_PER_SAMPLE / 1e3, "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}),
                  flush=True)
        del x, pk
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
