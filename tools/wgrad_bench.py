#!/usr/bin/env python
"""Per-layer timing + cross-check of the tcgen05 weight-gradient kernels (GPU box only): the per-tap kernel
(tc_wgrad.cu, POSEB200_WGRAD_V1=1) against the halo-resident one (tc_wgrad2.cu)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

from pose_estimation_amitai_b200 import ops

SHAPES = [
    ("conv1 lin", "linear", 64, 64, 192, 192, 1),
    ("conv2", "conv", 64, 64, 192, 192, 2),
    ("conv4", "conv", 64, 128, 96, 96, 2),
    ("conv5", "conv", 128, 128, 96, 96, 2),
    ("conv7", "conv", 128, 256, 48, 48, 2),
    ("conv8", "conv", 256, 256, 48, 48, 2),
    ("convT2", "convT1", 128, 128, 96, 96, 1),
    ("convT1", "convT2", 256, 128, 48, 48, 1),
    ("convT4", "convT2", 128, 36, 96, 96, 1),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    n = args.batch
    g = torch.Generator().manual_seed(0)
    for name, kind, cin, cout, h, w, dil in SHAPES:
        if args.only and args.only not in name:
            continue
        spec = ops.Contraction(kind, cin, cout, dilation=dil)
        wshape = (cout, cin, 3, 3) if kind == "conv" else ((cout, cin) if kind == "linear" else (cin, cout, 3, 3))
        oh, ow = spec.out_hw(h, w)
        cpad = (cout + 15) // 16 * 16 if cout % 8 else cout
        x = (torch.randint(-8, 9, (n, h, w, cin), generator=g).float() / 8).to(dev, torch.bfloat16)
        dc = torch.zeros((n, oh, ow, cpad), device=dev, dtype=torch.bfloat16)
        dc[..., :cout] = (torch.randint(-8, 9, (n, oh, ow, cout), generator=g).float() / 8).to(dev, torch.bfloat16)
        res = {}
        for mode in ("v1", "v2"):
            os.environ["POSEB200_WGRAD_V1"] = "1" if mode == "v1" else "0"
            dw = torch.full(wshape, float("nan"), device=dev)
            db = torch.full((cout,), float("nan"), device=dev)
            run = lambda: ops.wgrad("tc", spec, x, dc, n, h, w, dw, db, act_dtype=torch.bfloat16)
            try:
                run()
                torch.cuda.synchronize()
            except Exception as e:  # noqa: BLE001
                print(f"{name:10s} {mode} FAILED: {e}", flush=True)
                continue
            res[mode] = (dw.clone(), db.clone())
            for _ in range(2):
                run()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.iters
            macs = n * h * w * spec.ntaps * cin * cout
            diff = ""
            if mode == "v2" and "v1" in res:
                sc = res["v1"][0].abs().max().item() + 1e-30
                diff = f"max|dw-dw_v1|/max|dw| {(dw - res['v1'][0]).abs().max().item() / sc:9.3e}  db diff " \
                       f"{(db - res['v1'][1]).abs().max().item():9.3e}"
            print(f"{name:10s} {mode}  {ms * 1e3:8.1f} us (incl. bias + reduce)  {2 * macs / ms / 1e9:7.1f} TFLOP/s  {diff}",
                  flush=True)


if __name__ == "__main__":
    main()
