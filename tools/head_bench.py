#!/usr/bin/env python
"""Timing of the network head (ConvTranspose2d 128 -> C, k3 s2 + LeakyReLU at 96^2 -> 192^2) in its three forms --
NCHW fp32 heatmaps / fused arg-max / fused MSE + gradient -- on the folded-parity kernel (csrc/tc_head.cu) and on the
generic halo kernel (POSEB200_HEAD_V2=0).  GPU box only.   python tools/head_bench.py [--batch 64] [--joints 36]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

from pose_estimation_amitai_b200 import ops, tc_support


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--joints", type=int, default=36)
    ap.add_argument("--cin", type=int, default=128)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    n, cin, c, ih, iw = args.batch, args.cin, args.joints, 96, 96
    g = torch.Generator().manual_seed(0)
    spec = ops.Contraction("convT2", cin, c)
    wt = ((torch.rand(cin, c, 3, 3, generator=g) - 0.5) * (2.0 / (3 * cin ** 0.5))).to(dev)
    bias = (torch.rand(c, generator=g) - 0.5).to(dev)
    x = (torch.rand(n, ih, iw, cin, generator=g) - 0.5).to(dev, torch.bfloat16)
    wf = ops.pack_weights(wt, spec, "oi", torch.bfloat16, ipad=tc_support.pad_n(c))
    pts = torch.randint(8, 184, (n, c, 2), generator=g).float().to(dev)
    out = torch.empty((n, c, 2 * ih, 2 * iw), device=dev)
    cases = {
        "store": lambda: ops.conv("tc", x, wf, spec.fwd_taps(), n, ih, iw, cin, 2 * ih, 2 * iw, c, bias=bias,
                                  act=ops.PB_ACT_LRELU, act_dtype=torch.bfloat16, out_nchw=True, out=out),
        "argmax": lambda: ops.head_argmax_fused(x, wf, spec.fwd_taps(), n, ih, iw, cin, c, bias=bias),
        "mse": lambda: ops.head_mse_fused(x, wf, spec.fwd_taps(), n, ih, iw, cin, c, bias=bias, points=pts),
    }
    flops = 2.0 * n * ih * iw * 9 * cin * c
    for name, fn in cases.items():
        if args.only and args.only not in name:
            continue
        for v2 in ("1", "0"):
            os.environ["POSEB200_HEAD_V2"] = v2
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / args.iters * 1e3
            print(f"head {name:7s} {'folded ' if v2 == '1' else 'generic'}  {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s", flush=True)
    os.environ.pop("POSEB200_HEAD_V2", None)


if __name__ == "__main__":
    main()
