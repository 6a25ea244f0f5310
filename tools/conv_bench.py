#!/usr/bin/env python
"""Per-layer timing + cross-check of the tcgen05 conv kernels (GPU box only).

    python tools/conv_bench.py [--batch 64] [--mode NAME ...]

For every contraction shape of BasicNet (forward and input-gradient form) runs the per-tap kernel
(tc_conv.cu, POSEB200_CONV_V1=1) as the cross-check and each requested plan of tc_conv2.cu, prints the
max |difference| against the per-tap kernel's output and the CUDA-event time / algorithmic TFLOP/s.
Parity against the oracle is tests/test_gpu_kernels.py's job; this is the measurement loop.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

from pose_estimation_amitai_b200 import ops, tc_support

MODES = {
    "v1": {"POSEB200_CONV_V1": "1"},
    "default": {},
    "rolled": {"POSEB200_CONV_UNROLL": "0"},
    "rolled_mma": {"POSEB200_CONV_UNROLL": "0", "POSEB200_CONV_CLUSTER": "1", "POSEB200_CONV_DEBUG": "7"},
    "nopair": {"POSEB200_CONV_PAIR": "0"},
    "st128": {"POSEB200_CONV_DEBUG": "8"},
    "keepl2": {"POSEB200_CONV_KEEP_L2": "1"},
    "pair128": {"POSEB200_CONV_PAIR_MIN_N": "128"},
    "pair64": {"POSEB200_CONV_PAIR_MIN_N": "64"},
    "x_notmem": {"POSEB200_CONV_DEBUG": "16"},
    "T2_nopoll": {"POSEB200_TC_T": "2", "POSEB200_CONV_POLL_NS": "0"},
    "es4": {"POSEB200_CONV_ESTAGES": "4", "POSEB200_CONV_POLL_NS": "100"},
    "es4_T4": {"POSEB200_CONV_ESTAGES": "4", "POSEB200_CONV_POLL_NS": "100", "POSEB200_TC_T": "4"},
    "T4_poll": {"POSEB200_CONV_POLL_NS": "100", "POSEB200_TC_T": "4"},
    "poll20": {"POSEB200_CONV_POLL_NS": "20"},
    "poll50": {"POSEB200_CONV_POLL_NS": "50"},
    "poll100": {"POSEB200_CONV_POLL_NS": "100"},
    "poll200": {"POSEB200_CONV_POLL_NS": "200"},
    "unroll1": {"POSEB200_CONV_UNROLL": "1"},
    "unroll1_poll50": {"POSEB200_CONV_UNROLL": "1", "POSEB200_CONV_POLL_NS": "50"},
    # multicast clusters without cta pairs: slower than pairs on every layer measured (conv5 127 -> 155 / 170 us with
    # 2 / 4 CTAs, conv8 114 -> 130 / 136), and the stride-2 'up' layers did not terminate with it (run killed by its
    # timeout) -- do NOT use on convT1 / vitdc3
    "nopair_cl2": {"POSEB200_CONV_PAIR": "0", "POSEB200_CONV_CLUSTER": "2"},
    "nopair_cl4": {"POSEB200_CONV_PAIR": "0", "POSEB200_CONV_CLUSTER": "4"},
    "nospec": {"POSEB200_CONV_EPI_SPEC": "0"},
    "x_noq1": {"POSEB200_CONV_DEBUG": "256"},
    "x_noq2": {"POSEB200_CONV_DEBUG": "512"},
    "x_nosts": {"POSEB200_CONV_DEBUG": "32"},
    "x_nostore": {"POSEB200_CONV_DEBUG": "64"},
    "x_nomask": {"POSEB200_CONV_DEBUG": "128"},
    "x_nosts_nostore_nomask": {"POSEB200_CONV_DEBUG": "224"},
    "x_nostore_nomask": {"POSEB200_CONV_DEBUG": "192"},
    "np2": {"POSEB200_CONV_NPASS": "2"},
    "np2_T2": {"POSEB200_CONV_NPASS": "2", "POSEB200_TC_T": "2"},
    "np1_T2": {"POSEB200_CONV_NPASS": "1", "POSEB200_TC_T": "2"},
    "np_mmaonly": {"POSEB200_CONV_PAIR": "0", "POSEB200_CONV_DEBUG": "7"},
    "np_noB": {"POSEB200_CONV_PAIR": "0", "POSEB200_CONV_DEBUG": "2"},
    "np_noepi": {"POSEB200_CONV_PAIR": "0", "POSEB200_CONV_DEBUG": "1"},
    "halo_T1": {"POSEB200_TC_T": "1"},
    "halo_T2": {"POSEB200_TC_T": "2"},
    "halo_T4": {"POSEB200_TC_T": "4"},
    "strips": {"POSEB200_CONV_PLAN_HALO": "0"},
    "nostage": {"POSEB200_TC_NO_STAGED_EPI": "1"},
    "cl1": {"POSEB200_CONV_CLUSTER": "1"},
    "x_noepi": {"POSEB200_CONV_CLUSTER": "1", "POSEB200_CONV_DEBUG": "1"},
    "x_noB": {"POSEB200_CONV_CLUSTER": "1", "POSEB200_CONV_DEBUG": "2"},
    "x_noA": {"POSEB200_CONV_CLUSTER": "1", "POSEB200_CONV_DEBUG": "4"},
    "x_noAB": {"POSEB200_CONV_CLUSTER": "1", "POSEB200_CONV_DEBUG": "6"},
    "x_mmaonly": {"POSEB200_CONV_CLUSTER": "1", "POSEB200_CONV_DEBUG": "7"},
    "x_noepiB": {"POSEB200_CONV_CLUSTER": "1", "POSEB200_CONV_DEBUG": "3"},
    "cl2": {"POSEB200_CONV_CLUSTER": "2"},
    "cl4": {"POSEB200_CONV_CLUSTER": "4"},
    "cl2_T1": {"POSEB200_CONV_CLUSTER": "2", "POSEB200_TC_T": "1"},
    "cl4_T1": {"POSEB200_CONV_CLUSTER": "4", "POSEB200_TC_T": "1"},
    "cl4_T2": {"POSEB200_CONV_CLUSTER": "4", "POSEB200_TC_T": "2"},
    "nostage_T1": {"POSEB200_TC_NO_STAGED_EPI": "1", "POSEB200_TC_T": "1"},
}
KNOBS = ["POSEB200_CONV_V1", "POSEB200_TC_T", "POSEB200_CONV_COLS8", "POSEB200_CONV_PLAN_HALO", "POSEB200_CONV_BASEOFF",
         "POSEB200_TC_NO_BRES", "POSEB200_TC_NO_STAGED_EPI", "POSEB200_CONV_CLUSTER", "POSEB200_CONV_DEBUG", "POSEB200_CONV_PAIR", "POSEB200_CONV_PAIR_MIN_N", "POSEB200_CONV_NPASS", "POSEB200_CONV_KEEP_L2",
         "POSEB200_CONV_UNROLL", "POSEB200_CONV_POLL_NS", "POSEB200_CONV_ESTAGES", "POSEB200_CONV_EPI_SPEC"]

# (name, kind, cin, cout, h, w, dilation, what)
SHAPES = [
    ("conv1 lin", "linear", 64, 64, 192, 192, 1, "fwd_nores"),
    ("conv2 fwd", "conv", 64, 64, 192, 192, 2, "fwd"),
    ("conv2 fself", "conv", 64, 64, 192, 192, 2, "fwd_self"),
    ("conv2 nores", "conv", 64, 64, 192, 192, 2, "fwd_nores"),   # residual operand == the layer's own input (CNNs.py:75)
    ("conv5 fself", "conv", 128, 128, 96, 96, 2, "fwd_self"),
    ("conv2 dgrad", "conv", 64, 64, 192, 192, 2, "dgrad"),
    ("conv2 dg_noG", "conv", 64, 64, 192, 192, 2, "dgrad_noG"),
    ("conv2 dg_nomask", "conv", 64, 64, 192, 192, 2, "dgrad_nomask"),
    ("conv2 dg_noskip", "conv", 64, 64, 192, 192, 2, "dgrad_noskip"),
    ("conv4 nores", "conv", 64, 128, 96, 96, 2, "fwd_nores"),
    ("conv4 fwd", "conv", 64, 128, 96, 96, 2, "fwd"),
    ("conv4 dgrad", "conv", 64, 128, 96, 96, 2, "dgrad"),
    ("conv5 fwd", "conv", 128, 128, 96, 96, 2, "fwd"),
    ("conv7 fwd", "conv", 128, 256, 48, 48, 2, "fwd"),
    ("conv7 dgrad", "conv", 128, 256, 48, 48, 2, "dgrad"),
    ("conv8 fwd", "conv", 256, 256, 48, 48, 2, "fwd"),
    ("convT1 fwd", "convT2", 256, 128, 48, 48, 1, "fwd"),
    ("convT1 dgrad", "convT2", 256, 128, 48, 48, 1, "dgrad"),
    ("convT2 fwd", "convT1", 128, 128, 96, 96, 1, "fwd"),
    ("vitdc3 fwd", "convT2", 256, 256, 48, 48, 1, "fwd_nores"),
    ("vitdc3 dgrad", "convT2", 256, 256, 48, 48, 1, "dgrad"),
    ("convT4 fwd", "convT2", 128, 36, 96, 96, 1, "fwd"),
    ("convT4 dgrad", "convT2", 128, 36, 96, 96, 1, "dgrad"),
]


def set_mode(env):
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update(env)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--mode", nargs="*", default=["default"])
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    n = args.batch
    g = torch.Generator().manual_seed(0)
    for name, kind, cin, cout, h, w, dil, what in SHAPES:
        if args.only and not any(o.strip() in name for o in args.only.split(",")):
            continue
        spec = ops.Contraction(kind, cin, cout, dilation=dil)
        wshape = (cout, cin, 3, 3) if kind == "conv" else ((cout, cin) if kind == "linear" else (cin, cout, 3, 3))
        wt = ((torch.rand(wshape, generator=g) - 0.5) * (2.0 / (3 * cin ** 0.5))).to(dev)
        oh, ow = spec.out_hw(h, w)
        if what.startswith("fwd"):
            x = (torch.rand(n, h, w, cin, generator=g) - 0.5).to(dev, torch.bfloat16)
            wp = ops.pack_weights(wt, spec, "oi", torch.bfloat16, ipad=tc_support.pad_n(cout))
            res = (torch.rand(n, oh, ow, cout, generator=g) - 0.5).to(dev, torch.bfloat16)
            bias = (torch.rand(cout, generator=g) - 0.5).to(dev)
            mask = torch.zeros((n * oh * ow, (cout + 31) // 32), device=dev, dtype=torch.int32)
            if what == "fwd_nores":
                run = lambda: ops.conv("tc", x, wp, spec.fwd_taps(), n, h, w, cin, oh, ow, cout, bias=bias,
                                       act=ops.PB_ACT_LRELU, mask_out=mask, act_dtype=torch.bfloat16)
            elif what == "fwd_self":
                run = lambda: ops.conv("tc", x, wp, spec.fwd_taps(), n, h, w, cin, oh, ow, cout, bias=bias,
                                       act=ops.PB_ACT_LRELU, add1=x, mask_out=mask, act_dtype=torch.bfloat16)
            elif cout % 8 == 0:
                run = lambda: ops.conv("tc", x, wp, spec.fwd_taps(), n, h, w, cin, oh, ow, cout, bias=bias,
                                       act=ops.PB_ACT_LRELU, add1=res, mask_out=mask, act_dtype=torch.bfloat16)
            else:  # network head: NCHW fp32 output, bias + LeakyReLU only
                run = lambda: ops.conv("tc", x, wp, spec.fwd_taps(), n, h, w, cin, oh, ow, cout, bias=bias,
                                       act=ops.PB_ACT_LRELU, act_dtype=torch.bfloat16, out_nchw=True)
            macs = n * oh * ow * spec.ntaps * cin * cout // (4 if kind == "convT2" else 1)
        else:
            cpad = (cout + 7) // 8 * 8
            x = torch.zeros((n, oh, ow, cpad), device=dev, dtype=torch.bfloat16)
            x[..., :cout] = (torch.rand(n, oh, ow, cout, generator=g) - 0.5).to(dev, torch.bfloat16)
            wp = ops.pack_weights(wt, spec, "io", torch.bfloat16, jpad=cpad)
            skip = (torch.rand(n, h, w, cin, generator=g) - 0.5).to(dev, torch.bfloat16)
            mprev = torch.randint(-2 ** 31, 2 ** 31 - 1, (n * h * w, (cin + 31) // 32), generator=g,
                                  dtype=torch.int64).to(torch.int32).to(dev)
            gpre = torch.empty((n, h, w, cin), device=dev, dtype=torch.bfloat16)
            if what == "dgrad_noG":
                run = lambda: ops.conv("tc", x, wp, spec.dgrad_taps(), n, oh, ow, cpad, h, w, cin, add0=skip,
                                       act=ops.PB_ACT_MASKMUL, mask_in=mprev, act_dtype=torch.bfloat16)
            elif what == "dgrad_nomask":
                run = lambda: ops.conv("tc", x, wp, spec.dgrad_taps(), n, oh, ow, cpad, h, w, cin, add0=skip, pre_out=gpre,
                                       act=ops.PB_ACT_NONE, act_dtype=torch.bfloat16)
            elif what == "dgrad_noskip":
                run = lambda: ops.conv("tc", x, wp, spec.dgrad_taps(), n, oh, ow, cpad, h, w, cin, pre_out=gpre,
                                       act=ops.PB_ACT_MASKMUL, mask_in=mprev, act_dtype=torch.bfloat16)
            else:
                run = lambda: ops.conv("tc", x, wp, spec.dgrad_taps(), n, oh, ow, cpad, h, w, cin, add0=skip, pre_out=gpre,
                                       act=ops.PB_ACT_MASKMUL, mask_in=mprev, act_dtype=torch.bfloat16)
            macs = n * h * w * 9 * cin * cout // (1 if kind != "convT2" else 1) * (1 if kind != "convT2" else 1)
            if kind == "convT2":
                macs = n * h * w * 9 * cin * cout  # every input pixel x 9 taps
        set_mode(MODES["v1"])
        ref = run().float()
        torch.cuda.synchronize()
        for mode in ["v1"] + [m for m in args.mode if m != "v1"]:
            set_mode(MODES[mode])
            try:
                out = run().float()
                torch.cuda.synchronize()
            except Exception as e:  # noqa: BLE001
                print(f"{name:14s} {mode:10s} FAILED: {e}", flush=True)
                continue
            diff = (out - ref).abs().max().item()
            for _ in range(2):
                run()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.iters
            print(f"{name:14s} {mode:10s} maxdiff {diff:9.3e}  {ms * 1e3:8.1f} us  {2 * macs / ms / 1e9:7.1f} TFLOP/s",
                  flush=True)


if __name__ == "__main__":
    main()
