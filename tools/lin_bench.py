"""Time the ViT's nn.Linear contractions (pytorch_vit_encoder.py:20-23,52,55) at batch 64 (9216 token rows) through
pb_conv_tc, with the staged TMA-store epilogue on and off (POSEB200_LINEAR_TMA_EPI, read once per process -> the two
settings run in child processes).   usage: python tools/lin_bench.py [--iters 20]"""
import argparse
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SHAPES = [  # name, cin, cout, bias
    ("to_qkv fwd   256 -> 9216", 256, 9216, False),
    ("to_out dgrad 256 -> 3072", 256, 3072, False),
    ("ff2 dgrad    256 -> 1024", 256, 1024, False),
    ("to_qkv dgrad 9216 -> 256", 9216, 256, False),
    ("to_out fwd   3072 -> 256", 3072, 256, True),
]


def child(iters: int) -> None:
    import torch
    from pose_estimation_amitai_b200 import ops
    dev = torch.device("cuda")
    rows = 9216
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    for name, cin, cout, with_bias in SHAPES:
        x = (torch.rand(rows, cin, device=dev) - 0.5).bfloat16()
        wt = (torch.rand(cout, cin, device=dev) - 0.5) * (2.0 / cin ** 0.5)
        bias = (torch.rand(cout, device=dev) - 0.5) if with_bias else None
        lin = ops.Contraction("linear", cin, cout)
        wp = ops.pack_weights(wt, lin, "oi", torch.bfloat16)
        out = torch.empty((1, 1, rows, cout), device=dev, dtype=torch.bfloat16)
        run = lambda: ops.conv("tc", x, wp, lin.fwd_taps(), 1, 1, rows, cin, 1, rows, cout, bias=bias, out=out,
                               act_dtype=torch.bfloat16)
        for _ in range(3):
            run()
        ts = []
        for _ in range(iters):
            flush.zero_()                      # L2 flush between timed launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        us = ts[len(ts) // 2]
        want = x.float() @ wt.bfloat16().float().t()
        if with_bias:
            want = want + bias
        err = (out.view(rows, cout).float() - want).abs().max().item()
        print(f"  {name}: {us:7.1f} us  {2.0 * rows * cin * cout / us / 1e6:7.1f} TFLOP/s  "
              f"out {rows * cout * 2 / us / 1e3:6.0f} GB/s  max|err| {err:.3e}", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--child", action="store_true")
    a = ap.parse_args()
    if a.child:
        child(a.iters)
    else:
        for setting in ("0", "1"):
            print(f"POSEB200_LINEAR_TMA_EPI={setting}", flush=True)
            env = dict(os.environ, POSEB200_LINEAR_TMA_EPI=setting)
            subprocess.run([sys.executable, os.path.abspath(__file__), "--child", "--iters", str(a.iters)], env=env,
                           check=True, timeout=300)
