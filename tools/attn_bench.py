"""Time the fused attention launches (pytorch_vit_encoder.py:59-78) at the ViT's shape: batch 64, 144 tokens, 12 heads
of 256 features.  POSEB200_ATTN_EARLY is read once per process -> both settings run in child processes.
usage: python tools/attn_bench.py [--iters 20]"""
import argparse
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child(iters: int) -> None:
    import torch
    from pose_estimation_amitai_b200 import vit_ops
    dev = torch.device("cuda")
    b, s, h, d = 64, 144, 12, 256
    qkv = (torch.randn(b * s, 3 * h * d, device=dev) * 0.5).bfloat16()
    go = torch.randn(b * s, h * d, device=dev).bfloat16()
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    scale = d ** -0.5
    o, probs = vit_ops.attention_fwd(qkv, b, s, h, d, scale)

    def timed(fn):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        return ts[len(ts) // 2]

    f = timed(lambda: vit_ops.attention_fwd(qkv, b, s, h, d, scale))
    bw = timed(lambda: vit_ops.attention_bwd(qkv, probs, go, b, s, h, d, scale))
    gq = vit_ops.attention_bwd(qkv, probs, go, b, s, h, d, scale)
    print(f"  forward {f:6.1f} us   backward (A + B) {bw:6.1f} us   checksums {o.float().abs().sum().item():.6e} "
          f"{gq.float().abs().sum().item():.6e}", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--child", action="store_true")
    a = ap.parse_args()
    if a.child:
        child(a.iters)
    else:
        for setting in ("0", "1"):
            print(f"POSEB200_ATTN_EARLY={setting}", flush=True)
            env = dict(os.environ, POSEB200_ATTN_EARLY=setting)
            subprocess.run([sys.executable, os.path.abspath(__file__), "--child", "--iters", str(a.iters)], env=env,
                           check=True, timeout=300)
