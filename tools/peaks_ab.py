"""Time the planar arg-max / soft arg-max launches (Augmentor.py:105-148, utils.py:47-83) at the inference sweep's
shape (256 frames x 36 maps of 192 x 192) and the Gaussian target renderer (simple_data_generator.py:119-136, 64 x 36
maps) in their round-2 forms and in the forms they replace (POSEB200_ARGMAX_V1 / POSEB200_SOFTARGMAX_V1 /
POSEB200_GAUSS_V1, read at launch time).   usage: python tools/peaks_ab.py [--iters 10]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from pose_estimation_amitai_b200 import ops


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda")
    n, c, h, w = 256, 36, 192, 192
    hm = torch.rand(n, c, h, w, device=dev)
    hm_bf = hm.to(torch.bfloat16)
    peak = 6546.2
    try:
        import json
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                           "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass

    def timed(fn):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(a.iters):          # 680 MB / 1.36 GB operands >> 126 MB L2: no flush needed
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        return ts[len(ts) // 2]

    cases = [("argmax fp32", "POSEB200_ARGMAX_V1", hm, ops.peaks_argmax),
             ("argmax bf16", "POSEB200_ARGMAX_V1", hm_bf, ops.peaks_argmax),
             ("softargmax fp32", "POSEB200_SOFTARGMAX_V1", hm, ops.peaks_softargmax),
             ("softargmax bf16", "POSEB200_SOFTARGMAX_V1", hm_bf, ops.peaks_softargmax)]
    pts = torch.randint(8, 184, (64, c, 2), device=dev).float()
    cases.append(("gaussian render", "POSEB200_GAUSS_V1", pts, ops.gaussian_heatmaps))
    for name, var, t, fn in cases:
        nbytes = t.numel() * t.element_size() + 8 * n * c
        if fn is ops.gaussian_heatmaps:
            nbytes = 64 * c * h * w * 4 + 8 * 64 * c      # a write stream
        os.environ.pop(var, None)
        new = timed(lambda: fn(t))
        r_new = fn(t)
        os.environ[var] = "1"
        old = timed(lambda: fn(t))
        r_old = fn(t)
        os.environ.pop(var, None)
        same = torch.equal(r_new, r_old)
        print(f"{name:16s} {old:7.1f} us ({nbytes / old / 1e3 / peak:.3f} of HBM peak) -> {new:7.1f} us "
              f"({nbytes / new / 1e3 / peak:.3f})   identical results: {same}"
              + ("" if same else f" (max |diff| {(r_new - r_old).abs().max().item():.3e})"), flush=True)


if __name__ == "__main__":
    main()
