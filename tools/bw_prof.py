#!/usr/bin/env python
"""One launch of each HBM-bound heatmap / loss / peak kernel at the bench workload size (GPU box only), for
`ncu --set full -k regex:<kernel>`:   python tools/bw_prof.py [mse|gauss|argmax|argmax_bf16|softargmax|pool|affine|affine_u8|conv1|attn]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

from pose_estimation_amitai_b200 import ops

which = sys.argv[1] if len(sys.argv) > 1 else "mse"
dev = torch.device("cuda:0")
B, C, H, W = 64, 36, 192, 192
if which in ("mse", "mse_target"):
    out = torch.rand(B, C, H, W, device=dev) - 0.3
    pts = torch.randint(8, 184, (B, C, 2), device=dev).float()
    tgt = ops.gaussian_heatmaps(pts) if which == "mse_target" else None
    for _ in range(2):
        ops.mse_loss_fwd_bwd(out, tgt, points=None if which == "mse_target" else pts,
                             grad_nhwc_dtype=torch.bfloat16, cpad=48)
elif which == "gauss":
    pts = torch.randint(8, 184, (B, C, 2), device=dev).float()
    for _ in range(2):
        ops.gaussian_heatmaps(pts)
elif which in ("argmax", "argmax_bf16", "softargmax"):
    hm = torch.rand(256, C, H, W, device=dev)
    if which == "argmax_bf16":
        hm = hm.to(torch.bfloat16)
    for _ in range(2):
        ops.peaks_softargmax(hm) if which == "softargmax" else ops.peaks_argmax(hm)
elif which == "pool":
    x = (torch.rand(B, H, W, 64, device=dev) - 0.5).to(torch.bfloat16)
    gy = (torch.rand(B, H // 2, W // 2, 64, device=dev) - 0.5).to(torch.bfloat16)
    mask = torch.randint(-2 ** 31, 2 ** 31 - 1, (B * H * W, 2), device=dev, dtype=torch.int64).to(torch.int32)
    for _ in range(2):
        ops.maxpool_lrelu_fwd(x)
        ops.maxpool_lrelu_bwd(x, gy, mask)
elif which == "affine":
    import math
    import numpy as np
    rs = np.random.RandomState(0)
    th = np.array([[math.cos(a), math.sin(a), tx, -math.sin(a), math.cos(a), ty] for a, tx, ty in
                   zip(np.radians(rs.uniform(-30, 30, B)), rs.uniform(-10, 10, B), rs.uniform(-10, 10, B))], np.float32)
    theta = torch.from_numpy(th).to(dev)
    flips = torch.from_numpy(rs.randint(0, 4, B).astype(np.int32)).to(dev)
    src = torch.from_numpy(rs.permutation(256)[:B].astype(np.int32)).to(dev)
    hm = torch.rand(256, C, H, W, device=dev)
    for _ in range(2):
        ops.affine_nearest(hm, theta, flips, src_index=src)
elif which == "affine_u8":
    import math
    import time
    import numpy as np
    IB = 256
    rs = np.random.RandomState(0)
    th = np.array([[math.cos(a), math.sin(a), tx, -math.sin(a), math.cos(a), ty] for a, tx, ty in
                   zip(np.radians(rs.uniform(-30, 30, IB)), rs.uniform(-10, 10, IB), rs.uniform(-10, 10, IB))], np.float32)
    theta = torch.from_numpy(th).to(dev)
    flips = torch.from_numpy(rs.randint(0, 4, IB).astype(np.int32)).to(dev)
    src = torch.from_numpy(rs.permutation(IB).astype(np.int32)).to(dev)
    box = torch.randint(0, 256, (IB, 4, H, W), device=dev, dtype=torch.uint8)
    out = torch.empty(IB, 4, H, W, device=dev)
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    for _ in range(3):
        ops.affine_nearest(box, theta, flips, src_index=src, out=out)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.affine_nearest(box, theta, flips, src_index=src, out=out); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print("affine u8 256 x 4 x 192^2 (L2 flushed): median %.1f us, %.0f GB/s" % (ts[5], IB * 4 * H * W * 5 / ts[5] / 1e3))
elif which == "conv1":
    from pose_estimation_amitai_b200 import tc_support
    x = torch.rand(B, 4, H, W, device=dev)
    lin = ops.Contraction("linear", 36, 64)
    wt = (torch.rand(64, 4, 3, 3, device=dev) - 0.5) * 0.3
    wp = ops.pack_weights(wt, lin, "oi", torch.bfloat16, ipad=tc_support.pad_n(64), jpad=64)
    bias = torch.rand(64, device=dev) - 0.5
    mask = torch.zeros((B * H * W, 2), device=dev, dtype=torch.int32)
    import time
    for _ in range(3):
        ops.conv_first(x, wp, bias, 64, 2, torch.bfloat16, mask_out=mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.conv_first(x, wp, bias, 64, 2, torch.bfloat16, mask_out=mask)
    e1.record()
    torch.cuda.synchronize()
    print("conv_first batch 64: %.1f us" % (e0.elapsed_time(e1) / 5 * 1e3))
elif which == "attn":
    from pose_estimation_amitai_b200 import vit_ops
    b, s, h, d = 64, 144, 12, 256     # one encoder layer of the ViT bench step
    qkv = (torch.randn(b * s, 3 * h * d, device=dev) * 0.5).to(torch.bfloat16)
    go = (torch.randn(b * s, h * d, device=dev) * 0.5).to(torch.bfloat16)
    for _ in range(2):
        o, probs = vit_ops.attention_fwd(qkv, b, s, h, d, d ** -0.5)
        vit_ops.attention_bwd(qkv, probs, go, b, s, h, d, d ** -0.5)
torch.cuda.synchronize()
print("ok", which)
