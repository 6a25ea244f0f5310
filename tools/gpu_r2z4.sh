#!/bin/bash
# Gaussian renderer pass: kernel + trainer tests, A/B timing, ncu --set full of the separable form
TAG=${1:-r2z4}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 python -m pytest tests -q -m gpu -rf > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/${TAG}_tests.log)"
grep -E "^FAILED|^ERROR" gpurun_out/${TAG}_tests.log | head -20
python - > gpurun_out/${TAG}_gauss_ab.txt 2>&1 <<'P'
import os, torch
from pose_estimation_amitai_b200 import ops
dev = torch.device("cuda")
pts = torch.randint(8, 184, (64, 36, 2), device=dev).float()
def timed(fn, iters=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
nbytes = 64 * 36 * 192 * 192 * 4
new = timed(lambda: ops.gaussian_heatmaps(pts)); a = ops.gaussian_heatmaps(pts)
os.environ["POSEB200_GAUSS_V1"] = "1"
old = timed(lambda: ops.gaussian_heatmaps(pts)); b = ops.gaussian_heatmaps(pts)
print(f"gaussian render 64 x 36 x 192^2: per-pixel expf {old:.1f} us ({nbytes / old / 1e3:.0f} GB/s) -> separable {new:.1f} us "
      f"({nbytes / new / 1e3:.0f} GB/s); max rel diff {((a - b).abs() / b.clamp_min(1e-30)).max().item():.2e}")
P
cat gpurun_out/${TAG}_gauss_ab.txt | tail -3
python tools/bw_prof.py gauss > gpurun_out/plain_gauss.log 2>&1; echo "plain rc=$?"
timeout 240 $NCU -k regex:gaussian -s 1 -c 1 -o gpurun_out/${TAG}_gauss python tools/bw_prof.py gauss > gpurun_out/ncu_gauss.log 2>&1; echo "ncu rc=$?"
