#!/usr/bin/env python
"""Timing experiment (GPU box only): the training step replayed from a CUDA graph vs launched eagerly.  The captured
Adam launch freezes its bias-correction scalars, so this measures launch gaps only -- it is not a training path."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from bench import CFG, IMG, JOINTS
from pose_estimation_amitai_b200 import CNNs, parallel

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = CNNs.BasicNet(dict(CFG), np.array((IMG, IMG, 4)), JOINTS).to(dev)
dp = parallel.DataParallelStep(model, lr=1e-3)
x = torch.rand(64, 4, IMG, IMG, device=dev)
pts = torch.randint(8, IMG - 8, (64, JOINTS, 2), device=dev).float()


def timed(fn, n=20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for _ in range(5):
    dp.step(x, points=pts)
print("eager  ms/step", timed(lambda: dp.step(x, points=pts)))
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        dp.step(x, points=pts)
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    loss = dp.step(x, points=pts)
g.replay()
torch.cuda.synchronize()
print("graph  ms/step", timed(g.replay), "loss", loss.item())

# host side: how long does Python need to ENQUEUE one eager step (the GPU is kept busy behind a long spin, so the
# wall clock below is pure host work)?
import time
torch.cuda.synchronize()
torch.cuda._sleep(int(2e9))
t0 = time.perf_counter()
for _ in range(20):
    dp.step(x, points=pts)
t1 = time.perf_counter()
torch.cuda.synchronize()
print("host enqueue ms/step", (t1 - t0) * 1e3 / 20)
