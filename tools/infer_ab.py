#!/usr/bin/env python
"""A/B of inference-path knobs on one box: frames/s of BasicNet.predict_peaks (256 frames) with each env setting."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from bench import CFG, IMG, JOINTS
from pose_estimation_amitai_b200 import CNNs

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = CNNs.BasicNet(dict(CFG), np.array((IMG, IMG, 4)), JOINTS).to(dev)
x = torch.rand(256, 4, IMG, IMG, device=dev)
cases = [("default", {}), ("no pool fusion", {"POSEB200_NO_POOL_FUSION": "1"}), ("no conv1 direct", {"POSEB200_NO_CONV1_DIRECT": "1"}),
         ("generic head", {"POSEB200_HEAD_V2": "0"}), ("default", {})]
for name, env in cases * 2:
    for k in ("POSEB200_NO_POOL_FUSION", "POSEB200_NO_CONV1_DIRECT", "POSEB200_HEAD_V2"):
        os.environ.pop(k, None)
    os.environ.update(env)
    for _ in range(3):
        model.predict_peaks(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        model.predict_peaks(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name:16s} {ms:7.3f} ms  {256 / ms * 1e3:8.0f} frames/s", flush=True)
