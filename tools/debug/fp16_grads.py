import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import pose_oracle as po
from pose_estimation_amitai_b200 import CNNs
CFG = {"model type": "MODEL_18_POINTS_PER_WING", "number of base filters": 64, "convolution kernel size": 3,
       "dilation rate": 2, "dropout ratio": 0.5}
x = po.synthetic_crops(2, seed=1).cuda()
pts = po.synthetic_points(2, 36, seed=2)
tgt = torch.from_numpy(po.gaussian_targets(pts)).cuda()
res = {}
for prec in ("bf16", "fp16"):
    for path in ("autograd", "fused"):
        torch.manual_seed(0)
        m = CNNs.BasicNet(dict(CFG, precision=prec), np.array((192, 192, 4)), 36).cuda().train()
        if path == "autograd":
            loss = torch.nn.MSELoss()(m(x), tgt); loss.backward()
        else:
            m.train_step(x, tgt)
        res[(prec, path)] = {k: p.grad.double().norm().item() for k, p in m.named_parameters() if p.grad is not None}
base = res[("bf16", "fused")]
for k in base:
    print(f"{k:36s} " + " ".join(f"{res[key][k] / base[k]:8.4f}" for key in res))
