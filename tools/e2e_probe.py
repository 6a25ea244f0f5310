import sys, os, json
sys.path.insert(0, os.getcwd())
import bench, torch, argparse
args = argparse.Namespace(gpus=1)
ctx = bench._Ctx(args)
leg = bench.TrainLeg(ctx, "cnn", "bf16", 64, 36, graph=True)
for name in ["resident", "e2e", "resident", "e2e", "resident"]:
    if name == "resident":
        m = leg.measure(20, 3)
        print(name, round(m["value"]), round(m["ms_per_step"], 3), flush=True)
    else:
        e = leg.measure_e2e(20)
        print(name, round(e["value"]), round(e["ms_per_step"], 3), flush=True)
for probe in ["noh2d", "nod2h"]:
    os.environ["POSEB200_E2E_PROBE"] = probe
    e = leg.measure_e2e(20)
    print("e2e", probe, round(e["value"]), round(e["ms_per_step"], 3), flush=True)
