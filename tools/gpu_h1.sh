mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_augment.py -q -m gpu -x > gpurun_out/h1_pytest_aug.log 2>&1; echo "aug rc=$?"; tail -15 gpurun_out/h1_pytest_aug.log
timeout 120 python - > gpurun_out/h1_affbench.log 2>&1 <<'PY'
import json, torch, bench
dev = torch.device("cuda:0")
peaks = json.load(open("MEASURED_PEAKS.json"))
hbm = peaks.get("hbm_gbs") or 6546.0
for r in bench.bandwidth_kernels(dev, hbm)[:2]:
    print(json.dumps(r))
PY
echo "affbench rc=$?"; tail -5 gpurun_out/h1_affbench.log
