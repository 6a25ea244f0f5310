mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_augment.py tests/test_fourcam.py -q -m gpu -x > gpurun_out/h2_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/h2_pytest.log
timeout 120 python - > gpurun_out/h2_affbench.log 2>&1 <<'PY'
import json, torch, bench
dev = torch.device("cuda:0")
hbm = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs") or 6546.0
for r in bench.bandwidth_kernels(dev, hbm)[:2]:
    print(json.dumps(r))
PY
echo "affbench rc=$?"; tail -5 gpurun_out/h2_affbench.log
