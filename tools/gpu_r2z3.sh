#!/bin/bash
# affine gather pass: augmentation tests, timing + ncu --set full of the u8 form, bandwidth table of bench.py
TAG=${1:-r2z3}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 python -m pytest tests/test_augment.py tests/test_gpu_kernels.py -q -m gpu -rf > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/${TAG}_tests.log)"
grep -E "^FAILED|^ERROR" gpurun_out/${TAG}_tests.log | head -20
python tools/bw_prof.py affine_u8 > gpurun_out/plain_affine_u8.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/plain_affine_u8.log
timeout 240 $NCU -k regex:affine_nearest -s 2 -c 1 -o gpurun_out/${TAG}_affine_u8 python tools/bw_prof.py affine_u8 > gpurun_out/ncu_affine_u8.log 2>&1; echo "ncu rc=$?"
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-inference > gpurun_out/${TAG}_bench_short.json 2> gpurun_out/${TAG}_bench_short.err; echo "bench rc=$?"
python - <<'P'
import json
d = json.load(open("gpurun_out/r2z3_bench_short.json"))
for k in d["bandwidth_kernels"]["kernels"]:
    print(k["kernel"][:70], k["us"], k["frac_of_hbm_peak"])
P
