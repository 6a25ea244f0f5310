#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v2.log 2>&1
echo "pytest rc=$?"
tail -n 25 gpurun_out/pytest_v2.log
timeout 400 python bench.py --model vit --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_vit.log 2>&1
echo "bench vit rc=$?"
tail -n 1 gpurun_out/bench_vit.log | cut -c1-200
