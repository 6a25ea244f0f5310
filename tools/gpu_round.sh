#!/bin/bash
# one GPU-box visit: per-layer timings + cross-check, GPU tests, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 300 python tools/conv_bench.py --batch 4 --iters 2 --mode default strips > gpurun_out/triage.log 2>&1
echo "triage rc=$?"
timeout 600 python tools/conv_bench.py --batch 64 --mode default halo_T1 halo_T2 halo_T4 nostage > gpurun_out/convbench.log 2>&1
echo "convbench rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v2.log 2>&1
echo "pytest v2 rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v2.log 2>&1
echo "bench v2 rc=$?"
grep -v " v1 " gpurun_out/triage.log
