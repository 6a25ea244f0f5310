#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/wgrad_bench.py --batch 4 --iters 2 > gpurun_out/wgrad_triage.log 2>&1
echo "wgrad triage rc=$?"
timeout 600 python tools/wgrad_bench.py --batch 64 > gpurun_out/wgradbench.log 2>&1
echo "wgradbench rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v2.log 2>&1
echo "pytest v2 rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v2.log 2>&1
echo "bench v2 rc=$?"
cat gpurun_out/wgradbench.log; tail -3 gpurun_out/pytest_v2.log
