#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/conv_bench.py --batch 3 --iters 2 --mode default > gpurun_out/triage.log 2>&1
echo "triage rc=$?"
timeout 600 python tools/conv_bench.py --batch 64 --mode default > gpurun_out/convbench.log 2>&1
echo "convbench rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v2.log 2>&1
echo "pytest v2 rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v2.log 2>&1
echo "bench v2 rc=$?"
grep -v " v1 " gpurun_out/triage.log | awk '{printf "%-8s %-6s %-10s %s %8s us\n",$1,$2,$3,$5,$6}'; tail -3 gpurun_out/pytest_v2.log
grep -v " v1 " gpurun_out/convbench.log | awk '{printf "%-8s %-6s %-10s %s %8s us %8s TF\n",$1,$2,$3,$5,$6,$8}'
