#!/bin/bash
# multi-GPU pass: weak scaling at N GPUs (64 samples per GPU), the strong-scaling point (global batch 64), inference sweep
N=${1:-8}
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29521 bench.py --gpus $N --steps 10 --warmup 3 --no-bandwidth > gpurun_out/bench_${N}gpu.log 2>&1; echo "bench ${N}gpu rc=$?"
run 29522 bench.py --gpus $N --steps 10 --warmup 3 --no-bandwidth --no-inference --batch-per-gpu $((64 / N)) > gpurun_out/bench_${N}gpu_strong.log 2>&1; echo "strong ${N}gpu rc=$?"
run 29523 tools/infer_sweep.py --batches 256 1024 4096 > gpurun_out/infer_sweep_${N}gpu.log 2>&1; echo "sweep ${N}gpu rc=$?"
grep -h '^{' gpurun_out/bench_${N}gpu.log gpurun_out/bench_${N}gpu_strong.log | cut -c1-160; grep -h '^{' gpurun_out/infer_sweep_${N}gpu.log
