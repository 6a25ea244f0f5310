#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_8gpu.log 2>&1
echo "bench 8gpu rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_4gpu.log 2>&1
echo "bench 4gpu rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 tools/infer_sweep.py --batches 256 1024 4096 > gpurun_out/infer_sweep_8gpu.log 2>&1
echo "sweep 8gpu rc=$?"
grep '^{' gpurun_out/bench_8gpu.log | cut -c1-200; grep '^{' gpurun_out/bench_4gpu.log | cut -c1-200; grep '^{' gpurun_out/infer_sweep_8gpu.log
