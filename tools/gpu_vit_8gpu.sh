mkdir -p gpurun_out
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --model vit --gpus 8 --steps 10 --warmup 3 --no-bandwidth --no-cpu-baseline --no-inference > gpurun_out/h11_bench_vit_8gpu.log 2>&1; echo "vit 8gpu rc=$?"; grep -h '^{' gpurun_out/h11_bench_vit_8gpu.log | cut -c1-300
