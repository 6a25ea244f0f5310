mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "attention" > gpurun_out/h6_pytest_attn.log 2>&1; echo "attn rc=$?"; tail -12 gpurun_out/h6_pytest_attn.log
if grep -q "passed" gpurun_out/h6_pytest_attn.log && ! grep -q "failed" gpurun_out/h6_pytest_attn.log; then
timeout 300 python -m pytest tests/test_gpu_network.py -q -m gpu -x -k "vit" > gpurun_out/h6_pytest_vit.log 2>&1; echo "vit tests rc=$?"; tail -3 gpurun_out/h6_pytest_vit.log
timeout 300 python bench.py --model vit --steps 10 --warmup 3 --no-bandwidth --no-cpu-baseline --no-inference > gpurun_out/h6_bench_vit.log 2>&1; echo "vit rc=$?"; grep -h '^{' gpurun_out/h6_bench_vit.log | cut -c1-260
fi
