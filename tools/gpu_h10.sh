mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_network.py -q -m gpu -x -k "attention or vit" > gpurun_out/h10_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/h10_pytest.log
if grep -q " passed" gpurun_out/h10_pytest.log && ! grep -q "failed" gpurun_out/h10_pytest.log; then
timeout 300 python bench.py --model vit --steps 10 --warmup 3 --no-bandwidth --no-cpu-baseline --no-inference > gpurun_out/h10_bench_vit.log 2>&1; echo "vit rc=$?"; grep -h '^{' gpurun_out/h10_bench_vit.log | cut -c1-200
fi
