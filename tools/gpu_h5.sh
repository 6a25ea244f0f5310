mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "attention" > gpurun_out/h5_pytest_attn.log 2>&1; echo "attn rc=$?"; tail -30 gpurun_out/h5_pytest_attn.log
if grep -q "passed" gpurun_out/h5_pytest_attn.log && ! grep -q "failed" gpurun_out/h5_pytest_attn.log; then
timeout 400 python -m pytest tests -q -m gpu -x > gpurun_out/h5_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/h5_pytest.log
timeout 300 python bench.py --model vit --steps 10 --warmup 3 --no-bandwidth --no-cpu-baseline > gpurun_out/h5_bench_vit.log 2>&1; echo "vit rc=$?"; grep -h '^{' gpurun_out/h5_bench_vit.log | cut -c1-260
POSEB200_ATTN_UNFUSED=1 timeout 300 python bench.py --model vit --steps 10 --warmup 3 --no-bandwidth --no-cpu-baseline --no-inference > gpurun_out/h5_bench_vit_unfused.log 2>&1; echo "vit unfused rc=$?"; grep -h '^{' gpurun_out/h5_bench_vit_unfused.log | cut -c1-260
fi
