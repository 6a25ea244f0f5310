mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu -x > gpurun_out/h9_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/h9_pytest.log
if grep -q " passed" gpurun_out/h9_pytest.log && ! grep -q "failed" gpurun_out/h9_pytest.log; then
timeout 300 python bench.py --model vit --steps 10 --warmup 3 --no-bandwidth --no-cpu-baseline --no-inference > gpurun_out/h9_bench_vit.log 2>&1; echo "vit rc=$?"; grep -h '^{' gpurun_out/h9_bench_vit.log | cut -c1-200
timeout 300 python bench.py --steps 10 --warmup 3 --no-bandwidth --no-cpu-baseline --no-inference > gpurun_out/h9_bench_cnn.log 2>&1; echo "cnn rc=$?"; grep -h '^{' gpurun_out/h9_bench_cnn.log | cut -c1-200
fi
