mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu -x > gpurun_out/h4_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/h4_pytest.log
timeout 300 python bench.py --model vit --steps 10 --warmup 3 --no-bandwidth --no-cpu-baseline --no-inference > gpurun_out/h4_bench_vit.log 2>&1; echo "vit rc=$?"; grep -h '^{' gpurun_out/h4_bench_vit.log | cut -c1-260
timeout 300 python bench.py --steps 10 --warmup 3 --no-bandwidth --no-cpu-baseline --no-inference > gpurun_out/h4_bench_cnn.log 2>&1; echo "cnn rc=$?"; grep -h '^{' gpurun_out/h4_bench_cnn.log | cut -c1-260
