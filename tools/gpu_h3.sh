mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_fourcam.py tests/test_augment.py -q -m gpu -x -k "tc_wgrad or tc_layer_fwd_dgrad or fourcam or affine" > gpurun_out/h3_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/h3_pytest.log
timeout 300 python bench.py --model fourcam --steps 5 --warmup 3 --no-bandwidth --no-cpu-baseline > gpurun_out/h3_bench_fourcam.log 2>&1; echo "bench rc=$?"; tail -3 gpurun_out/h3_bench_fourcam.log | cut -c1-3000
