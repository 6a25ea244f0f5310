mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu -x > gpurun_out/h8_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/h8_pytest.log
if grep -q " passed" gpurun_out/h8_pytest.log && ! grep -q "failed" gpurun_out/h8_pytest.log; then
timeout 300 python bench.py --model vit --steps 10 --warmup 3 --no-bandwidth --no-cpu-baseline > gpurun_out/h8_bench_vit.log 2>&1; echo "vit rc=$?"; grep -h '^{' gpurun_out/h8_bench_vit.log | cut -c1-260
timeout 120 python tools/bw_prof.py attn > gpurun_out/plain_attn2.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -f -k regex:tc_attn_kernel -s 3 -c 3 -o gpurun_out/r1j_attn python tools/bw_prof.py attn > gpurun_out/ncu_attn2.log 2>&1
echo "ncu attn rc=$?"
timeout 400 python bench.py > gpurun_out/h8_bench_cnn_default.log 2>&1; echo "cnn rc=$?"; grep -h '^{' gpurun_out/h8_bench_cnn_default.log | cut -c1-400
fi
