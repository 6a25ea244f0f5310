mkdir -p gpurun_out
timeout 200 python -m pytest tests -q -m gpu -x > gpurun_out/h12_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/h12_pytest.log
timeout 100 python __graft_entry__.py --smoke > gpurun_out/h12_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/h12_smoke.log
