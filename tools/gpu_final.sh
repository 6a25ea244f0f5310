#!/bin/bash
# final state of the round: full GPU test run, smoke, default bench line
TAG=${1:-r2zz}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -rf > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/${TAG}_tests.log)"
grep -E "^FAILED|^ERROR" gpurun_out/${TAG}_tests.log | head -20
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
( time timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err ); echo "bench rc=$?"
tail -3 gpurun_out/${TAG}_bench.err; cut -c1-200 gpurun_out/${TAG}_bench.json
