#!/bin/bash
# follow-up GPU pass: tests, peak-kernel A/B, ncu --set full of the arg-max scan in both element types
TAG=${1:-r2z2}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 900 python -m pytest tests -q -m gpu -rf > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/${TAG}_tests.log)"
grep -E "^FAILED|^ERROR" gpurun_out/${TAG}_tests.log | head -20
timeout 200 python tools/peaks_ab.py > gpurun_out/${TAG}_peaks_ab.txt 2>&1; echo "peaks_ab rc=$?"; cat gpurun_out/${TAG}_peaks_ab.txt | tail -6
run() {  # name, kernel regex, launch-skip, launch-count, command...
  local name=$1 kern=$2 skip=$3 cnt=$4; shift 4
  "$@" > gpurun_out/plain_$name.log 2>&1 || { echo "plain $name FAILED"; return; }
  timeout 240 $NCU -k regex:$kern -s $skip -c $cnt -o gpurun_out/${TAG}_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "ncu $name rc=$?"
}
run argmax_bf16 argmax_planar 1 1 python tools/bw_prof.py argmax_bf16
run argmax_f32 argmax_planar 1 1 python tools/bw_prof.py argmax
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
