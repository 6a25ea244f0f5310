#!/bin/bash
# GPU passes of the last session of round 2 (the HBM-bound peak / target / gather kernels).  Each ncu command runs
# only after the same command exited 0 without ncu.   usage: tools/gpu_r2z.sh <tag> [all|tests|ab|prof]
TAG=${1:-r2z}; WHAT=${2:-all}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
if [[ $WHAT == all || $WHAT == tests ]]; then
  timeout 900 python -m pytest tests -q -m gpu -rf > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/${TAG}_tests.log)"
  grep -E "^FAILED|^ERROR" gpurun_out/${TAG}_tests.log | head -20
  timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
fi
if [[ $WHAT == all || $WHAT == ab ]]; then
  # arg-max / soft arg-max: old and new form in one process, results compared
  timeout 200 python tools/peaks_ab.py > gpurun_out/${TAG}_peaks_ab.txt 2>&1; echo "peaks_ab rc=$?"; tail -6 gpurun_out/${TAG}_peaks_ab.txt
  # u8 affine gather: four pixels per thread (default) vs one
  python tools/bw_prof.py affine_u8 | tail -2; POSEB200_AFFINE_PX1=1 python tools/bw_prof.py affine_u8 | tail -2
  ( time timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err ); echo "bench rc=$?"
  tail -3 gpurun_out/${TAG}_bench.err; cut -c1-300 gpurun_out/${TAG}_bench.json
  timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2>/dev/null; echo "reference arm rc=$?"
fi
if [[ $WHAT == all || $WHAT == prof ]]; then
  run() {  # name, kernel regex, launch-skip, launch-count, command...
    local name=$1 kern=$2 skip=$3 cnt=$4; shift 4
    "$@" > gpurun_out/plain_$name.log 2>&1 || { echo "plain $name FAILED"; return; }
    timeout 240 $NCU -k regex:$kern -s $skip -c $cnt -o gpurun_out/${TAG}_$name "$@" > gpurun_out/ncu_$name.log 2>&1
    echo "ncu $name rc=$?"
  }
  run argmax_bf16 argmax_planar 1 1 python tools/bw_prof.py argmax_bf16
  run argmax_f32 argmax_planar 1 1 python tools/bw_prof.py argmax
  run softargmax softargmax_kernel 1 1 python tools/bw_prof.py softargmax
  run gauss gaussian 1 1 python tools/bw_prof.py gauss
  run affine_u8 affine_nearest 2 1 python tools/bw_prof.py affine_u8
  run attn tc_attn_kernel 3 3 python tools/bw_prof.py attn
fi
