#!/bin/bash
# round-2 GPU pass: tests, default bench, launch list, ncu --set full of the 64-channel layers (each ncu command only
# after the same command exited 0 without ncu).   usage: tools/gpu_r2.sh <tag> [tests|bench|prof|all]
TAG=${1:-r2h}; WHAT=${2:-all}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -c 1 -f"
if [[ $WHAT == all || $WHAT == tests ]]; then
  timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/${TAG}_tests.log)"
  timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/${TAG}_smoke.log
fi
if [[ $WHAT == all || $WHAT == bench ]]; then
  ( time timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err ); echo "bench rc=$?"
  tail -3 gpurun_out/${TAG}_bench.err; cut -c1-400 gpurun_out/${TAG}_bench.json
fi
if [[ $WHAT == all || $WHAT == prof ]]; then
  timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-bandwidth --no-extras > gpurun_out/${TAG}_plain2.log 2>&1 && \
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-bandwidth --no-extras > gpurun_out/${TAG}_ncu_bench2.log 2>&1
  echo "launch list rc=$?"
  run() {  # name, kernel regex, launch-skip, command...
    local name=$1 kern=$2 skip=$3; shift 3
    "$@" > gpurun_out/plain_$name.log 2>&1 || { echo "plain $name FAILED"; return; }
    timeout 300 $NCU -k regex:$kern -s $skip -o gpurun_out/${TAG}_$name "$@" > gpurun_out/ncu_$name.log 2>&1
    echo "ncu $name rc=$?"
  }
  run conv2fwd tc_conv2_kernel 3 python tools/conv_bench.py --batch 64 --iters 1 --mode default --only "conv2 fself"
  run conv2dgrad tc_conv2_kernel 3 python tools/conv_bench.py --batch 64 --iters 1 --mode default --only "conv2 dgrad"
  run conv5fwd tc_conv2_kernel 3 python tools/conv_bench.py --batch 64 --iters 1 --mode default --only "conv5 fself"
  run conv8fwd tc_conv2_kernel 3 python tools/conv_bench.py --batch 64 --iters 1 --mode default --only "conv8 fwd"
  run wgrad2_conv5 tc_wgrad2_kernel 2 python tools/wgrad_bench.py --batch 64 --iters 1 --only conv5
fi
if [[ $WHAT == final ]]; then
  timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/${TAG}_tests.log)"
  timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
  timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2>/dev/null; echo "reference arm rc=$?"
  ( time timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err ); echo "bench rc=$?"
  tail -3 gpurun_out/${TAG}_bench.err; cut -c1-300 gpurun_out/${TAG}_bench.json
  timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-bandwidth --no-extras --no-graph > gpurun_out/${TAG}_plain2.log 2>&1 && \
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-bandwidth --no-extras --no-graph > gpurun_out/${TAG}_ncu_bench2.log 2>&1
  echo "launch list rc=$?"
  run() {  # name, kernel regex, launch-skip, command...
    local name=$1 kern=$2 skip=$3; shift 3
    "$@" > gpurun_out/plain_$name.log 2>&1 || { echo "plain $name FAILED"; return; }
    timeout 300 $NCU -k regex:$kern -s $skip -o gpurun_out/${TAG}_$name "$@" > gpurun_out/ncu_$name.log 2>&1
    echo "ncu $name rc=$?"
  }
  run head_store tc_head_kernel 3 python tools/head_bench.py --only store --iters 2
  run head_argmax tc_head_kernel 3 python tools/head_bench.py --only argmax --iters 2
  run head_mse tc_head_kernel 3 python tools/head_bench.py --only mse --iters 2
  run conv1 tc_conv1_kernel 3 python tools/bw_prof.py conv1
  run conv2fwd tc_conv2_kernel 3 python tools/conv_bench.py --batch 64 --iters 1 --mode default --only "conv2 fself"
  run conv2dgrad tc_conv2_kernel 3 python tools/conv_bench.py --batch 64 --iters 1 --mode default --only "conv2 dgrad"
  cat gpurun_out/plain_head_*.log gpurun_out/plain_conv1.log | grep -v "^$" | tail -12
fi
