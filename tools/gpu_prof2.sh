#!/bin/bash
# round-1 (session 3) profile pass: every ncu command runs only after the same command exited 0 without ncu
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -c 1 -f"
run() {  # name, kernel regex, launch-skip, command...
  local name=$1 kern=$2 skip=$3; shift 3
  "$@" > gpurun_out/plain_$name.log 2>&1 || { echo "plain $name FAILED"; return; }
  timeout 300 $NCU -k regex:$kern -s $skip -o gpurun_out/r1g_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "ncu $name rc=$?"
}
run conv5fwd_pair tc_conv2_kernel 3 python tools/conv_bench.py --batch 64 --iters 1 --mode default --only "conv5 fwd"
run conv8fwd_pair tc_conv2_kernel 3 python tools/conv_bench.py --batch 64 --iters 1 --mode default --only "conv8 fwd"
run conv2fwd_pair tc_conv2_kernel 3 python tools/conv_bench.py --batch 64 --iters 1 --mode default --only "conv2 fwd"
run conv2dgrad_pair tc_conv2_kernel 3 python tools/conv_bench.py --batch 64 --iters 1 --mode default --only "conv2 dgrad"
run wgrad2_conv5 tc_wgrad2_kernel 2 python tools/wgrad_bench.py --batch 64 --iters 1 --only conv5
run mse mse_nhwc_bf16_kernel 1 python tools/bw_prof.py mse

run argmax argmax_planar_kernel 1 python tools/bw_prof.py argmax



timeout 120 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-bandwidth > gpurun_out/plain_bench2.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r1g_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-bandwidth > gpurun_out/ncu_bench2.log 2>&1
echo "launch list rc=$?"
