#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/conv_bench.py --batch 64 --iters 1 --mode halo_T2 --only conv2"
$CMD > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tc_conv2_kernel -c 2 -o gpurun_out/r1_conv2_halo -f $CMD > gpurun_out/ncu_conv2.log 2>&1
echo "ncu conv2 rc=$?"
CMD="python tools/conv_bench.py --batch 64 --iters 1 --mode halo_T2 --only conv8"
$CMD > gpurun_out/prof_plain8.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tc_conv2_kernel -c 2 -o gpurun_out/r1_conv8_halo -f $CMD > gpurun_out/ncu_conv8.log 2>&1
echo "ncu conv8 rc=$?"
