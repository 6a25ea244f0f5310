#!/bin/bash
mkdir -p gpurun_out
for L in "conv2 fwd" "conv5 fwd" "conv8 fwd"; do
T=$(echo $L | tr -d ' ')
python tools/conv_bench.py --batch 64 --iters 1 --mode default --only "$L" > gpurun_out/prof_plain_$T.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tc_conv2_kernel -c 1 -o gpurun_out/r1c_conv2k_$T -f python tools/conv_bench.py --batch 64 --iters 1 --mode default --only "$L" > gpurun_out/ncu_$T.log 2>&1
echo "ncu $T rc=$?"
done
