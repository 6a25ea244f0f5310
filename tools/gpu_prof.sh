#!/bin/bash
mkdir -p gpurun_out
for L in conv2 conv8; do
CMD="python tools/wgrad_bench.py --batch 64 --iters 1 --only $L"
$CMD > gpurun_out/prof_plain_$L.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tc_wgrad2_kernel -c 1 -o gpurun_out/r1_wgrad2_$L -f $CMD > gpurun_out/ncu_wgrad2_$L.log 2>&1
echo "ncu $L rc=$?"
done
