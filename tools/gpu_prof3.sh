#!/bin/bash
# session-4 profile pass: every ncu command runs only after the same command exited 0 without ncu
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 120 python tools/bw_prof.py attn > gpurun_out/plain_attn.log 2>&1 && \
timeout 300 $NCU -k regex:tc_attn_kernel -s 3 -c 3 -o gpurun_out/r1i_attn python tools/bw_prof.py attn > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn rc=$?"
timeout 120 python tools/bw_prof.py affine > gpurun_out/plain_affine.log 2>&1 && \
timeout 300 $NCU -k regex:affine_nearest_kernel -s 1 -c 1 -o gpurun_out/r1i_affine python tools/bw_prof.py affine > gpurun_out/ncu_affine.log 2>&1
echo "ncu affine rc=$?"
timeout 200 python bench.py --model vit --steps 2 --warmup 1 --no-cpu-baseline --no-bandwidth --no-inference > gpurun_out/plain_vit2.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r1i_vit_launches.csv python bench.py --model vit --steps 2 --warmup 1 --no-cpu-baseline --no-bandwidth --no-inference > gpurun_out/ncu_vit2.log 2>&1
echo "vit launch list rc=$?"
