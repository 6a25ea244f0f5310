#!/bin/bash
# compute-sanitizer over a reduced `-m gpu` set (GPU box): memcheck, racecheck (shared-memory hazards of the
# warp-specialised kernels), synccheck.  Each tool runs under its own timeout; logs -> gpurun_out/r2_sanitize_<tool>.log,
# the error summaries are collected into gpurun_out/r2_sanitize_summary.txt (copied to profiles/ by hand).
#   usage: tools/sanitize.sh [per-tool timeout seconds, default 900]
T=${1:-240}
mkdir -p gpurun_out
SAN=$(command -v compute-sanitizer || echo /usr/local/cuda/bin/compute-sanitizer)
# small shapes of every kernel family: tcgen05 conv forward / input gradient / pairs / fused head, both weight-gradient
# kernels, fused attention, the bandwidth kernels, the optimiser
SEL='test_head_folded_parities_vs_torch_and_generic_kernel or test_head_argmax_fused_is_bit_exact or test_head_mse_fused_equals_head_then_mse or test_first_layer_direct_vs_im2col_form_and_torch or test_tc_pair_ragged_group_counts or test_tc_wgrad or test_mse_loss_and_grad or test_argmax_kat_nhwc_and_nchw or test_adam_matches_oracle or test_tcgen05_selftest_gemm'
: > gpurun_out/r2_sanitize_summary.txt
for tool in memcheck racecheck synccheck; do
  log=gpurun_out/r2_sanitize_$tool.log
  timeout $T $SAN --tool $tool --target-processes all --launch-timeout 0 --error-exitcode 0 --print-limit 20 \
      python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "$SEL" -p no:cacheprovider > $log 2>&1
  rc=$?
  {
    echo "== $tool rc=$rc ($(grep -c 'Error\|Hazard\|Race' $log) matching lines)"
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" $log | tail -8
  } >> gpurun_out/r2_sanitize_summary.txt
done
cat gpurun_out/r2_sanitize_summary.txt
