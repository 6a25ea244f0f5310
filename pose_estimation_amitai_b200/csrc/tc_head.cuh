// Fused network head: what the last layer's epilogue does with its heatmaps (shared by tc_conv2.cu and tc_head.cu).
#pragma once

#include "tc_common.cuh"

namespace pb {

struct V2Head {   // host-side carrier
  int mode;       // 0 (or no V2Head): store NCHW fp32 heatmaps; 1: per-map arg-max keys; 2: MSE loss + bf16 NHWC gradient
  unsigned long long* keys;
  const float* target;
  const float* points;
  float negk2, gscale;
  float* loss;
  void* grad;
  int cpad;
  float* dbias;   // mode 2: bias gradient accumulated by the epilogue (tc_head.cu only), or nullptr
};

// tc_head.cu: the stride-2 transposed 3x3 head with its four output parities folded into the MMA's N.
// PB_ERR_UNSUPPORTED: shape outside its tiling (the caller falls back to the generic halo kernel).
int head_tc(const pb_conv_args* a, const V2Head* head, cudaStream_t stream);

}  // namespace pb
