// Shared helpers for libposeb200 (sm_100a).  Error plumbing, dtype load/store, reductions.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/poseb200.h"

namespace pb {

// ---- error plumbing (thread local message, negative status codes; no exceptions cross the ABI)
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
bool is_device_ptr(const void* p);

#define PB_REQUIRE(cond, ...)                 \
  do {                                        \
    if (!(cond)) {                            \
      pb::set_error(__VA_ARGS__);             \
      return PB_ERR_INVALID;                  \
    }                                         \
  } while (0)

#define PB_REQUIRE_DEV(ptr, name)                                        \
  do {                                                                   \
    if ((ptr) != nullptr && !pb::is_device_ptr(ptr)) {                   \
      pb::set_error("%s: '%s' is not a device pointer (no CPU fallback)", __func__, name); \
      return PB_ERR_NOT_DEVICE;                                          \
    }                                                                    \
  } while (0)

// every kernel launch of the library is counted (pb_launch_count(); bench.py's gpu_launches)
void note_launches(int n);

#define PB_LAUNCH_CHECK(what)                              \
  do {                                                     \
    cudaError_t e__ = cudaGetLastError();                  \
    if (e__ != cudaSuccess) return pb::cuda_fail(e__, what); \
    pb::note_launches(1);                                  \
  } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
int sm_count();

// ---- dtype helpers
template <typename T>
__device__ __forceinline__ float ldf(const T* p, long long i);
template <>
__device__ __forceinline__ float ldf<float>(const float* p, long long i) { return p[i]; }
template <>
__device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p, long long i) {
  return __bfloat162float(p[i]);
}
template <typename T>
__device__ __forceinline__ void stf(T* p, long long i, float v);
template <>
__device__ __forceinline__ void stf<float>(float* p, long long i, float v) { p[i] = v; }
template <>
__device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, long long i, float v) {
  p[i] = __float2bfloat16_rn(v);
}

template <>
__device__ __forceinline__ float ldf<__half>(const __half* p, long long i) { return __half2float(p[i]); }
template <>
__device__ __forceinline__ void stf<__half>(__half* p, long long i, float v) { p[i] = __float2half_rn(v); }

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : slope * v; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum, result valid in thread 0 (blockDim.x multiple of 32, <= 1024)
__device__ __forceinline__ float block_sum(float v, float* smem32) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) smem32[w] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? smem32[threadIdx.x] : 0.f;
  if (w == 0) v = warp_sum(v);
  __syncthreads();
  return v;
}

// 16-byte streaming load / store (read-once data: skip L1 allocation)
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream16(void* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w));
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float bf16lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// arg-max keys: (order-preserving float bits << 32) | (~flat_index), see bandwidth.cu "Peaks"
__device__ __forceinline__ uint32_t order_key(float v) {
  if (v != v) return 0xFFFFFFFFu;  // NaN is the maximum (torch.max semantics, Augmentor.py:131)
  v = v + 0.0f;                    // -0 -> +0 so they tie
  const uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
  if (k == 0xFFFFFFFFu) return __uint_as_float(0x7FC00000u);
  const uint32_t u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
  return __uint_as_float(u);
}
__device__ __forceinline__ unsigned long long make_key(float v, uint32_t idx) {
  return ((unsigned long long)order_key(v) << 32) | (unsigned long long)(0xFFFFFFFFu - idx);
}

// fp16 twins ("fp16" precision: forward operands in IEEE half -- 11 significand bits against bf16's 8 -- while
// gradients stay bf16 for their exponent range)
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  __half2 t = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float f16lo(uint32_t v) { return __half2float(__ushort_as_half((unsigned short)(v & 0xffffu))); }
__device__ __forceinline__ float f16hi(uint32_t v) { return __half2float(__ushort_as_half((unsigned short)(v >> 16))); }

// 16-bit storage format chosen at compile time: F16 = IEEE half, otherwise bf16
template <bool F16>
__device__ __forceinline__ uint32_t pack16x2(float lo, float hi) { return F16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi); }
template <bool F16>
__device__ __forceinline__ float lo16(uint32_t v) { return F16 ? f16lo(v) : bf16lo(v); }
template <bool F16>
__device__ __forceinline__ float hi16(uint32_t v) { return F16 ? f16hi(v) : bf16hi(v); }

}  // namespace pb
