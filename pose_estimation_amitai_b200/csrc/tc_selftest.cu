// Single-tile tcgen05 GEMM used by the GPU tests to pin the UMMA descriptor encodings
// (K-major and MN-major shared-memory operands, 128-byte swizzle written by TMA).
//   D[128, N] (fp32) = A x B,  A: [M][K] (K-major) or [K][M] (MN-major), B: [N][K] or [K][N].
// Unpipelined on purpose: one TMA round trip and 4 MMAs per 64-wide K chunk.
#include <string.h>

#include "tc_common.cuh"

namespace pb {
using namespace tc;

struct SelfP {
  int N, K, a_mn, b_mn, a_f16, b_f16;
  float* d;
};

struct SelfMaps {
  CUtensorMap a, b;
};

__global__ void __launch_bounds__(128, 1)
tc_selftest_kernel(const __grid_constant__ SelfMaps maps, const SelfP p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t load_bar, mma_bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;
  uint8_t* sb = smem + 16384;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&load_bar, 1);
    mbar_init(&mma_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp == 0 && elect_one()) {
    const uint32_t idesc = make_idesc(128, p.N, p.a_mn, p.b_mn, p.a_f16, p.b_f16);
    const uint32_t bytes = 16384u + (uint32_t)p.N * 128u;
    uint32_t phase = 0;
    for (int kc = 0; kc < p.K / 64; ++kc) {
      mbar_expect_tx(&load_bar, bytes);
      if (p.a_mn) {
        tma_load_2d(sa, &maps.a, &load_bar, 0, kc * 64);
        tma_load_2d(sa + 8192, &maps.a, &load_bar, 64, kc * 64);
      } else {
        tma_load_2d(sa, &maps.a, &load_bar, kc * 64, 0);
      }
      if (p.b_mn) {
        for (int nb = 0; nb < p.N / 64; ++nb) tma_load_2d(sb + nb * 8192, &maps.b, &load_bar, nb * 64, kc * 64);
      } else {
        tma_load_2d(sb, &maps.b, &load_bar, kc * 64, 0);
      }
      mbar_wait(&load_bar, phase);
      tc_fence_after();
      for (int j = 0; j < 4; ++j) {
        const uint64_t ad = p.a_mn ? smem_desc_sw128(smem_u32(sa) + j * 2048, 8192, 1024)
                                   : smem_desc_sw128(smem_u32(sa) + j * 32, 16, 1024);
        const uint64_t bd = p.b_mn ? smem_desc_sw128(smem_u32(sb) + j * 2048, 8192, 1024)
                                   : smem_desc_sw128(smem_u32(sb) + j * 32, 16, 1024);
        umma_bf16(tmem_base, ad, bd, idesc, (kc > 0 || j > 0) ? 1u : 0u);
      }
      umma_commit(&mma_bar);
      mbar_wait(&mma_bar, phase);
      phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < p.N; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) p.d[(long long)row * p.N + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace pb

using namespace pb;

extern "C" int pb_gemm_selftest(const pb_gemm_selftest_args* a, void* stream) {
  PB_REQUIRE(a != nullptr && a->a && a->b && a->d, "pb_gemm_selftest: null args");
  PB_REQUIRE(a->M == 128 && a->N >= 16 && a->N <= 256 && a->N % 16 == 0 && a->K >= 64 && a->K % 64 == 0,
             "pb_gemm_selftest: M must be 128, N a multiple of 16 <= 256, K a multiple of 64");
  PB_REQUIRE(!a->b_mn_major || a->N % 64 == 0, "pb_gemm_selftest: MN-major B needs N %% 64 == 0");
  PB_REQUIRE_DEV(a->a, "a");
  PB_REQUIRE_DEV(a->b, "b");
  PB_REQUIRE_DEV(a->d, "d");
  SelfMaps maps;
  memset(&maps, 0, sizeof(maps));
  int rc;
  if (a->a_mn_major) {
    const uint64_t dims[2] = {(uint64_t)a->M, (uint64_t)a->K};
    const uint64_t str[1] = {(uint64_t)a->M * 2};
    const uint32_t box[2] = {64, 64};
    rc = encode_tmap_bf16(&maps.a, a->a, 2, dims, str, box);
  } else {
    const uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->M};
    const uint64_t str[1] = {(uint64_t)a->K * 2};
    const uint32_t box[2] = {64, 128};
    rc = encode_tmap_bf16(&maps.a, a->a, 2, dims, str, box);
  }
  if (rc != PB_OK) return rc;
  if (a->b_mn_major) {
    const uint64_t dims[2] = {(uint64_t)a->N, (uint64_t)a->K};
    const uint64_t str[1] = {(uint64_t)a->N * 2};
    const uint32_t box[2] = {64, 64};
    rc = encode_tmap_bf16(&maps.b, a->b, 2, dims, str, box);
  } else {
    const uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->N};
    const uint64_t str[1] = {(uint64_t)a->K * 2};
    const uint32_t box[2] = {64, (uint32_t)a->N};
    rc = encode_tmap_bf16(&maps.b, a->b, 2, dims, str, box);
  }
  if (rc != PB_OK) return rc;
  SelfP p;
  p.N = a->N; p.K = a->K; p.a_mn = a->a_mn_major; p.b_mn = a->b_mn_major; p.d = a->d;
  p.a_f16 = a->a_f16; p.b_f16 = a->b_f16;
  const size_t smem = 16384 + 32768 + 1024;
  cudaError_t e = cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "pb_gemm_selftest: smem attribute");
  tc_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(maps, p);
  PB_LAUNCH_CHECK("tc_selftest_kernel");
  return PB_OK;
}

