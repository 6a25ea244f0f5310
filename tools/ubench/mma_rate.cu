// Micro-benchmark (GPU box only): issue rate of tcgen05.mma kind::f16, M=128, cta_group::1, both operands in shared
// memory, as a function of N and of the shared-memory layout / swizzle mode of the operands.  No loads, no epilogue:
// the operands are whatever the shared memory holds.  Prints cycles per MMA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I pose_estimation_amitai_b200/csrc -I include tools/ubench/mma_rate.cu -o gpurun_out/mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace pb::tc;

__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

// MODE 0: SWIZZLE_128B K-major, k-step j at +32 B inside the 128-byte rows (what tc_conv2 does), SBO 1024
// MODE 1: SWIZZLE_32B  K-major, one [rows x 32 B] sub-tile per k-step, SBO 256
// MODE 4: SWIZZLE_128B with a halo-style SBO (20 columns x 128 B) and a 128-byte-aligned (not 1024) start
// MODE 5: MN-major A and B (weight-gradient kernels): SWIZZLE_128B, A tile rows 2048 B apart with its two 64-channel
//         halves 256 B apart (two taps of one halo), B 64-channel blocks 16 KB apart
// ROT: consecutive MMAs accumulate into ROT different accumulators (independent dependency chains); the loop is
// fully unrolled over 8 MMAs with compile-time descriptor offsets so the issuing thread is not the limiter.
template <int MODE, int ROT>
__global__ void __launch_bounds__(128, 1) k(int n, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1 && elect_one()) {
    const uint32_t idesc = make_idesc(128, n, MODE == 5 ? 1 : 0, MODE == 5 ? 1 : 0);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 96 * 1024;
    uint64_t ad[8], bd[8];
    uint32_t acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = u & 3, tap = u >> 2;
      if (MODE == 0) { ad[u] = desc(a0 + tap * 4096 + j * 32, 16, 1024, 2); bd[u] = desc(b0 + j * 32, 16, 1024, 2); }
      else if (MODE == 1) { ad[u] = desc(a0 + j * 16384 + tap * 1024, 16, 256, 6); bd[u] = desc(b0 + j * 8192, 16, 256, 6); }
      else if (MODE == 5) {   // MN-major operands as in tc_wgrad2: K = pixel rows of 128 B, 2 tile rows per K step
        ad[u] = desc(a0 + tap * 256 + j * 2 * 2048, 256, 2048, 2); bd[u] = desc(b0 + j * 2048, 16384, 1024, 2);
      }
      else { ad[u] = desc(a0 + tap * 5248 + j * 32, 16, 2560, 2); bd[u] = desc(b0 + j * 32, 16, 1024, 2); }
      acc[u] = tm + (uint32_t)((u % ROT) * n);
    }
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u) umma_bf16(acc[u], ad[u], bd[u], idesc, 1u);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int MODE, int ROT>
void run(const char* name, int n, long long* d) {
  if (ROT * n > 512) return;
  const int iters = 1000;
  cudaFuncSetAttribute(k<MODE, ROT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 2; ++rep) {
    k<MODE, ROT><<<148, 128, 200 * 1024>>>(n, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s n %d: %s\n", name, n, cudaGetErrorString(e)); exit(1); }
  }
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += (double)h[i];
  avg /= 148.0 * iters * 8;
  printf("%-36s N=%3d chains=%d  %7.1f cycles/MMA  (math floor %d)\n", name, n, ROT, avg, n / 2);
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  for (int n : {64, 128, 256}) {
    run<0, 1>("SW128 (+32B k-steps)", n, d);
    run<0, 2>("SW128 (+32B k-steps)", n, d);
    run<0, 4>("SW128 (+32B k-steps)", n, d);
    run<0, 8>("SW128 (+32B k-steps)", n, d);
    run<4, 1>("SW128 halo SBO, 128B-aligned start", n, d);
    run<4, 4>("SW128 halo SBO, 128B-aligned start", n, d);
    run<1, 1>("SW32 sub-tiles", n, d);
    run<1, 4>("SW32 sub-tiles", n, d);
    run<5, 1>("MN-major A and B (wgrad)", n, d);
    run<5, 4>("MN-major A and B (wgrad)", n, d);
  }
  return 0;
}
