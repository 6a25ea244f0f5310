// Micro-benchmark (GPU box only): cycles per tcgen05.mma kind::f16 (M=128, cta_group::1, SS operands, SWIZZLE_128B
// K-major) as a function of N in steps of 16 -- does a narrow or odd N (the C=36 head pads to 48) run at the
// operand-fetch floor (4096 + 32 N) / 128 cycles, or in slices?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I pose_estimation_amitai_b200/csrc -I include tools/ubench/mma_n_sweep.cu -o tools/ubench/mma_n_sweep.bin
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace pb::tc;

__global__ void __launch_bounds__(128, 1) k(int n, int rot, int iters, long long* out) {
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
    const uint32_t idesc = make_idesc(128, n, 0, 0);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 96 * 1024;
    uint64_t ad[8], bd[8];
    uint32_t acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = u & 3, tap = u >> 2;
      ad[u] = smem_desc_sw128(a0 + tap * 4096 + j * 32, 16, 1024);
      bd[u] = smem_desc_sw128(b0 + j * 32, 16, 1024);
      acc[u] = tm + (uint32_t)((u % rot) * n);
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

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 1000;
  for (int n = 16; n <= 256; n += 16) {
    for (int rot : {1, 4}) {
      if (rot * n > 512) continue;
      for (int rep = 0; rep < 2; ++rep) {
        k<<<148, 128, 200 * 1024>>>(n, rot, iters, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("n %d: %s\n", n, cudaGetErrorString(e)); return 1; }
      }
      long long h[148];
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      double avg = 0;
      for (int i = 0; i < 148; ++i) avg += (double)h[i];
      avg /= 148.0 * iters * 8;
      printf("N=%3d chains=%d  %7.1f cycles/MMA   math floor %3d   fetch floor %5.1f\n", n, rot, avg, n / 2,
             (4096.0 + 32.0 * n) / 128.0);
    }
  }
  return 0;
}
