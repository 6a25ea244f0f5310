// tcgen05 batched GEMM for the ViT attention products (pytorch_vit_encoder.py:59-78 and their autograd):
//   C[z][m][n] = epilogue( alpha * sum_k A[z][m][k] * B[z][n][k] ),   z = (zb, zh) = (sample, head)
// S = 144 tokens, head dim 256: every product has M <= 256, N <= 256, K <= 256, so a CTA loads the WHOLE K extent
// of its [128 x N] tile in one TMA round (no ring), issues <= 16 MMAs and runs a row-per-thread epilogue:
//   mode 0  store alpha * acc                                     (P V, dV, dQ, dK)
//   mode 1  row softmax of alpha * acc  -> probabilities (bf16)    (Q K^T; Attention.forward :66-70)
//   mode 2  dS = P * (acc - sum_n acc * P) * alpha  (bf16)         (softmax backward on dP = dO V^T)
// Either operand may be K-major (k contiguous) or MN-major (m / n contiguous): all six products read q, k, v, dO
// and the probability matrices in place, no transposed copies.
// A CTA (256 threads) owns BOTH 128-row tiles of a 144-token problem: the B operand is loaded once, the second
// tile's MMAs (mostly padding rows) run behind the first tile's, and eight warps share the row-per-thread epilogue
// (warps 0-3 tile 0, warps 4-7 tile 1) -- the kernel is bound by its load -> MMA -> epilogue latency chain, one CTA per
// SM, so halving the CTA count nearly halves its time.
#include <string.h>

#include "tc_common.cuh"

namespace pb {

using namespace tc;

struct BgMaps {
  CUtensorMap a, b;
};

struct BgP {
  int M, N, K, ZH;
  int a_mn, b_mn;
  int n_mma, kchunks, nb;        // UMMA N (multiple of 16), 64-wide K chunks, 64-wide N blocks (MN-major B)
  uint32_t a_chunk_bytes, b_chunk_bytes;
  void* C;
  long long c_zb, c_zh, c_m;
  int c_bf16;
  float alpha;
  int mode;
  const __nv_bfloat16* P;        // mode 2: probabilities, indexed like C
};

__device__ __forceinline__ void tma_load_4d_(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                             int c3) {
  tma_load_4d(dst, m, bar, c0, c1, c2, c3);
}

constexpr int BG_THREADS = 256;

__global__ void __launch_bounds__(BG_THREADS, 1)
tc_bgemm_kernel(const __grid_constant__ BgMaps maps, const BgP p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t load_bar, mma_bar[2];
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 256;
  const int ntile = (p.M - m0 > 128) ? 2 : 1;   // 128-row tiles of this CTA that hold any valid row
  uint8_t* sa = smem;                            // [tile][kchunk] A chunks
  uint8_t* sb = smem + (size_t)(2 * p.kchunks) * p.a_chunk_bytes;
  const int z = blockIdx.y, zb = z / p.ZH, zh = z - zb * p.ZH;

  if (threadIdx.x == 0) {
    prefetch_tmap(&maps.a);
    prefetch_tmap(&maps.b);
    mbar_init(&load_bar, 1);
    mbar_init(&mma_bar[0], 1);
    mbar_init(&mma_bar[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0 && elect_one()) {
    const uint32_t idesc = make_idesc(128, p.n_mma, p.a_mn, p.b_mn);
    mbar_expect_tx(&load_bar, (uint32_t)p.kchunks * ((uint32_t)ntile * p.a_chunk_bytes + p.b_chunk_bytes));
    for (int kc = 0; kc < p.kchunks; ++kc) {
      uint8_t* db = sb + (size_t)kc * p.b_chunk_bytes;
      for (int tile = 0; tile < ntile; ++tile) {
        uint8_t* da = sa + (size_t)(tile * p.kchunks + kc) * p.a_chunk_bytes;
        const int mt = m0 + tile * 128;
        if (p.a_mn) {
          tma_load_4d_(da, &maps.a, &load_bar, mt, kc * 64, zh, zb);
          tma_load_4d_(da + 8192, &maps.a, &load_bar, mt + 64, kc * 64, zh, zb);
        } else {
          tma_load_4d_(da, &maps.a, &load_bar, kc * 64, mt, zh, zb);
        }
      }
      if (p.b_mn) {
        for (int j = 0; j < p.nb; ++j) tma_load_4d_(db + j * 8192, &maps.b, &load_bar, j * 64, kc * 64, zh, zb);
      } else {
        tma_load_4d_(db, &maps.b, &load_bar, kc * 64, 0, zh, zb);
      }
    }
    mbar_wait(&load_bar, 0);
    tc_fence_after();
    for (int tile = 0; tile < ntile; ++tile) {
      for (int kc = 0; kc < p.kchunks; ++kc) {
        const uint32_t a_addr = smem_u32(sa + (size_t)(tile * p.kchunks + kc) * p.a_chunk_bytes);
        const uint32_t b_addr = smem_u32(sb + (size_t)kc * p.b_chunk_bytes);
        const int ksteps = min(4, (p.K - kc * 64 + 15) / 16);
        for (int j = 0; j < ksteps; ++j) {
          const uint64_t ad = p.a_mn ? smem_desc_sw128(a_addr + j * 2048, 8192, 1024)
                                     : smem_desc_sw128(a_addr + j * 32, 16, 1024);
          const uint64_t bd = p.b_mn ? smem_desc_sw128(b_addr + j * 2048, 8192, 1024)
                                     : smem_desc_sw128(b_addr + j * 32, 16, 1024);
          umma_bf16(tmem_base + (uint32_t)(tile * p.n_mma), ad, bd, idesc, (kc > 0 || j > 0) ? 1u : 0u);
        }
      }
      umma_commit(&mma_bar[tile]);   // tile 0's epilogue starts while tile 1's MMAs run
    }
  }
  __syncwarp();
  // ---- epilogue: thread = output row; warps 0-3 tile 0, warps 4-7 tile 1 (TMEM lane quadrant = warp % 4)
  const int etile = warp >> 2;
  if (etile < ntile) {
  mbar_wait(&mma_bar[etile], 0);
  tc_fence_after();
  const int row = m0 + etile * 128 + (warp & 3) * 32 + lane;
  const bool row_ok = row < p.M;
  const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(etile * p.n_mma);
  const long long coff = (long long)zb * p.c_zb + (long long)zh * p.c_zh + (long long)row * p.c_m;
  const int nchunks = p.n_mma >> 4;
  float r_max = -INFINITY, r_sum = 0.f, r_dot = 0.f;
  if (p.mode == 1) {
    for (int c = 0; c < nchunks; ++c) {
      uint32_t r[16];
      tmem_ld16(lane_base + c * 16, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (c * 16 + j < p.N) r_max = fmaxf(r_max, p.alpha * __uint_as_float(r[j]));
    }
    for (int c = 0; c < nchunks; ++c) {
      uint32_t r[16];
      tmem_ld16(lane_base + c * 16, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (c * 16 + j < p.N) r_sum += expf(p.alpha * __uint_as_float(r[j]) - r_max);
    }
  } else if (p.mode == 2) {
    for (int c = 0; c < nchunks; ++c) {
      uint32_t r[16];
      tmem_ld16(lane_base + c * 16, r);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c * 16 + j < p.N) r_dot += __uint_as_float(r[j]) * __bfloat162float(p.P[coff + c * 16 + j]);
      }
    }
  }
  const float inv_sum = p.mode == 1 ? 1.f / r_sum : 0.f;
  for (int c = 0; c < nchunks; ++c) {
    uint32_t r[16];
    tmem_ld16(lane_base + c * 16, r);
    tmem_ld_wait();
    if (!row_ok) continue;
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float acc = __uint_as_float(r[j]);
      if (p.mode == 1) v[j] = expf(p.alpha * acc - r_max) * inv_sum;
      else if (p.mode == 2)
        v[j] = (c * 16 + j < p.N) ? __bfloat162float(p.P[coff + c * 16 + j]) * (acc - r_dot) * p.alpha : 0.f;
      else v[j] = p.alpha * acc;
    }
    const int n0 = c * 16;
    if (p.c_bf16) {
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.C) + coff + n0;
      if (n0 + 16 <= p.N && ((coff + n0) & 7) == 0) {
        uint4 t0, t1;
        t0.x = pack_bf16x2(v[0], v[1]); t0.y = pack_bf16x2(v[2], v[3]);
        t0.z = pack_bf16x2(v[4], v[5]); t0.w = pack_bf16x2(v[6], v[7]);
        t1.x = pack_bf16x2(v[8], v[9]); t1.y = pack_bf16x2(v[10], v[11]);
        t1.z = pack_bf16x2(v[12], v[13]); t1.w = pack_bf16x2(v[14], v[15]);
        reinterpret_cast<uint4*>(dst)[0] = t0;
        reinterpret_cast<uint4*>(dst)[1] = t1;
      } else {
        for (int j = 0; j < 16; ++j)
          if (n0 + j < p.N) dst[j] = __float2bfloat16_rn(v[j]);
      }
    } else {
      float* dst = reinterpret_cast<float*>(p.C) + coff + n0;
      for (int j = 0; j < 16; ++j)
        if (n0 + j < p.N) dst[j] = v[j];
    }
  }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// operand: element (z, i, k) at  ptr + zb*s_zb + zh*s_zh + i*s_i + k*s_k   with exactly one of s_i, s_k equal to 1
struct BgOperand {
  const void* ptr;
  long long s_zb, s_zh, s_i, s_k;
  int extent_i;   // M (for A) or N (for B)
};

static int encode_operand(CUtensorMap* map, const BgOperand& o, int K, int ZH, int ZB, int tile_i, bool* mn_major) {
  const bool mn = (o.s_i == 1);
  *mn_major = mn;
  if (!mn && o.s_k != 1) {
    set_error("tc_bgemm: operand has no unit stride");
    return PB_ERR_UNSUPPORTED;
  }
  const long long ld = mn ? o.s_k : o.s_i;
  if ((ld & 7) || (o.s_zh & 7) || (o.s_zb & 7) || (((uintptr_t)o.ptr) & 15)) {
    set_error("tc_bgemm: operand strides must be multiples of 8 elements (16 bytes)");
    return PB_ERR_UNSUPPORTED;
  }
  if (mn) {
    const uint64_t dims[4] = {(uint64_t)o.extent_i, (uint64_t)K, (uint64_t)ZH, (uint64_t)ZB};
    const uint64_t str[3] = {(uint64_t)ld * 2, (uint64_t)o.s_zh * 2, (uint64_t)o.s_zb * 2};
    const uint32_t box[4] = {64, 64, 1, 1};
    return encode_tmap_bf16(map, o.ptr, 4, dims, str, box);
  }
  const uint64_t dims[4] = {(uint64_t)K, (uint64_t)o.extent_i, (uint64_t)ZH, (uint64_t)ZB};
  const uint64_t str[3] = {(uint64_t)ld * 2, (uint64_t)o.s_zh * 2, (uint64_t)o.s_zb * 2};
  const uint32_t box[4] = {64, (uint32_t)tile_i, 1, 1};
  return encode_tmap_bf16(map, o.ptr, 4, dims, str, box);
}

// returns PB_ERR_UNSUPPORTED when the shape is outside this kernel (caller falls back to the CUDA-core GEMM)
int tc_bgemm(const BgOperand& A, const BgOperand& B, int M, int N, int K, int ZH, int ZB, void* C, long long c_zb,
             long long c_zh, long long c_m, bool c_bf16, float alpha, int mode, const void* P, cudaStream_t st) {
  if (N < 16 || N > 256 || K < 16 || K > 256 || M < 1 || (N & 7) || (K & 7)) return PB_ERR_UNSUPPORTED;
  BgP p;
  memset(&p, 0, sizeof(p));
  BgMaps maps;
  memset(&maps, 0, sizeof(maps));
  p.M = M; p.N = N; p.K = K; p.ZH = ZH;
  p.n_mma = cdiv(N, 16) * 16;
  p.kchunks = cdiv(K, 64);
  p.nb = cdiv(N, 64);
  bool amn, bmn;
  int rc = encode_operand(&maps.a, A, K, ZH, ZB, 128, &amn);
  if (rc != PB_OK) return rc;
  rc = encode_operand(&maps.b, B, K, ZH, ZB, p.n_mma, &bmn);
  if (rc != PB_OK) return rc;
  p.a_mn = amn ? 1 : 0;
  p.b_mn = bmn ? 1 : 0;
  p.a_chunk_bytes = 16384;
  p.b_chunk_bytes = bmn ? (uint32_t)p.nb * 8192u : (((uint32_t)p.n_mma * 128u + 1023u) & ~1023u);
  // TMA transfers whole boxes: the transaction count must match what the boxes deliver
  if (!bmn && p.b_chunk_bytes != (uint32_t)p.n_mma * 128u) return PB_ERR_UNSUPPORTED;
  p.C = C; p.c_zb = c_zb; p.c_zh = c_zh; p.c_m = c_m; p.c_bf16 = c_bf16 ? 1 : 0;
  p.alpha = alpha; p.mode = mode; p.P = (const __nv_bfloat16*)P;
  const size_t smem = (size_t)p.kchunks * (2 * p.a_chunk_bytes + p.b_chunk_bytes) + 1024;
  if (smem > 227 * 1024 - 2048) return PB_ERR_UNSUPPORTED;
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(tc_bgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "tc_bgemm: smem attribute");
    attr = smem;
  }
  dim3 grid(cdiv(M, 256), ZH * ZB);
  tc_bgemm_kernel<<<grid, BG_THREADS, smem, st>>>(maps, p);
  PB_LAUNCH_CHECK("tc_bgemm_kernel");
  return PB_OK;
}

}  // namespace pb
