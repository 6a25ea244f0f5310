// Weight (and bias) gradient of the first layer on tcgen05 straight from the NCHW fp32 crops:
//
//   dW[co][k = ci*9 + r*3 + s] = sum over pixels  col[pixel][k] * g[pixel][co],      dbias[co] = sum over pixels g[pixel][co]
//
// where col is the crop's im2col (pytorch/CNNs.py:24 conv1; autograd of it, pytorch/train_pytorch.py:137).  Round 1
// materialised col as a [pixels][64] 16-bit tensor -- 302 MB written and read again per batch-64 step for a 9 MB
// input -- only because this contraction needs it as a TMA-able operand.  Here the producer warps of csrc/tc_conv1.cu
// build the SAME shared-memory image (one 128-byte row per pixel, SWIZZLE_128B by hand) and the tensor core reads it
// as an MN-major A operand (K = pixel rows, M = the 64 k columns); the gradient tile [128 pixels][64 co] arrives by TMA
// as the MN-major B operand.  M is padded to 128 with a block of zeros the descriptor's LBO points at.
// The unused k column 36 of every pixel row holds 1.0, so accumulator row 36 IS the bias gradient.
//
// One persistent CTA per SM accumulates its tiles (4 image rows x 32 columns = 128 pixels) in ONE TMEM accumulator
// and writes one fp32 partial tile [64 k][Cg] (+ the bias row) in pb_wgrad_reduce's layout: ksplit = grid size.
//
// Warp roles (416 threads): warp 0 gradient-tile TMA + MMA issuer, warps 1-8 operand producers (two sets on alternate
// tiles), warps 9-12 epilogue (once, at the end).
#include <stdlib.h>
#include <string.h>

#include "tc_epilogue.cuh"

namespace pb {

using namespace tc;

constexpr int W1_THREADS = 416;
constexpr int W1_STAGES = 4;                  // even: the two producer sets own disjoint stages
constexpr int W1_TILE_ROWS = 4, W1_TILE_COLS = 32;
constexpr uint32_t W1_TILE_BYTES = 128 * 128; // [128 pixels][128 B]
constexpr uint32_t W1_STAGE_BYTES = 2 * W1_TILE_BYTES;   // col tile + gradient tile
constexpr int W1_ONES_K = 36;                 // first padding column of the k axis (Cin * 9 <= 36)

struct W1P {
  const float* in;
  float* partial;
  int N, C, H, W, dil, Cg, Ca;
  int groups_w, tiles_per_img, total_tiles;
  long long L;
};

__global__ void __launch_bounds__(W1_THREADS, 1)
tc_wgrad1_kernel(const __grid_constant__ CUtensorMap gmap, const W1P p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[W1_STAGES];
  __shared__ __align__(8) uint64_t g_full[W1_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[W1_STAGES];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_slot;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* zero_blk = smem + W1_STAGES * W1_STAGE_BYTES;     // 16 KB of zeros: rows 64..127 of the A operand
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (uint32_t i = threadIdx.x; i < W1_TILE_BYTES / 16; i += W1_THREADS)
    reinterpret_cast<uint4*>(zero_blk)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&gmap);
    for (int s = 0; s < W1_STAGES; ++s) {
      mbar_init(&a_full[s], 4);      // lane 0 of the four producer warps of the set that owns the stage
      mbar_init(&g_full[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 64);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the zero block is read by the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int ntiles = (p.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0) {
    if (elect_one()) {
      // ---------------------------------------------------------------- gradient tiles (TMA) + MMA issuer
      const uint32_t idesc = make_idesc(128, 64, 1, 1);
      // stage-0 descriptors; a later stage moves the start address up and the LBO (distance to the zero block) down by
      // the same amount -- both fields live in the low word, so it is one 32-bit add per stage
      const uint32_t sa0 = smem_u32(smem);
      // MN-major: 8 pixel rows = one 1024-byte swizzle group (SBO), the second 64-column block of M = the zero block
      const uint64_t ad = smem_desc_sw128(sa0, smem_u32(zero_blk) - sa0, 1024);
      const uint64_t gd = smem_desc_sw128(sa0 + W1_TILE_BYTES, 16, 1024);
      const uint32_t a_lo0 = (uint32_t)ad, a_hi0 = (uint32_t)(ad >> 32), g_lo0 = (uint32_t)gd, g_hi = (uint32_t)(gd >> 32);
      constexpr uint32_t S16 = W1_STAGE_BYTES >> 4;
      constexpr uint32_t A_STEP = S16 - (S16 << 16);
      auto load_g = [&](int it) {
        const int stage = it % W1_STAGES;
        const int tile = (int)blockIdx.x + it * (int)gridDim.x;
        const int img = tile / p.tiles_per_img;
        const int rem = tile - img * p.tiles_per_img;
        const int gh = rem / p.groups_w, gw = rem - gh * p.groups_w;
        mbar_wait(&empty_bar[stage], ((uint32_t)(it / W1_STAGES) & 1u) ^ 1u);
        mbar_expect_tx(&g_full[stage], W1_TILE_BYTES);
        tma_load_4d(smem + (size_t)stage * W1_STAGE_BYTES + W1_TILE_BYTES, &gmap, &g_full[stage], 0, gw * W1_TILE_COLS,
                    gh * W1_TILE_ROWS, img);
      };
      // the same thread feeds the gradient ring (W1_STAGES - 1 tiles ahead) and issues the MMAs
      for (int it = 0; it < ntiles && it < W1_STAGES - 1; ++it) load_g(it);
      for (int it = 0; it < ntiles; ++it) {
        const int stage = it % W1_STAGES;
        const uint32_t par = (uint32_t)(it / W1_STAGES) & 1u;
        mbar_wait(&g_full[stage], par);
        mbar_wait(&a_full[stage], par);
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < 8; ++j)     // 16 pixels = 2048 bytes per K step
          umma_bf16_lohi(tmem_base, a_lo0 + (uint32_t)stage * A_STEP + (uint32_t)j * 128u, a_hi0,
                         g_lo0 + (uint32_t)stage * S16 + (uint32_t)j * 128u, g_hi, idesc, (it > 0 || j > 0) ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (it + W1_STAGES - 1 < ntiles) load_g(it + W1_STAGES - 1);
      }
      umma_commit(&done_bar);
    }
  } else if (warp <= 8) {
    // ------------------------------------------------------------------ col-tile producers (as csrc/tc_conv1.cu)
    const int set = (warp - 1) >> 2;
    const int ml = ((warp - 1) & 3) * 32 + lane;
    const long long plane = (long long)p.H * p.W;
    for (int it = set; it < ntiles; it += 2) {
      const int tile = (int)blockIdx.x + it * (int)gridDim.x;
      const int img = tile / p.tiles_per_img;
      const int rem = tile - img * p.tiles_per_img;
      const int gh = rem / p.groups_w, gw = rem - gh * p.groups_w;
      const int y = gh * W1_TILE_ROWS + (ml >> 5), x = gw * W1_TILE_COLS + (ml & 31);
      float v[40];
#pragma unroll
      for (int k = 36; k < 40; ++k) v[k] = 0.f;
      v[W1_ONES_K] = 1.f;      // ones column: accumulator row 36 sums the gradient = dbias (out-of-image pixels have g = 0)
      {
        const bool inside = y < p.H && x < p.W;
        const float* src = p.in + (long long)img * p.C * plane + (inside ? (long long)y * p.W + x : 0ll);
        int off[9];
        uint32_t msk[9];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int dy = p.dil * (r - 1);
          const bool oky = inside && (unsigned)(y + dy) < (unsigned)p.H;
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            const int dx = p.dil * (s - 1);
            const bool okt = oky && (unsigned)(x + dx) < (unsigned)p.W;
            msk[r * 3 + s] = okt ? 0xFFFFFFFFu : 0u;
            off[r * 3 + s] = okt ? dy * p.W + dx : 0;
          }
        }
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const float* sc = src + (ci < p.C ? ci : 0) * plane;
          const uint32_t cm = ci < p.C ? 0xFFFFFFFFu : 0u;
#pragma unroll
          for (int t = 0; t < 9; ++t) v[ci * 9 + t] = __uint_as_float(__float_as_uint(__ldg(sc + off[t])) & msk[t] & cm);
        }
      }
      const int stage = it % W1_STAGES;
      mbar_wait(&empty_bar[stage], ((uint32_t)(it / W1_STAGES) & 1u) ^ 1u);
      uint8_t* row = smem + (size_t)stage * W1_STAGE_BYTES + ml * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint4 t = make_uint4(0u, 0u, 0u, 0u);
        if (j < 5) t = pack16x8<false>(v + 8 * j);
        *reinterpret_cast<uint4*>(row + ((j ^ (ml & 7)) << 4)) = t;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[stage]);
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps 9..12: the partial tile
    const int q = warp & 3;
    const int m = q * 32 + lane;        // accumulator row = k index (rows >= 64: the zero block)
    if (ntiles > 0) {
      mbar_wait(&done_bar, 0);
      tc_fence_after();
    }
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    float* base = p.partial + (long long)blockIdx.x * p.L;
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t r[32];
      if (ntiles > 0) {
        tmem_ld32(lane_base + (uint32_t)c0, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      float* dst = nullptr;
      if (m < p.C * 9) dst = base + (long long)m * p.Cg;              // weight rows: partial[(0 * Ca + k) * Cg + co]
      else if (m == W1_ONES_K) dst = base + (p.L - p.Cg);              // the bias row
      if (dst != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c0 + j < p.Cg) dst[c0 + j] = __uint_as_float(r[j]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

}  // namespace pb

using namespace pb;

extern "C" int pb_wgrad_first_tc(const pb_wgrad_first_args* a, void* stream) {
  PB_REQUIRE(a != nullptr && a->in && a->g && a->partial, "pb_wgrad_first_tc: null args");
  PB_REQUIRE_DEV(a->in, "in");
  PB_REQUIRE_DEV(a->g, "g");
  PB_REQUIRE_DEV(a->partial, "partial");
  PB_REQUIRE(a->N > 0 && a->H > 0 && a->W > 0 && a->dilation >= 1 && a->ksplit >= 1, "pb_wgrad_first_tc: bad geometry");
  if (a->C < 1 || a->C > 4 || a->ksize != 3 || a->Cg != 64 || a->Ca != 64 || a->act_dtype != PB_BF16) {
    set_error("pb_wgrad_first_tc: supports Cin <= 4, k = 3, 64 output channels, bf16 gradients, Ca (stored k width) 64");
    return PB_ERR_UNSUPPORTED;
  }
  W1P p;
  memset((void*)&p, 0, sizeof(p));
  p.in = a->in; p.partial = a->partial;
  p.N = a->N; p.C = a->C; p.H = a->H; p.W = a->W; p.dil = a->dilation; p.Cg = a->Cg; p.Ca = a->Ca;
  p.groups_w = cdiv(a->W, W1_TILE_COLS);
  p.tiles_per_img = cdiv(a->H, W1_TILE_ROWS) * p.groups_w;
  p.total_tiles = a->N * p.tiles_per_img;
  p.L = (long long)a->Ca * a->Cg + a->Cg;
  CUtensorMap gmap;
  {
    const uint64_t C = (uint64_t)a->Cg;
    const uint64_t dims[4] = {C, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->N};
    const uint64_t str[3] = {C * 2, (uint64_t)a->W * C * 2, (uint64_t)a->H * a->W * C * 2};
    const uint32_t box[4] = {64, W1_TILE_COLS, W1_TILE_ROWS, 1};
    int rc = encode_tmap_bf16(&gmap, a->g, 4, dims, str, box);
    if (rc != PB_OK) return rc;
  }
  // every split writes its partial tile (zero when it has no tiles): the grid IS the split count
  const int grid = a->ksplit;
  PB_REQUIRE(grid <= sm_count(), "pb_wgrad_first_tc: ksplit must not exceed the SM count (one persistent CTA per split)");
  const size_t smem = (size_t)W1_STAGES * W1_STAGE_BYTES + W1_TILE_BYTES + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_wgrad1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, "pb_wgrad_first_tc: smem attribute");
    attr_set = true;
  }
  tc_wgrad1_kernel<<<grid, W1_THREADS, smem, (cudaStream_t)stream>>>(gmap, p);
  PB_LAUNCH_CHECK("tc_wgrad1_kernel");
  return PB_OK;
}
