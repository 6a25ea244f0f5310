// First layer on tcgen05 straight from the network input: Conv2d(Cin -> Cout, k3, dilation d, "same") + bias + LeakyReLU
// on the NCHW fp32 crop tensor (pytorch/CNNs.py:24,74; the crops are [B, 4, 192, 192], pytorch/preprocessor.py:33-39).
//
// Cin * 9 = 36 contraction elements are far below a 64-channel K chunk, so the other kernels' halo scheme does not
// apply; the training path materialises an im2col tensor ([pixels][64] 16-bit, 32 x the input's bytes) because the
// weight gradient needs it as a TMA-able operand anyway.  Inference does not: here eight producer warps gather each
// pixel's 36 taps from the fp32 image (L1 / L2 hits: every input element is used by nine pixels), convert and write
// the pixel's 128-byte K-major row -- SWIZZLE_128B by hand, zero padded to K = 64 -- into shared memory, and one MMA
// per 16 K elements (M = 128 pixels, N = Cout) consumes it.  HBM traffic per batch-64 launch: 9.4 MB read + 302 MB
// written, against 9.4 + 302 (im2col write) + 302 (read) + 302 MB for the two-kernel form.
//
// Warp roles (544 threads, persistent, one CTA per SM):
//   warp 0      barrier setup, resident weight tile (TMA), MMA issuer
//   warps 1-8   A-operand producers, two sets of four warps on alternate tiles (a tile = 4 image rows x 32 columns:
//               a warp's 32 lanes read 128 contiguous bytes per tap)
//   warps 9-16  epilogue, two per TMEM lane quadrant on alternate 32-channel chunks: bias, LeakyReLU, sign mask,
//               256-bit NHWC stores (a thread owns its pixel's 64-byte half row)
#include <stdlib.h>
#include <string.h>

#include "tc_epilogue.cuh"

namespace pb {

using namespace tc;

constexpr int C1_THREADS = 544;
constexpr int C1_STAGES = 6;              // even: the two producer sets own disjoint stages
constexpr int C1_TILE_ROWS = 4, C1_TILE_COLS = 32;
constexpr uint32_t C1_A_BYTES = 128 * 128;

struct C1P {
  const float* in;
  void* out;
  __nv_bfloat16* out2;
  uint32_t* mask_out;
  const float* bias;
  int N, C, H, W, dil, Cout;
  int groups_w, tiles_per_img, total_tiles;
  float slope;
  int f16;
};

template <bool F16, bool MASK>
__global__ void __launch_bounds__(C1_THREADS, 1)
tc_conv1_kernel(const __grid_constant__ CUtensorMap wmap, const C1P p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[C1_STAGES];
  __shared__ __align__(8) uint64_t a_empty[C1_STAGES];
  __shared__ __align__(8) uint64_t b_full;
  __shared__ __align__(8) uint64_t tmem_full[2];
  __shared__ __align__(8) uint64_t tmem_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float sbias[128];

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sb = smem + C1_STAGES * C1_A_BYTES;        // resident weights [Cout rows][128 B]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x < 128) sbias[threadIdx.x] = (p.bias != nullptr && (int)threadIdx.x < p.Cout) ? __ldg(p.bias + threadIdx.x) : 0.f;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&wmap);
    for (int s = 0; s < C1_STAGES; ++s) {
      mbar_init(&a_full[s], 4);      // lane 0 of the four producer warps of the set that owns the stage
      mbar_init(&a_empty[s], 1);
    }
    mbar_init(&b_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 8);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int ntiles = (p.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0) {
    if (elect_one()) {
      // ---------------------------------------------------------------- weights + MMA issuer
      mbar_expect_tx(&b_full, (uint32_t)p.Cout * 128u);
      tma_load_3d(sb, &wmap, &b_full, 0, 0, 0);
      const uint32_t idesc = make_idesc(128, p.Cout, 0, 0, F16, F16);
      const uint64_t ad = smem_desc_sw128(smem_u32(smem), 16, 1024);
      const uint64_t bd = smem_desc_sw128(smem_u32(sb), 16, 1024);
      const uint32_t a_lo0 = (uint32_t)ad, a_hi = (uint32_t)(ad >> 32), b_lo = (uint32_t)bd, b_hi = (uint32_t)(bd >> 32);
      mbar_wait(&b_full, 0);
      tc_fence_after();
      for (int it = 0; it < ntiles; ++it) {
        const int as = it & 1, stage = it % C1_STAGES;
        mbar_wait(&tmem_empty[as], (((uint32_t)it >> 1) & 1u) ^ 1u);
        mbar_wait(&a_full[stage], (uint32_t)(it / C1_STAGES) & 1u);
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(as * p.Cout);
        const uint32_t a_lo = a_lo0 + (((uint32_t)stage * C1_A_BYTES) >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_lohi(d, a_lo + 2u * k, a_hi, b_lo + 2u * k, b_hi, idesc, k ? 1u : 0u);
        umma_commit(&a_empty[stage]);
        umma_commit(&tmem_full[as]);
      }
    }
  } else if (warp <= 8) {
    // ------------------------------------------------------------------ A-operand producers
    const int set = (warp - 1) >> 2;                 // tiles it = set, set + 2, ...
    const int ml = ((warp - 1) & 3) * 32 + lane;     // operand row = pixel of the tile: image row ml >> 5, column ml & 31
    const long long plane = (long long)p.H * p.W;
    for (int it = set; it < ntiles; it += 2) {
      const int tile = (int)blockIdx.x + it * (int)gridDim.x;
      const int img = tile / p.tiles_per_img;
      const int rem = tile - img * p.tiles_per_img;
      const int gh = rem / p.groups_w, gw = rem - gh * p.groups_w;
      const int y = gh * C1_TILE_ROWS + (ml >> 5), x = gw * C1_TILE_COLS + (ml & 31);
      float v[40];
#pragma unroll
      for (int k = 36; k < 40; ++k) v[k] = 0.f;
      {
        // Unconditional loads: a tap outside the image (zero padding) reads the pixel itself instead and is masked to
        // zero afterwards, an out-of-image pixel of a partial tile reads pixel (0, 0) -- no predicated loads, no
        // per-tap branches; offsets and masks are computed once per pixel and shared by the input channels.
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
      const int stage = it % C1_STAGES;
      mbar_wait(&a_empty[stage], ((uint32_t)(it / C1_STAGES) & 1u) ^ 1u);
      uint8_t* row = smem + (size_t)stage * C1_A_BYTES + ml * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint4 t = make_uint4(0u, 0u, 0u, 0u);
        if (j < 5) t = pack16x8<F16>(v + 8 * j);
        *reinterpret_cast<uint4*>(row + ((j ^ (ml & 7)) << 4)) = t;     // SWIZZLE_128B: 16-byte chunk index ^ (row & 7)
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[stage]);
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps 9..16
    const int ew = warp - 9;
    const int q = warp & 3;
    const int half = ew >> 2;
    const int ml = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const int words = p.Cout >> 5;
    const float slope = p.slope;
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
    for (int it = 0; it < ntiles; ++it) {
      const int tile = (int)blockIdx.x + it * (int)gridDim.x;
      const int img = tile / p.tiles_per_img;
      const int rem = tile - img * p.tiles_per_img;
      const int gh = rem / p.groups_w, gw = rem - gh * p.groups_w;
      const int y = gh * C1_TILE_ROWS + (ml >> 5), x = gw * C1_TILE_COLS + (ml & 31);
      const bool ok = y < p.H && x < p.W;
      const long long pix = ((long long)img * p.H + y) * p.W + x;
      const int as = it & 1;
      mbar_wait(&tmem_full[as], ((uint32_t)it >> 1) & 1u);
      tc_fence_after();
      for (int cc = half; cc < words; cc += 2) {
        uint32_t rr[32];
        tmem_ld32(lane_base + (uint32_t)(as * p.Cout + cc * 32), rr);
        tmem_ld_wait();
        float v[32];
        uint32_t bits = 0;
#pragma unroll
        for (int j4 = 7; j4 >= 0; --j4) {
          const float4 b4 = *reinterpret_cast<const float4*>(sbias + cc * 32 + 4 * j4);
          const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int e = 3; e >= 0; --e) {
            const int j = 4 * j4 + e;
            const float t = __uint_as_float(rr[j]) + bb[e];
            if (MASK) bits = __funnelshift_l((uint32_t)(-(int)__float_as_uint(t)), bits, 1);    // sign of -t = (t > 0)
            v[j] = fmaxf(t, slope * t);
          }
        }
        if (ok) {
          if (MASK) p.mask_out[pix * words + cc] = bits;
          __nv_bfloat16* dst = out + pix * p.Cout + cc * 32;
          st_global_256(dst, pack16x8<F16>(v), pack16x8<F16>(v + 8));
          st_global_256(dst + 16, pack16x8<F16>(v + 16), pack16x8<F16>(v + 24));
          if (F16 && p.out2 != nullptr) {     // training in the "fp16" precision: the bf16 twin the next weight gradient reads
            __nv_bfloat16* d2 = p.out2 + pix * p.Cout + cc * 32;
            st_global_256(d2, pack16x8<false>(v), pack16x8<false>(v + 8));
            st_global_256(d2 + 16, pack16x8<false>(v + 16), pack16x8<false>(v + 24));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
    }
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

extern "C" int pb_conv_first_tc(const pb_conv_first_args* a, void* stream) {
  PB_REQUIRE(a != nullptr && a->in && a->w && a->out, "pb_conv_first_tc: null args");
  PB_REQUIRE_DEV(a->in, "in");
  PB_REQUIRE_DEV(a->w, "w");
  PB_REQUIRE_DEV(a->out, "out");
  PB_REQUIRE_DEV(a->bias, "bias");
  PB_REQUIRE_DEV(a->mask_out, "mask_out");
  PB_REQUIRE_DEV(a->out2, "out2");
  PB_REQUIRE(a->N >= 0 && a->H > 0 && a->W > 0 && a->dilation >= 1, "pb_conv_first_tc: bad geometry");
  PB_REQUIRE(a->act_dtype == PB_BF16 || a->act_dtype == PB_F16, "pb_conv_first_tc: bf16 / fp16 activations only");
  if (a->C < 1 || a->C > 4 || a->ksize != 3 || (a->Cout != 32 && a->Cout != 64 && a->Cout != 128) ||
      !(a->slope > 0.f && a->slope <= 1.f)) {
    set_error("pb_conv_first_tc: supports Cin <= 4, k = 3, Cout in {32, 64, 128}, LeakyReLU slope in (0, 1]");
    return PB_ERR_UNSUPPORTED;
  }
  if (a->N == 0) return PB_OK;
  C1P p;
  memset((void*)&p, 0, sizeof(p));
  p.in = a->in; p.out = a->out; p.mask_out = a->mask_out; p.bias = a->bias;
  p.out2 = a->act_dtype == PB_F16 ? reinterpret_cast<__nv_bfloat16*>(a->out2) : nullptr;
  p.N = a->N; p.C = a->C; p.H = a->H; p.W = a->W; p.dil = a->dilation; p.Cout = a->Cout;
  p.groups_w = cdiv(a->W, C1_TILE_COLS);
  p.tiles_per_img = cdiv(a->H, C1_TILE_ROWS) * p.groups_w;
  p.total_tiles = a->N * p.tiles_per_img;
  p.slope = a->slope;
  CUtensorMap wmap;
  {
    const uint64_t dims[3] = {64, (uint64_t)a->Cout, 1};
    const uint64_t str[2] = {128, (uint64_t)a->Cout * 128};
    const uint32_t box[3] = {64, (uint32_t)a->Cout, 1};
    int rc = encode_tmap_bf16(&wmap, a->w, 3, dims, str, box);
    if (rc != PB_OK) return rc;
  }
  const size_t smem = (size_t)C1_STAGES * C1_A_BYTES + (size_t)a->Cout * 128 + 1024;
  typedef void (*Kern)(const CUtensorMap, const C1P);
  const bool f16 = a->act_dtype == PB_F16, mk = a->mask_out != nullptr;
  const Kern kern = f16 ? (mk ? tc_conv1_kernel<true, true> : tc_conv1_kernel<true, false>)
                        : (mk ? tc_conv1_kernel<false, true> : tc_conv1_kernel<false, false>);
  static bool attr_set[4] = {false, false, false, false};
  const int ki = (f16 ? 2 : 0) + (mk ? 1 : 0);
  if (!attr_set[ki]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, "pb_conv_first_tc: smem attribute");
    attr_set[ki] = true;
  }
  const int grid = p.total_tiles < sm_count() ? p.total_tiles : sm_count();
  kern<<<grid, C1_THREADS, smem, (cudaStream_t)stream>>>(wmap, p);
  PB_LAUNCH_CHECK("tc_conv1_kernel");
  return PB_OK;
}
