// Network head on tcgen05: the last layer of Decoder2d -- ConvTranspose2d(k3, s2, p1, op1) + LeakyReLU
// (pytorch/CNNs.py:125-128,155) -- with its FOUR OUTPUT PARITIES FOLDED INTO THE MMA's N.
//
// A stride-2 transposed 3x3 convolution is four small convolutions over the INPUT grid, one per output parity
// (py, px): out[2y+py, 2x+px] = sum over the taps of that parity of W_t * in[y+sy, x+sx], with shifts (sy, sx) in
// {0,1}^2.  The generic halo kernel (tc_conv2.cu) gives every parity its own accumulator and issues one N = Cout MMA
// per tap -- 36 narrow (N = 48) MMAs per 64-channel K chunk, each paying the N-independent 4 KB fetch of its A
// operand, and a single-tile issue loop that runs at twice that floor.  Here the nine taps are grouped by SHIFT:
//
//     accumulator columns   [ P00 | P01 | P11 | P10 ]                (NT = Cout rounded up to 16, each)
//     shift (0,0)  N = 4 NT  taps (1,1) (1,2) (2,2) (2,1)  -> columns [0, 4NT)
//     shift (0,1)  N = 2 NT  taps (1,0) (2,0)              -> columns [NT, 3NT)
//     shift (1,0)  N = 2 NT  taps (0,2) (0,1)              -> columns [2NT, 4NT)
//     shift (1,1)  N = 1 NT  tap  (0,0)                    -> columns [2NT, 3NT)
//
// (the cyclic parity order makes every group's columns contiguous), so one K step is FOUR MMAs -- 16 instead of 36 per
// K chunk -- whose B operands are simply runs of consecutive tap tiles of the packed weights, loaded once per CTA in
// this order and kept resident.  The A operand of every shift is the same shared-memory halo box
// [(16+1) x (8+1) pixels x 64 channels] seen through a shifted UMMA descriptor (as tc_conv2.cu).
//
// Epilogue (8 warps; a TMEM lane quadrant is shared by two warps, one per output-row parity): a thread owns one input
// pixel = the two horizontally adjacent output pixels of its row parity, all channels.
//   MODE 0  bias + LeakyReLU, NCHW fp32 heatmaps                      (pb_conv_tc, out_nchw_f32)
//   MODE 1  per-map arg-max, heatmaps never written                   (pb_convT_argmax_fused)
//           -- every thread keeps a running (key, index) best per channel in registers; they are folded across the
//              warp (redux.sync) and the CTA (shared-memory atomicMax) and reach HBM (atomicMax on the order key) only
//              when the CTA's contiguous tile range moves on to the next image
//   MODE 2  MSE loss against Gaussian targets rendered on the fly + bf16 NHWC gradient w.r.t. the pre-activation
//   MODE 3  the same against a target tensor                           (both: pb_convT_mse_fused)
// Warp 0: TMA producer (resident weights once, then halo boxes), warp 1: MMA issuer, two accumulator stages in TMEM.
#include <stdlib.h>
#include <string.h>

#include "tc_epilogue.cuh"
#include "tc_head.cuh"

namespace pb {

using namespace tc;

constexpr int HD_THREADS = 320;
constexpr int HD_MAX_A_STAGES = 6;
constexpr int HD_TILE_H = 16, HD_TILE_W = 8;
constexpr int HD_BOX_ROWS = HD_TILE_H + 1, HD_BOX_COLS = HD_TILE_W + 1;
constexpr uint32_t HD_A_TX = HD_BOX_ROWS * HD_BOX_COLS * 128u;   // bytes one halo box brings
constexpr uint32_t HD_A_STAGE = (HD_A_TX + 1023u) & ~1023u;

struct HdMaps {
  CUtensorMap a;   // input [N, IH, IW, Cin], box [64 ch][9][17]
  CUtensorMap b;   // packed weights [9 taps][NT rows][Cin], box [64 ch][NT][1]
};

struct HdP {
  int N, IH, IW, OH, OW, Cout, kchunks, a_stages;
  int groups_w, tiles_per_img, total_tiles, tiles_per_cta;
  uint32_t b_off;     // resident weights behind the halo ring
  int f16;            // operands are IEEE half instead of bf16
  float slope;
  const float* bias;
  float* out;
  unsigned long long* keys;
  const float* target;
  const float* points;
  float negk2, gscale;
  float* loss;
  __nv_bfloat16* grad;
  float* dbias;
};

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// order_key (common.cuh) without its early return: NaN greatest, -0 ties with +0
__device__ __forceinline__ uint32_t order_key_bf(float v) {
  const uint32_t u = __float_as_uint(v + 0.0f);
  const uint32_t k = u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u);
  return (v != v) ? 0xFFFFFFFFu : k;
}

template <int NT, int MODE>
__global__ void __launch_bounds__(HD_THREADS, 1)
tc_head_kernel(const __grid_constant__ HdMaps maps, const HdP p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[HD_MAX_A_STAGES];
  __shared__ __align__(8) uint64_t a_empty[HD_MAX_A_STAGES];
  __shared__ __align__(8) uint64_t b_full;
  __shared__ __align__(8) uint64_t tmem_full[2];
  __shared__ __align__(8) uint64_t tmem_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float sbias[64];
  __shared__ __align__(8) unsigned long long s_best[64];
  __shared__ __align__(8) float2 s_pts[64];
  __shared__ float s_dbias[8][64];   // per epilogue warp: summed in a fixed order (deterministic)

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x < 64) {
    sbias[threadIdx.x] = (p.bias != nullptr && (int)threadIdx.x < p.Cout) ? __ldg(p.bias + threadIdx.x) : 0.f;
    s_best[threadIdx.x] = 0ull;
    s_pts[threadIdx.x] = make_float2(0.f, 0.f);
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.a);
    prefetch_tmap(&maps.b);
    for (int s = 0; s < HD_MAX_A_STAGES; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    mbar_init(&b_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 8);   // lane 0 of the eight epilogue warps
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  // this CTA's contiguous range of tiles: a tile is 16 x 8 INPUT pixels (= 32 x 16 output pixels) of one image
  const int tile0 = (int)blockIdx.x * p.tiles_per_cta;
  const int ntiles = min(p.tiles_per_cta, p.total_tiles - tile0);

  if (warp == 0) {
    if (elect_one()) {
      // ---------------------------------------------------------------- TMA producer
      uint8_t* sb = smem + p.b_off;
      // resident weights, tap tiles in group order (see the header): positions 0-3 shift (0,0), 4-5 shift (0,1),
      // 6-7 shift (1,0), 8 shift (1,1); tap index = r*3 + s of the 3x3 kernel
      const int perm[9] = {4, 5, 8, 7, 3, 6, 2, 1, 0};
      mbar_expect_tx(&b_full, (uint32_t)(9 * p.kchunks) * (uint32_t)(NT * 128));
      for (int kc = 0; kc < p.kchunks; ++kc)
#pragma unroll
        for (int pos = 0; pos < 9; ++pos)
          tma_load_3d(sb + (size_t)(kc * 9 + pos) * (NT * 128), &maps.b, &b_full, kc * 64, 0, perm[pos]);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < ntiles; ++it) {
        const int tile = tile0 + it;
        const int img = tile / p.tiles_per_img;
        const int rem = tile - img * p.tiles_per_img;
        const int gh = rem / p.groups_w, gw = rem - gh * p.groups_w;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&a_empty[stage], phase ^ 1);
          mbar_expect_tx(&a_full[stage], HD_A_TX);
          tma_load_4d(smem + (size_t)stage * HD_A_STAGE, &maps.a, &a_full[stage], kc * 64, gw * HD_TILE_W, gh * HD_TILE_H, img);
          if (++stage == p.a_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ---------------------------------------------------------------- MMA issuer
      const uint32_t id4 = make_idesc(128, 4 * NT, 0, 0, p.f16, p.f16);
      const uint32_t id2 = make_idesc(128, 2 * NT, 0, 0, p.f16, p.f16);
      const uint32_t id1 = make_idesc(128, NT, 0, 0, p.f16, p.f16);
      // A: tile row = 8 consecutive 128-byte pixel rows of the box, tile rows one box row (9 pixels) apart; the shift
      // only moves the start (128-byte aligned starts: the swizzle is a function of the absolute address)
      const uint32_t sa0 = smem_u32(smem);
      const uint64_t ad = smem_desc_sw128(sa0, 16, HD_BOX_COLS * 128);
      const uint32_t a_hi = (uint32_t)(ad >> 32);
      const uint32_t a00 = (uint32_t)ad, a01 = a00 + (128u >> 4), a10 = a00 + ((HD_BOX_COLS * 128u) >> 4),
                     a11 = a10 + (128u >> 4);
      const uint64_t bd = smem_desc_sw128(smem_u32(smem + p.b_off), 16, 1024);
      const uint32_t b_hi = (uint32_t)(bd >> 32);
      const uint32_t b0 = (uint32_t)bd;
      constexpr uint32_t TAP16 = (NT * 128) >> 4;         // one tap tile, in descriptor units
      mbar_wait(&b_full, 0);
      tc_fence_after();
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < ntiles; ++it) {
        const int as = it & 1;
        mbar_wait(&tmem_empty[as], (((uint32_t)it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(as * 4 * NT);
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&a_full[stage], phase);
          tc_fence_after();
          const uint32_t so = ((uint32_t)stage * HD_A_STAGE) >> 4;
          const uint32_t bk = b0 + (uint32_t)kc * 9u * TAP16;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t ko = 2u * (uint32_t)k;   // 16 K elements = 32 bytes
            umma_bf16_lohi(d, a00 + so + ko, a_hi, bk + ko, b_hi, id4, (kc | k) ? 1u : 0u);
            umma_bf16_lohi(d + NT, a01 + so + ko, a_hi, bk + 4 * TAP16 + ko, b_hi, id2, 1u);
            umma_bf16_lohi(d + 2 * NT, a10 + so + ko, a_hi, bk + 6 * TAP16 + ko, b_hi, id2, 1u);
            umma_bf16_lohi(d + 2 * NT, a11 + so + ko, a_hi, bk + 8 * TAP16 + ko, b_hi, id1, 1u);
          }
          umma_commit(&a_empty[stage]);
          if (++stage == p.a_stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[as]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps 2..9
    const int ew = warp - 2;
    const int q = warp & 3;            // TMEM lane quadrant this warp may read
    const int py = ew >> 2;            // output-row parity of this warp
    const int et = ew * 32 + lane;     // 0..255
    const int ml = q * 32 + lane;      // pixel of the tile: row ml >> 3, column ml & 7
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t col_px0 = py ? 3u * NT : 0u, col_px1 = py ? 2u * NT : (uint32_t)NT;
    const long long plane = (long long)p.OH * p.OW;
    const float slope = p.slope;
    constexpr int NC = NT / 16;
    float loss_acc = 0.f;
    // MODE 1: this thread's running best (order key, flat index) of every channel over the tiles of the current image.
    // A thread meets its pixels in increasing flat-index order (tiles row-major, px = 0 before 1), so a strict '>'
    // keeps the lowest index among equal maxima (Augmentor.py:131: first occurrence) with no per-tile exchange.
    // MODE 2 / 3: this thread's share of the bias gradient (sum of the gradient over its pixels), per channel
    float bsum[MODE >= 2 ? NT : 1];
    if (MODE >= 2) {
#pragma unroll
      for (int c = 0; c < NT; ++c) bsum[c] = 0.f;
    }
    uint32_t bk[MODE == 1 ? NT : 1], bi[MODE == 1 ? NT : 1];
    if (MODE == 1) {
#pragma unroll
      for (int c = 0; c < NT; ++c) { bk[c] = 0u; bi[c] = 0u; }
    }
    // warp-level fold of the running bests into the CTA's per-channel state (once per image, not per tile)
    auto fold_bests = [&]() {
#pragma unroll
      for (int c = 0; c < (MODE == 1 ? NT : 0); ++c) {
        const uint32_t m = __reduce_max_sync(0xffffffffu, bk[c]);
        if (bk[c] == m && m != 0u)
          atomicMax(&s_best[c], ((unsigned long long)m << 32) | (unsigned long long)(0xFFFFFFFFu - bi[c]));
        bk[c] = 0u;
      }
    };
    int prev_img = -1;
    for (int it = 0; it < ntiles; ++it) {
      const int tile = tile0 + it;
      const int img = tile / p.tiles_per_img;
      const int rem = tile - img * p.tiles_per_img;
      const int gh = rem / p.groups_w, gw = rem - gh * p.groups_w;
      if (MODE != 0 && MODE != 3 && img != prev_img) {
        // every epilogue warp walks the same tile sequence, so all eight meet here: the image's state changes hands
        if (MODE == 1 && prev_img >= 0) fold_bests();
        epi_bar_sync();
        if (et < NT) {
          if (MODE == 1) {
            const unsigned long long best = s_best[et];
            if (prev_img >= 0 && et < p.Cout && best != 0ull) atomicMax(p.keys + (long long)prev_img * p.Cout + et, best);
            s_best[et] = 0ull;
          } else if (et < p.Cout) {
            s_pts[et] = __ldg(reinterpret_cast<const float2*>(p.points) + (long long)img * p.Cout + et);
          }
        }
        epi_bar_sync();
        prev_img = img;
      }
      const int as = it & 1;
      const int by = gh * HD_TILE_H + (ml >> 3), bx = gw * HD_TILE_W + (ml & 7);
      const bool ok = by < p.IH && bx < p.IW;
      const int oy = 2 * by + py, ox0 = 2 * bx;
      const uint32_t tcol = lane_base + (uint32_t)(as * 4 * NT);
      const uint32_t idx0 = (uint32_t)(oy * p.OW + ox0);
      const long long opix = (long long)oy * p.OW + ox0;
      mbar_wait(&tmem_full[as], ((uint32_t)it >> 1) & 1u);
      tc_fence_after();
      // accumulator chunks (16 channels of both output pixels) double-buffered in registers: chunk cc + 1 is in
      // flight while chunk cc is worked on
      uint32_t r0[2][16], r1[2][16];
      tmem_ld16(tcol + col_px0, r0[0]);
      tmem_ld16(tcol + col_px1, r1[0]);
#pragma unroll
      for (int cc = 0; cc < NC; ++cc) {
        tmem_ld_wait();
        if (cc + 1 < NC) {
          tmem_ld16(tcol + col_px0 + (uint32_t)((cc + 1) * 16), r0[(cc + 1) & 1]);
          tmem_ld16(tcol + col_px1 + (uint32_t)((cc + 1) * 16), r1[(cc + 1) & 1]);
        } else {
          // the accumulator stage is in registers: hand it back before the last chunk's arithmetic
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[as]);
        }
        const uint32_t(&a0)[16] = r0[cc & 1];
        const uint32_t(&a1)[16] = r1[cc & 1];
        // every chunk but the last is full (NT = Cout rounded up to 16): only the last one tests the channel count
        const int nvalid = p.Cout - cc * 16;
#define HD_VALID(j) (cc + 1 < NC || (j) < nvalid)
        if (MODE == 0) {
          if (ok) {
            float* dst = p.out + ((long long)img * p.Cout + cc * 16) * plane + opix;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float b = sbias[cc * 16 + j];
              float v0 = __uint_as_float(a0[j]) + b, v1 = __uint_as_float(a1[j]) + b;
              v0 = fmaxf(v0, slope * v0);      // LeakyReLU for 0 < slope <= 1 (checked on the host)
              v1 = fmaxf(v1, slope * v1);
              if (HD_VALID(j)) __stcs(reinterpret_cast<float2*>(dst + j * plane), make_float2(v0, v1));
            }
          }
        } else if (MODE == 1) {
          // branch-free: pad channels (zero weights, zero bias) and out-of-image pixels run too; the former are never
          // read back, the latter carry key 0, which no value maps to
          const uint32_t okm = ok ? 0xFFFFFFFFu : 0u;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float b = sbias[cc * 16 + j];
            float v0 = __uint_as_float(a0[j]) + b, v1 = __uint_as_float(a1[j]) + b;
            v0 = fmaxf(v0, slope * v0);
            v1 = fmaxf(v1, slope * v1);
            const uint32_t k0 = order_key_bf(v0), k1 = order_key_bf(v1);
            const bool hi = k1 > k0;                                   // ties keep the lower index
            const uint32_t kb = (hi ? k1 : k0) & okm;
            const uint32_t ib = hi ? idx0 + 1u : idx0;
            const bool up = kb > bk[cc * 16 + j];
            bk[cc * 16 + j] = up ? kb : bk[cc * 16 + j];
            bi[cc * 16 + j] = up ? ib : bi[cc * 16 + j];
          }
        } else {
          float g0[16], g1[16];
          const float* tgt = MODE == 3 ? p.target + ((long long)img * p.Cout + cc * 16) * plane + opix : nullptr;
          const float gs = p.gscale, gsl = p.gscale * slope;
          const float fx0 = (float)ox0, fx1 = (float)(ox0 + 1), fy = (float)oy;
          float part = 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float b = sbias[cc * 16 + j];
            float v0 = __uint_as_float(a0[j]) + b, v1 = __uint_as_float(a1[j]) + b;
            const float gs0 = v0 > 0.f ? gs : gsl, gs1 = v1 > 0.f ? gs : gsl;
            v0 = fmaxf(v0, slope * v0);
            v1 = fmaxf(v1, slope * v1);
            float t0, t1;
            if (MODE == 3) {
              float2 tt = make_float2(0.f, 0.f);
              if (ok && HD_VALID(j)) tt = __ldcs(reinterpret_cast<const float2*>(tgt + j * plane));
              t0 = tt.x; t1 = tt.y;
            } else {
              // Gaussian target rendered here: exp(-r^2 / 2 sigma^2) as one ex2.approx, exactly as mse_nhwc_bf16_kernel
              const float2 mxy = s_pts[cc * 16 + j];
              const float dx0 = fx0 - mxy.x, dx1 = fx1 - mxy.x, dy = fy - mxy.y;
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"((dx0 * dx0 + dy * dy) * p.negk2));
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"((dx1 * dx1 + dy * dy) * p.negk2));
            }
            float d0 = v0 - t0, d1 = v1 - t1;
            if (!HD_VALID(j)) d0 = d1 = 0.f;     // channel padding: no loss, zero gradient
            part += d0 * d0 + d1 * d1;
            g0[j] = d0 * gs0;
            g1[j] = d1 * gs1;
          }
          if (ok) {
            loss_acc += part;
            if (p.dbias != nullptr) {
#pragma unroll
              for (int j = 0; j < 16; ++j) bsum[cc * 16 + j] += g0[j] + g1[j];
            }
            // gradient rows of the two output pixels are adjacent in NHWC: channels [cc*16, cc*16+16) of each
            __nv_bfloat16* gd = p.grad + ((long long)img * plane + opix) * NT + cc * 16;
            st_global_256(gd, pack16x8<false>(g0), pack16x8<false>(g0 + 8));
            st_global_256(gd + NT, pack16x8<false>(g1), pack16x8<false>(g1 + 8));
          }
        }
#undef HD_VALID
      }
    }
    if (MODE == 1) {
      if (prev_img >= 0) fold_bests();
      epi_bar_sync();
      if (et < p.Cout && prev_img >= 0) {
        const unsigned long long best = s_best[et];
        if (best != 0ull) atomicMax(p.keys + (long long)prev_img * p.Cout + et, best);
      }
    }
    if (MODE >= 2) {
      loss_acc = warp_sum(loss_acc);
      if (lane == 0 && loss_acc != 0.f) atomicAdd(p.loss, loss_acc);
      if (p.dbias != nullptr) {
        // bias gradient: warp fold (shuffles), CTA fold over the eight warps in a fixed order, one row per CTA
#pragma unroll
        for (int c = 0; c < (MODE >= 2 ? NT : 0); ++c) {
          const float sc = warp_sum(bsum[c]);
          if (lane == 0) s_dbias[ew][c] = sc;
        }
        epi_bar_sync();
        if (et < p.Cout) {
          float t = 0.f;
#pragma unroll
          for (int w8 = 0; w8 < 8; ++w8) t += s_dbias[w8][et];
          p.dbias[(long long)blockIdx.x * p.Cout + et] += t;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static int hd_env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v != nullptr && *v) ? atoi(v) : dflt;
}

template <int NT>
static int head_launch(int mode, const HdMaps& maps, const HdP& p, int grid, size_t smem, cudaStream_t stream) {
  typedef void (*Kern)(const HdMaps, const HdP);
  const Kern kern = mode == 1 ? tc_head_kernel<NT, 1> : mode == 2 ? tc_head_kernel<NT, 2>
                    : mode == 3 ? tc_head_kernel<NT, 3> : tc_head_kernel<NT, 0>;
  static bool attr_set[4] = {false, false, false, false};
  if (!attr_set[mode]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, "tc_head_kernel: smem attribute");
    attr_set[mode] = true;
  }
  kern<<<grid, HD_THREADS, smem, stream>>>(maps, p);
  PB_LAUNCH_CHECK("tc_head_kernel");
  return PB_OK;
}

int head_tc(const pb_conv_args* a, const V2Head* head, cudaStream_t stream) {
  if (hd_env_int("POSEB200_HEAD_V2", 1) == 0) return PB_ERR_UNSUPPORTED;
  const pb_taps& tp = a->taps;
  if (!(tp.out_mul == 1 && tp.in_div == 2) || tp.ntaps != 9 || !a->out_nchw_f32) return PB_ERR_UNSUPPORTED;
  for (int t = 0; t < 9; ++t)
    if (tp.dy[t] != 1 - t / 3 || tp.dx[t] != 1 - t % 3) return PB_ERR_UNSUPPORTED;   // ConvTranspose2d(k3, s2, p1, op1)
  if (a->OH != 2 * a->IH || a->OW != 2 * a->IW || (a->Cin & 63) != 0) return PB_ERR_UNSUPPORTED;
  if (a->add0 || a->add1 || a->pre_out || a->mask_out || (a->act != PB_ACT_NONE && a->act != PB_ACT_LRELU)) return PB_ERR_UNSUPPORTED;
  if (a->act_dtype != PB_BF16 && a->act_dtype != PB_F16) return PB_ERR_UNSUPPORTED;
  const int nt = cdiv(a->Cout, 16) * 16;
  if (nt > 48) return PB_ERR_UNSUPPORTED;   // the arg-max form keeps 2 x NT running bests in registers
  int mode = head != nullptr ? head->mode : 0;
  if (mode == 2 && head->cpad != nt) return PB_ERR_UNSUPPORTED;
  if (mode == 2 && head->target != nullptr) mode = 3;   // MSE against a target tensor (2: Gaussian targets from keypoints)
  HdP p;
  HdMaps maps;
  memset((void*)&p, 0, sizeof(p));
  memset(&maps, 0, sizeof(maps));
  p.kchunks = a->Cin / 64;
  const uint32_t w_bytes = (uint32_t)(9 * p.kchunks * nt * 128);
  const uint32_t budget = 220u * 1024u - 1024u;
  if (w_bytes + 2u * HD_A_STAGE > budget) return PB_ERR_UNSUPPORTED;   // weights must stay resident next to >= 2 halo stages
  p.a_stages = (int)((budget - w_bytes) / HD_A_STAGE);
  if (p.a_stages > HD_MAX_A_STAGES) p.a_stages = HD_MAX_A_STAGES;
  p.b_off = (uint32_t)p.a_stages * HD_A_STAGE;
  p.N = a->N; p.IH = a->IH; p.IW = a->IW; p.OH = a->OH; p.OW = a->OW; p.Cout = a->Cout;
  const int groups_h = cdiv(a->IH, HD_TILE_H);
  p.groups_w = cdiv(a->IW, HD_TILE_W);
  p.tiles_per_img = groups_h * p.groups_w;
  p.total_tiles = a->N * p.tiles_per_img;
  if (p.total_tiles <= 0) return PB_OK;
  int grid = p.total_tiles < sm_count() ? p.total_tiles : sm_count();
  p.tiles_per_cta = cdiv(p.total_tiles, grid);
  grid = cdiv(p.total_tiles, p.tiles_per_cta);
  p.f16 = a->act_dtype == PB_F16 ? 1 : 0;
  p.slope = a->act == PB_ACT_LRELU ? a->slope : 1.f;
  if (!(p.slope > 0.f && p.slope <= 1.f)) return PB_ERR_UNSUPPORTED;   // the epilogue's LeakyReLU is max(v, slope * v)
  p.bias = a->bias;
  p.out = reinterpret_cast<float*>(a->out);
  if (head != nullptr) {
    p.keys = head->keys;
    p.target = head->target; p.points = head->points;
    p.negk2 = head->negk2; p.gscale = head->gscale;
    p.loss = head->loss; p.grad = reinterpret_cast<__nv_bfloat16*>(head->grad);
    p.dbias = head->dbias;
  }
  {
    const uint64_t C = (uint64_t)a->Cin;
    const uint64_t dims[4] = {C, (uint64_t)a->IW, (uint64_t)a->IH, (uint64_t)a->N};
    const uint64_t str[3] = {C * 2, (uint64_t)a->IW * C * 2, (uint64_t)a->IH * a->IW * C * 2};
    const uint32_t box[4] = {64, (uint32_t)HD_BOX_COLS, (uint32_t)HD_BOX_ROWS, 1};
    int rc = encode_tmap_bf16(&maps.a, a->in, 4, dims, str, box);
    if (rc != PB_OK) return rc;
    const uint64_t wdims[3] = {C, (uint64_t)nt, 9};
    const uint64_t wstr[2] = {C * 2, (uint64_t)nt * C * 2};
    const uint32_t wbox[3] = {64, (uint32_t)nt, 1};
    rc = encode_tmap_bf16(&maps.b, a->w, 3, wdims, wstr, wbox);
    if (rc != PB_OK) return rc;
  }
  const size_t smem = (size_t)p.b_off + w_bytes + 1024;
  switch (nt) {
    case 16: return head_launch<16>(mode, maps, p, grid, smem, stream);
    case 32: return head_launch<32>(mode, maps, p, grid, smem, stream);
    default: return head_launch<48>(mode, maps, p, grid, smem, stream);
  }
}

}  // namespace pb
