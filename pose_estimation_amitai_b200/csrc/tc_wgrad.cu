// tcgen05 weight-gradient contraction:  dW[t][ci][co] = sum_pixels a(p + off_a(t))[ci] * g(p + off_g(t))[co]
//
// The reduction dimension (K) is the pixel index, which is the SLOW axis of both NHWC operands, so
// both UMMA operands are MN-major: a TMA box [64 channels x TH x TW pixels] lands in shared memory as
// [K = pixels][128 B of channels] rows with the 128-byte swizzle -- exactly the canonical MN-major
// SWIZZLE_128B layout (8-row groups 1024 B apart = SBO, 64-channel blocks LBO apart).  No transposed
// copy of any activation is ever materialised.
//
// Work item = (unit, pixel split).  A unit is 128 accumulator rows:
//   Cin >= 128 : (tap t, 128-channel block of ci)
//   Cin == 64  : a PAIR of taps (2j, 2j+1) x 64 ci  -- the two 64-row halves are two differently
//                shifted activation boxes placed LBO apart, so M stays 128 (full-rate MMA)
// N = Cout (<= 256).  Each CTA accumulates its pixel range in TMEM and writes one fp32 partial
// tile; pb_wgrad_reduce sums the splits into the parameter-shaped gradient.
#include <string.h>

#include "tc_common.cuh"

namespace pb {

using namespace tc;

constexpr int WG_THREADS = 192;
constexpr int WG_KPIX = 64;                 // pixels per K chunk (one TMA box per 64-channel block)
constexpr int WG_BLK_BYTES = WG_KPIX * 128; // 8 KB: [64 pixels][64 channels] bf16
constexpr int WG_MAX_STAGES = 8;

struct WgMaps {
  CUtensorMap a;
  CUtensorMap g[4];
};

struct WgTap {
  int8_t ady, adx, gdy, gdx, gmap;
};

struct WgP {
  int N, TH, TW, tiles_h, tiles_w;
  int ntaps, Ca, Cg, n_mma, nb;       // n_mma: UMMA N (multiple of 16), nb: 64-channel blocks of g
  int pair_mode;                      // 1: Cin == 64, unit = tap pair
  int units, ksplit, chunks_per_split, total_chunks;
  int base_units, ncg, cg_chunk;      // Cg > 256 (wide layers): ncg column chunks of cg_chunk = 256 or 128, units = base_units * ncg
  int stages;
  uint32_t stage_bytes;
  WgTap taps[PB_MAX_TAPS];
  float* partial;
  long long L;                        // floats per split: ntaps*Ca*Cg + Cg
};

__global__ void __launch_bounds__(WG_THREADS, 1)
tc_wgrad_kernel(const __grid_constant__ WgMaps maps, const WgP p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[WG_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[WG_MAX_STAGES];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int unit_all = blockIdx.x % p.units;
  const int split = blockIdx.x / p.units;
  const int unit = unit_all % p.base_units;
  const int gcol0 = (unit_all / p.base_units) * p.cg_chunk;   // first g channel of this unit's column chunk
  const int c_begin = split * p.chunks_per_split;
  const int c_end = min(p.total_chunks, c_begin + p.chunks_per_split);
  const int nchunks = max(0, c_end - c_begin);

  // which taps / channel block feed the two 64-row halves of this unit
  int tap_lo, tap_hi, ci_lo, ci_hi;
  if (p.pair_mode) {
    tap_lo = 2 * unit;
    tap_hi = min(2 * unit + 1, p.ntaps - 1);
    ci_lo = ci_hi = 0;
  } else {
    const int cblocks = p.Ca / 128;
    tap_lo = tap_hi = unit / cblocks;
    ci_lo = (unit % cblocks) * 128;
    ci_hi = ci_lo + 64;
  }

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.a);
    for (int i = 0; i < 4; ++i) prefetch_tmap(&maps.g[i]);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      const WgTap tl = p.taps[tap_lo], th = p.taps[tap_hi];
      int stage = 0;
      uint32_t phase = 0;
      for (int c = c_begin; c < c_end; ++c) {
        int r = c;
        const int tw = r % p.tiles_w; r /= p.tiles_w;
        const int thh = r % p.tiles_h;
        const int img = r / p.tiles_h;
        const int h0 = thh * p.TH, w0 = tw * p.TW;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + (size_t)stage * p.stage_bytes;
        uint8_t* sb = sa + 2 * WG_BLK_BYTES;
        mbar_expect_tx(&full_bar[stage], p.stage_bytes);
        tma_load_4d(sa, &maps.a, &full_bar[stage], ci_lo, w0 + tl.adx, h0 + tl.ady, img);
        tma_load_4d(sa + WG_BLK_BYTES, &maps.a, &full_bar[stage], ci_hi, w0 + th.adx, h0 + th.ady, img);
        // in pair mode both taps must see the SAME g pixels: only the a-side is shifted (plain convs)
        for (int b = 0; b < p.nb; ++b)
          tma_load_4d(sb + b * WG_BLK_BYTES, &maps.g[tl.gmap], &full_bar[stage], gcol0 + b * 64, w0 + tl.gdx,
                      h0 + tl.gdy, img);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc(128, p.n_mma, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + (size_t)stage * p.stage_bytes);
        const uint32_t b_addr = a_addr + 2 * WG_BLK_BYTES;
#pragma unroll
        for (int j = 0; j < WG_KPIX / 16; ++j) {
          const uint64_t ad = smem_desc_sw128(a_addr + j * 2048, WG_BLK_BYTES, 1024);
          const uint64_t bd = smem_desc_sw128(b_addr + j * 2048, WG_BLK_BYTES, 1024);
          umma_bf16(tmem_base, ad, bd, idesc, (c > 0 || j > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      umma_commit(&done_bar);
    }
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    if (nchunks > 0) {
      mbar_wait(&done_bar, 0);
      tc_fence_after();
    }
    int tap, ci;
    bool row_ok = true;
    if (p.pair_mode) {
      tap = 2 * unit + (m >> 6);
      ci = m & 63;
      row_ok = tap < p.ntaps;
    } else {
      tap = tap_lo;
      ci = ci_lo + m;
    }
    float* dst = p.partial + (long long)split * p.L + ((long long)tap * p.Ca + ci) * p.Cg + gcol0;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int c0 = 0; c0 < p.n_mma; c0 += 16) {
      uint32_t r[16];
      if (nchunks > 0) {
        tmem_ld16(lane_base + c0, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = 0u;
      }
      if (row_ok) {
        if (gcol0 + c0 + 16 <= p.Cg && (p.Cg & 3) == 0) {
#pragma unroll
          for (int v = 0; v < 4; ++v)
            *reinterpret_cast<uint4*>(dst + c0 + 4 * v) = make_uint4(r[4 * v], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]);
        } else {
          for (int j = 0; j < 16; ++j)
            if (gcol0 + c0 + j < p.Cg) dst[c0 + j] = __uint_as_float(r[j]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

int launch_bias_partial(const pb_wgrad_args* a, cudaStream_t st);  // simt_conv.cu
int wgrad_tc_v2(const pb_wgrad_args* a, cudaStream_t stream);      // tc_wgrad2.cu
int wgrad_tc_up(const pb_wgrad_args* a, cudaStream_t stream);      // tc_wgrad_up.cu

}  // namespace pb

using namespace pb;

extern "C" int pb_wgrad_tc(const pb_wgrad_args* a, void* stream) {
  PB_REQUIRE(a != nullptr && a->a && a->g && a->partial, "pb_wgrad_tc: null args");
  PB_REQUIRE(a->N > 0 && a->PH > 0 && a->PW > 0 && a->ksplit >= 1 && a->ntaps >= 1 && a->ntaps <= PB_MAX_TAPS,
             "pb_wgrad_tc: bad shape");
  PB_REQUIRE_DEV(a->a, "a");
  PB_REQUIRE_DEV(a->g, "g");
  PB_REQUIRE_DEV(a->partial, "partial");
  {
    const int rc2 = wgrad_tc_v2(a, (cudaStream_t)stream);  // halo-resident kernel; falls through when it does not tile the shape
    if (rc2 != PB_ERR_UNSUPPORTED) return rc2;
    const int rc3 = wgrad_tc_up(a, (cudaStream_t)stream);  // all nine taps of a narrow stride-2 transposed conv (the head) per CTA
    if (rc3 != PB_ERR_UNSUPPORTED) return rc3;
  }
  const int gcs = a->g_cstride ? a->g_cstride : a->Cg;
  if (a->act_dtype != PB_BF16 || a->a_nchw_f32 || a->Ca % 64 != 0 || (a->Ca > 64 && a->Ca % 128 != 0) ||
      gcs % 8 != 0 || gcs < a->Cg || (a->Cg > 256 && a->Cg % 128 != 0) || a->mul_a != 1 ||
      (a->mul_g != 1 && a->mul_g != 2)) {
    set_error("pb_wgrad_tc: shape outside the tcgen05 tiling (Ca=%d Cg=%d mul_a=%d mul_g=%d)", a->Ca, a->Cg, a->mul_a,
              a->mul_g);
    return PB_ERR_UNSUPPORTED;
  }
  WgP p;
  memset(&p, 0, sizeof(p));
  p.N = a->N;
  {
    const int cand[][2] = {{4, 16}, {8, 8}, {2, 32}, {16, 4}, {1, 64}};
    long long best = -1;
    for (auto& c : cand) {
      const long long cover = (long long)cdiv(a->PH, c[0]) * c[0] * cdiv(a->PW, c[1]) * c[1];
      if (best < 0 || cover < best) { best = cover; p.TH = c[0]; p.TW = c[1]; }
    }
  }
  p.tiles_h = cdiv(a->PH, p.TH);
  p.tiles_w = cdiv(a->PW, p.TW);
  p.ntaps = a->ntaps; p.Ca = a->Ca; p.Cg = a->Cg;
  const int cg_chunk = a->Cg <= 256 ? a->Cg : (a->Cg % 256 == 0 ? 256 : 128);
  p.ncg = a->Cg / cg_chunk;
  p.cg_chunk = cg_chunk;
  p.n_mma = cdiv(cg_chunk, 16) * 16;
  p.nb = cdiv(cg_chunk, 64);
  p.pair_mode = (a->Ca == 64) ? 1 : 0;
  p.base_units = p.pair_mode ? (a->ntaps + 1) / 2 : a->ntaps * (a->Ca / 128);
  p.units = p.base_units * p.ncg;
  p.ksplit = a->ksplit;
  p.total_chunks = a->N * p.tiles_h * p.tiles_w;
  p.chunks_per_split = cdiv(p.total_chunks, a->ksplit);
  p.stage_bytes = (uint32_t)((2 + p.nb) * WG_BLK_BYTES);
  p.stages = (int)((200 * 1024) / p.stage_bytes);
  if (p.stages > WG_MAX_STAGES) p.stages = WG_MAX_STAGES;
  p.partial = a->partial;
  p.L = (long long)a->ntaps * a->Ca * a->Cg + a->Cg;
  for (int t = 0; t < a->ntaps; ++t) {
    WgTap& k = p.taps[t];
    k.ady = a->dya[t]; k.adx = a->dxa[t];
    if (a->mul_g == 1) {
      k.gdy = a->dyg[t]; k.gdx = a->dxg[t]; k.gmap = 0;
    } else {
      const int py = a->dyg[t] & 1, px = a->dxg[t] & 1;
      k.gdy = (int8_t)((a->dyg[t] - py) / 2); k.gdx = (int8_t)((a->dxg[t] - px) / 2);
      k.gmap = (int8_t)(py * 2 + px);
    }
  }
  if (p.pair_mode) {
    // the two taps of a pair share one g tile: requires identical g offsets (true for plain convs)
    for (int t = 0; t + 1 < a->ntaps; t += 2)
      if (p.taps[t].gdy != p.taps[t + 1].gdy || p.taps[t].gdx != p.taps[t + 1].gdx ||
          p.taps[t].gmap != p.taps[t + 1].gmap) {
        set_error("pb_wgrad_tc: Ca == 64 needs tap-invariant g offsets");
        return PB_ERR_UNSUPPORTED;
      }
  }
  WgMaps maps;
  memset(&maps, 0, sizeof(maps));
  int rc;
  {
    const uint64_t C = (uint64_t)a->Ca;
    const uint64_t dims[4] = {C, (uint64_t)a->AW, (uint64_t)a->AH, (uint64_t)a->N};
    const uint64_t str[3] = {C * 2, (uint64_t)a->AW * C * 2, (uint64_t)a->AH * a->AW * C * 2};
    const uint32_t box[4] = {64, (uint32_t)p.TW, (uint32_t)p.TH, 1};
    rc = encode_tmap_bf16(&maps.a, a->a, 4, dims, str, box);
    if (rc != PB_OK) return rc;
  }
  {
    const uint64_t C = (uint64_t)gcs;
    const uint32_t box[4] = {64, (uint32_t)p.TW, (uint32_t)p.TH, 1};
    if (a->mul_g == 1) {
      const uint64_t dims[4] = {C, (uint64_t)a->GW, (uint64_t)a->GH, (uint64_t)a->N};
      const uint64_t str[3] = {C * 2, (uint64_t)a->GW * C * 2, (uint64_t)a->GH * a->GW * C * 2};
      rc = encode_tmap_bf16(&maps.g[0], a->g, 4, dims, str, box);
      if (rc != PB_OK) return rc;
      for (int i = 1; i < 4; ++i) maps.g[i] = maps.g[0];
    } else {
      for (int ph = 0; ph < 4; ++ph) {
        const int py = ph >> 1, px = ph & 1;
        const __nv_bfloat16* base = (const __nv_bfloat16*)a->g + ((size_t)py * a->GW + px) * C;
        const uint64_t dims[4] = {C, (uint64_t)a->GW / 2, (uint64_t)a->GH / 2, (uint64_t)a->N};
        const uint64_t str[3] = {2 * C * 2, 2 * (uint64_t)a->GW * C * 2, (uint64_t)a->GH * a->GW * C * 2};
        rc = encode_tmap_bf16(&maps.g[ph], base, 4, dims, str, box);
        if (rc != PB_OK) return rc;
      }
    }
  }
  const size_t smem = (size_t)p.stages * p.stage_bytes + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, "pb_wgrad_tc: smem attribute");
    attr_set = true;
  }
  tc_wgrad_kernel<<<p.units * a->ksplit, WG_THREADS, smem, (cudaStream_t)stream>>>(maps, p);
  PB_LAUNCH_CHECK("tc_wgrad_kernel");
  if (a->want_bias) return launch_bias_partial(a, (cudaStream_t)stream);
  return PB_OK;
}
