// tcgen05 weight gradient of a stride-2 transposed 3x3 convolution with a narrow output (the network head,
// ConvTranspose2d(128 -> C, k3, s2, p1, op1), pytorch/CNNs.py:125-128):
//
//   dW[t = (r, s)][ci][co] = sum over INPUT pixels i = (y, x):  a[y, x][ci] * g[2y + r - 1, 2x + s - 1][co]
//
// The per-tap kernel (tc_wgrad.cu) gives each tap its own CTAs, so the activation tensor is pulled through L2 nine
// times and the gradient tensor once per tap: 1.9 GB of L2 -> SM traffic for a 0.38 GB layer, 194 us.  Here one CTA
// owns ALL nine taps of a 128-row (ci) block for its pixel range, like tc_wgrad2.cu does for the stride-1 layers:
// per 16 x 8 input-pixel tile it loads the activation tile once and the gradient map as four PARITY boxes (the
// gradient pixels a tap touches all share the parity of (r - 1, s - 1); an odd parity needs one extra row / column
// for the r = 0 / s = 0 taps) through stride-2 tensor maps; the nine taps are nine (box, offset) views.
//
//   A (MN-major): [128 pixels][64 ci] x 2 blocks, LBO apart -> M = 128 channels
//   B (MN-major): parity box [(16 + py) x (8 + px) pixels][64 co, the channel padding zero-filled by TMA], N = C
//                 rounded up to 16; a 16-pixel K step = two box rows
//   D: nine accumulators of N fp32 columns in TMEM (9 x 48 = 432), kept across the CTA's whole pixel range
//
// Work item = (128-channel ci block, pixel split); fp32 partial tiles in pb_wgrad_reduce's layout.
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

namespace pb {

using namespace tc;

constexpr int WGU_THREADS = 192;
constexpr int WGU_A_BYTES = 2 * 128 * 128;   // two 64-channel blocks of [128 pixels][128 B]

struct WguMaps {
  CUtensorMap a;
  CUtensorMap g[4];   // parity (py, px) = index py * 2 + px: the gradient map sampled at stride 2
};

struct WguP {
  int N, tiles_h, tiles_w, total_tiles, tiles_per_split;
  int Ca, Cg, nco, units, ksplit, stages;
  uint32_t g_off[4], g_sbo[4];        // parity boxes inside a stage (behind the A tile), bytes between 8-pixel rows
  uint32_t tap_off[9];                // byte offset of the tap's first pixel inside its parity box
  int8_t tap_box[9];
  uint32_t g_tx, stage_bytes;
  float* partial;
  long long L;
};

__global__ void __launch_bounds__(WGU_THREADS, 1)
tc_wgrad_up_kernel(const __grid_constant__ WguMaps maps, const WguP p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[4];
  __shared__ __align__(8) uint64_t empty_bar[4];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int unit = (int)blockIdx.x % p.units;        // 128-channel block of ci
  const int split = (int)blockIdx.x / p.units;
  const int t_begin = split * p.tiles_per_split;
  const int t_end = min(p.total_tiles, t_begin + p.tiles_per_split);
  const int ntiles = max(0, t_end - t_begin);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.a);
    for (int i = 0; i < 4; ++i) prefetch_tmap(&maps.g[i]);
    for (int s = 0; s < 4; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      // ---------------------------------------------------------------- TMA producer
      int stage = 0;
      uint32_t phase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        int r = t;
        const int tw = r % p.tiles_w; r /= p.tiles_w;
        const int th = r % p.tiles_h;
        const int img = r / p.tiles_h;
        const int h0 = th * 16, w0 = tw * 8;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + (size_t)stage * p.stage_bytes;
        mbar_expect_tx(&full_bar[stage], (uint32_t)WGU_A_BYTES + p.g_tx);
        tma_load_4d(sa, &maps.a, &full_bar[stage], unit * 128, w0, h0, img);
        tma_load_4d(sa + 128 * 128, &maps.a, &full_bar[stage], unit * 128 + 64, w0, h0, img);
#pragma unroll
        for (int b = 0; b < 4; ++b)   // odd parities start one (parity-grid) row / column early: negative starts zero-fill
          tma_load_4d(sa + p.g_off[b], &maps.g[b], &full_bar[stage], 0, w0 - (b & 1), h0 - (b >> 1), img);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ---------------------------------------------------------------- MMA issuer
      const uint32_t idesc = make_idesc(128, p.nco, 1, 1);
      const uint64_t ad = smem_desc_sw128(smem_u32(smem), 128 * 128, 1024);
      const uint32_t a_lo0 = (uint32_t)ad, a_hi = (uint32_t)(ad >> 32);
      uint32_t g_lo0[9], g_hi[9], g_step[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int b = p.tap_box[t];
        const uint64_t gd = smem_desc_sw128(smem_u32(smem) + p.g_off[b] + p.tap_off[t], 16, p.g_sbo[b]);
        g_lo0[t] = (uint32_t)gd;
        g_hi[t] = (uint32_t)(gd >> 32);
        g_step[t] = (2u * p.g_sbo[b]) >> 4;          // 16 pixels per K step = two 8-pixel box rows
      }
      const uint32_t stage16 = p.stage_bytes >> 4;
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < ntiles; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t s16 = (uint32_t)stage * stage16;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const uint32_t d = tmem_base + (uint32_t)(t * p.nco);
          uint32_t g_lo = g_lo0[t] + s16;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            umma_bf16_lohi(d, a_lo0 + s16 + (uint32_t)j * 128u, a_hi, g_lo, g_hi[t], idesc, (it > 0 || j > 0) ? 1u : 0u);
            g_lo += g_step[t];
          }
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      umma_commit(&done_bar);
    }
  } else {
    // ------------------------------------------------------------------ epilogue: accumulators -> fp32 partial tiles
    const int q = warp & 3;
    const int m = q * 32 + lane;                       // accumulator row = channel of the block
    if (ntiles > 0) {
      mbar_wait(&done_bar, 0);
      tc_fence_after();
    }
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const int ci = unit * 128 + m;
    for (int t = 0; t < 9; ++t) {
      float* dst = p.partial + (long long)split * p.L + ((long long)t * p.Ca + ci) * p.Cg;
      for (int c0 = 0; c0 < p.nco; c0 += 16) {
        uint32_t r[16];
        if (ntiles > 0) {
          tmem_ld16(lane_base + (uint32_t)(t * p.nco + c0), r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) r[j] = 0u;
        }
        if (ci < p.Ca) {
          if (c0 + 16 <= p.Cg && (p.Cg & 3) == 0) {
#pragma unroll
            for (int v = 0; v < 4; ++v)
              *reinterpret_cast<uint4*>(dst + c0 + 4 * v) = make_uint4(r[4 * v], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c0 + j < p.Cg) dst[c0 + j] = __uint_as_float(r[j]);
          }
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

int launch_bias_partial(const pb_wgrad_args* a, cudaStream_t st);  // simt_conv.cu

// tc_wgrad.cu calls this first for stride-2 transposed convs; PB_ERR_UNSUPPORTED means "use the per-tap kernel"
int wgrad_tc_up(const pb_wgrad_args* a, cudaStream_t stream) {
  const char* off = getenv("POSEB200_WGRAD_UP");
  if (off != nullptr && off[0] == '0') return PB_ERR_UNSUPPORTED;
  const int gcs = a->g_cstride ? a->g_cstride : a->Cg;
  const int nco = cdiv(a->Cg, 16) * 16;
  if (a->act_dtype != PB_BF16 || a->a_nchw_f32 || a->mul_a != 1 || a->mul_g != 2 || a->ntaps != 9 || (a->Ca % 128) != 0 ||
      9 * nco > 512 || (gcs & 7) != 0 || gcs < a->Cg || a->PH < 16 || a->PW < 8 || a->GH != 2 * a->PH || a->GW != 2 * a->PW ||
      a->AH != a->PH || a->AW != a->PW)
    return PB_ERR_UNSUPPORTED;
  for (int t = 0; t < 9; ++t)
    if (a->dya[t] != 0 || a->dxa[t] != 0 || a->dyg[t] != t / 3 - 1 || a->dxg[t] != t % 3 - 1) return PB_ERR_UNSUPPORTED;
  WguP p;
  memset(&p, 0, sizeof(p));
  p.N = a->N;
  p.tiles_h = cdiv(a->PH, 16);
  p.tiles_w = cdiv(a->PW, 8);
  p.total_tiles = a->N * p.tiles_h * p.tiles_w;
  p.tiles_per_split = cdiv(p.total_tiles, a->ksplit);
  p.Ca = a->Ca; p.Cg = a->Cg; p.nco = nco;
  p.units = a->Ca / 128;
  p.ksplit = a->ksplit;
  p.partial = a->partial;
  p.L = (long long)a->ntaps * a->Ca * a->Cg + a->Cg;
  // parity boxes behind the activation tile
  uint32_t off_b = (uint32_t)WGU_A_BYTES;
  int box_rows[4], box_cols[4];
  for (int b = 0; b < 4; ++b) {
    box_rows[b] = 16 + (b >> 1);
    box_cols[b] = 8 + (b & 1);
    p.g_off[b] = off_b;
    p.g_sbo[b] = (uint32_t)box_cols[b] * 128u;
    const uint32_t bytes = (uint32_t)box_rows[b] * box_cols[b] * 128u;
    p.g_tx += bytes;
    off_b += (bytes + 1023u) & ~1023u;
  }
  p.stage_bytes = off_b;
  for (int t = 0; t < 9; ++t) {
    const int r = t / 3, s = t % 3;
    const int py = r != 1, px = s != 1;                 // parity of the gradient rows / columns the tap touches
    const int roff = (r == 2) ? 1 : 0, coff = (s == 2) ? 1 : 0;   // r = 0 reads parity row y - 1 = box row 0, r = 2 row y = box row 1
    const int b = py * 2 + px;
    p.tap_box[t] = (int8_t)b;
    p.tap_off[t] = (uint32_t)(roff * box_cols[b] + coff) * 128u;
  }
  static int dyn_max = 0;
  if (dyn_max == 0) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, tc_wgrad_up_kernel);
    if (e != cudaSuccess) return cuda_fail(e, "pb_wgrad_tc(up): func attributes");
    const int lim = 227 * 1024 - (int)fa.sharedSizeBytes;
    e = cudaFuncSetAttribute(tc_wgrad_up_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    if (e != cudaSuccess) return cuda_fail(e, "pb_wgrad_tc(up): smem attribute");
    dyn_max = lim;
  }
  p.stages = (int)(((uint32_t)dyn_max - 1024u) / p.stage_bytes);
  if (p.stages > 4) p.stages = 4;
  if (p.stages < 2) return PB_ERR_UNSUPPORTED;

  WguMaps maps;
  memset(&maps, 0, sizeof(maps));
  {
    const uint64_t C = (uint64_t)a->Ca;
    const uint64_t dims[4] = {C, (uint64_t)a->AW, (uint64_t)a->AH, (uint64_t)a->N};
    const uint64_t str[3] = {C * 2, (uint64_t)a->AW * C * 2, (uint64_t)a->AH * a->AW * C * 2};
    const uint32_t box[4] = {64, 8, 16, 1};
    int rc = encode_tmap_bf16(&maps.a, a->a, 4, dims, str, box);
    if (rc != PB_OK) return rc;
  }
  for (int b = 0; b < 4; ++b) {
    const int py = b >> 1, px = b & 1;
    const uint64_t C = (uint64_t)gcs;
    const __nv_bfloat16* base = (const __nv_bfloat16*)a->g + ((size_t)py * a->GW + px) * C;
    const uint64_t dims[4] = {C, (uint64_t)a->GW / 2, (uint64_t)a->GH / 2, (uint64_t)a->N};
    const uint64_t str[3] = {2 * C * 2, 2 * (uint64_t)a->GW * C * 2, (uint64_t)a->GH * a->GW * C * 2};
    const uint32_t box[4] = {64, (uint32_t)box_cols[b], (uint32_t)box_rows[b], 1};   // channels beyond C zero-fill
    int rc = encode_tmap_bf16(&maps.g[b], base, 4, dims, str, box);
    if (rc != PB_OK) return rc;
  }
  const size_t smem = (size_t)p.stages * p.stage_bytes + 1024;
  tc_wgrad_up_kernel<<<p.units * a->ksplit, WGU_THREADS, smem, stream>>>(maps, p);
  PB_LAUNCH_CHECK("tc_wgrad_up_kernel");
  if (a->want_bias) return launch_bias_partial(a, stream);
  return PB_OK;
}

}  // namespace pb
