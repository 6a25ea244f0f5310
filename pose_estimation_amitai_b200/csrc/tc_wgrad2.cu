// tcgen05 weight-gradient contraction, second generation: one CTA computes ALL taps of a
// 64(ci) x 64(co) block of dW from a shared-memory resident input halo.
//
//   dW[t][ci][co] = sum_pixels a(p + off(t))[ci] * g(p)[co]          (conv / stride-1 transposed conv)
//
// tc_wgrad.cu gives every (tap, 128-channel block) its own CTA, so the gradient tile g and the
// (shifted) activation tile are pulled through L2 once per tap: 2.7 GB of L2->SM traffic for a
// 0.3 GB layer (profiles/r1a_tc_wgrad_ncu_summary.txt, 32 % tensor-pipe active).  Here the reduction
// runs over pixel tiles of 16 rows x 8 columns (K = 128 pixels); per tile ONE halo box
// [(16+hy) x (8+hx) pixels x 64 ci] and ONE gradient box [128 pixels x 64 co] are loaded, and the
// nine taps are nine shifted views of the halo (MN-major UMMA descriptors: K = pixel index, an
// 8-pixel tile row = one 8-row swizzle group, SBO = box_cols*128 between tile rows).  Two taps share
// one M=128 accumulator: rows 0-63 are tap 2j, rows 64-127 tap 2j+1, the second 64-channel block
// simply sits LBO = off(2j+1) - off(2j) bytes further inside the same halo.  9 taps -> 5 accumulators
// of 64 fp32 columns in TMEM.  L2->SM bytes per MMA cycle drop ~5x.
//
// Work item = (ci block, co block, pixel split); fp32 partial tiles go to `partial` exactly like
// tc_wgrad.cu (pb_wgrad_reduce folds the splits).
//
// Wide mode (Cout % 128 == 0): the co block is 128 wide (N = 128), because a M=128 x N=64 MMA is paced by its 4 KB
// A-operand fetch (64 cycles) and not by its 32 cycles of math -- N = 128 does twice the work in the same time.
// Four tap pairs x 128 columns fill TMEM, so a layer with more than 8 taps is covered by two KINDS of CTA in the
// same launch: kind A owns pairs 0-3, kind B the remaining pair(s).  A kind-B CTA has a quarter of the MMAs per
// pixel tile, so it takes WG2_B_RATIO consecutive pixel splits (and zero-fills the partial slots it skips).
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

namespace pb {

using namespace tc;

constexpr int WG2_THREADS = 192;
constexpr int WG2_MAX_STAGES = 6;
constexpr int WG2_G_BYTES = 128 * 128;  // [128 pixels][64 co] bf16
constexpr int WG2_B_RATIO = 3;          // pixel splits per kind-B CTA (ops.py:choose_ksplit mirrors this)

struct Wg2Maps {
  CUtensorMap a, g;
};

struct Wg2P {
  int N, tiles_h, tiles_w, total_tiles, tiles_per_split;
  int ntaps, npairs, Ca, Cg, cob_n, units;
  int8_t tap_lo[PB_MAX_TAPS / 2 + 1], tap_hi[PB_MAX_TAPS / 2 + 1];  // per pair: rows 0-63 / 64-127 (offset order)
  uint32_t a_off[PB_MAX_TAPS];
  uint32_t sbo_a, a_bytes, a_tx, stage_bytes;
  int box_dx0, box_dy0, stages;
  float* partial;
  long long L;
  int want_bias;   // CTAs of ci block 0 also column-sum their gradient tiles (dbias) from shared memory
  int nco;         // co block width = UMMA N: 64, or 128 in wide mode
  int grp_pairs;   // tap pairs per CTA (512 / nco TMEM columns each)
  int ksplit, n_kind_a;   // pixel splits; CTAs of kind A (= units * ksplit)
};

__global__ void __launch_bounds__(WG2_THREADS, 1)
tc_wgrad2_kernel(const __grid_constant__ Wg2Maps maps, const Wg2P p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[WG2_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[WG2_MAX_STAGES];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float sred[16][128];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const bool kind_b = (int)blockIdx.x >= p.n_kind_a;
  const int bi = kind_b ? (int)blockIdx.x - p.n_kind_a : (int)blockIdx.x;
  const int unit = bi % p.units;
  const int split = (bi / p.units) * (kind_b ? WG2_B_RATIO : 1);      // partial slot this CTA writes
  const int nsplit = kind_b ? min(WG2_B_RATIO, p.ksplit - split) : 1;  // pixel splits it covers
  const int cib = unit / p.cob_n, cob = unit % p.cob_n;
  const int t_begin = split * p.tiles_per_split;
  const int t_end = min(p.total_tiles, t_begin + nsplit * p.tiles_per_split);
  const int ntiles = max(0, t_end - t_begin);
  const int pr_begin = kind_b ? p.grp_pairs : 0;
  const int pr_end = min(p.npairs, pr_begin + p.grp_pairs);
  const bool do_bias = p.want_bias != 0 && cib == 0 && !kind_b;
  const int nhalf = p.nco >> 6;   // 64-channel blocks of the gradient tile

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.a);
    prefetch_tmap(&maps.g);
    for (int s = 0; s < WG2_MAX_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], do_bias ? 5 : 1);  // MMA commit (+ the four column-summing warps)
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
        mbar_expect_tx(&full_bar[stage], p.a_tx + (uint32_t)nhalf * WG2_G_BYTES);
        tma_load_4d(sa, &maps.a, &full_bar[stage], cib * 64, w0 + p.box_dx0, h0 + p.box_dy0, img);
        for (int h = 0; h < nhalf; ++h)
          tma_load_4d(sa + p.a_bytes + h * WG2_G_BYTES, &maps.g, &full_bar[stage], (cob * nhalf + h) * 64, w0, h0, img);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc(128, p.nco, 1, 1);
      // descriptors of stage 0 (only the 14-bit start-address field of the low word moves afterwards): per pair the
      // halo view of its first tap (LBO = distance to its second tap); the MN-major gradient tile (64-channel blocks
      // WG2_G_BYTES apart = LBO, 16 pixels = 2048 B per K step).  The issue loop is then 32-bit adds only.
      uint32_t pa_lo[PB_MAX_TAPS / 2 + 1], pa_hi[PB_MAX_TAPS / 2 + 1];
#pragma unroll
      for (int i = 0; i < PB_MAX_TAPS / 2 + 1; ++i) {
        const int pr = min(pr_begin + i, p.npairs - 1);
        const int t1 = p.tap_lo[pr], t2 = p.tap_hi[pr];
        const uint64_t d = smem_desc_sw128(smem_u32(smem) + p.a_off[t1], p.a_off[t2] - p.a_off[t1], p.sbo_a);
        pa_lo[i] = (uint32_t)d;
        pa_hi[i] = (uint32_t)(d >> 32);
      }
      const uint64_t gd = smem_desc_sw128(smem_u32(smem) + p.a_bytes, WG2_G_BYTES, 1024);
      const uint32_t g_lo0 = (uint32_t)gd, g_hi = (uint32_t)(gd >> 32);
      const uint32_t a_step16 = (2u * p.sbo_a) >> 4;   // two tile rows = 16 pixels per K step
      const uint32_t stage16 = p.stage_bytes >> 4;
      const int npr = pr_end - pr_begin;
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < ntiles; ++t) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t s16 = (uint32_t)stage * stage16;
        const uint32_t g_lo = g_lo0 + s16;
#pragma unroll
        for (int i = 0; i < PB_MAX_TAPS / 2 + 1; ++i) {
          if (i < npr) {
            uint32_t a_lo = pa_lo[i] + s16;
            const uint32_t d_tmem = tmem_base + (uint32_t)(i * p.nco);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              umma_bf16_lohi(d_tmem, a_lo, pa_hi[i], g_lo + (uint32_t)j * 128u, g_hi, idesc, (t > 0 || j > 0) ? 1u : 0u);
              a_lo += a_step16;
            }
          }
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      umma_commit(&done_bar);
    }
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    if (do_bias) {
      // ---- dbias: column sums of every gradient tile, read from the same shared-memory stage the MMAs use
      const int et = (warp - 2) * 32 + lane;      // 0..127
      const int j = et & 7, rg = et >> 3;         // 16-byte channel chunk, group of 8 pixel rows
      float acc[2][8];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[0][k] = acc[1][k] = 0.f;
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < ntiles; ++t) {
        mbar_wait(&full_bar[stage], phase);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (h < nhalf) {
            const uint8_t* gt = smem + (size_t)stage * p.stage_bytes + p.a_bytes + h * WG2_G_BYTES + rg * 1024;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint4 v = *reinterpret_cast<const uint4*>(gt + i * 128 + ((j ^ i) << 4));
              const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                acc[h][2 * k] += bf16lo(w[k]);
                acc[h][2 * k + 1] += bf16hi(w[k]);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int k = 0; k < 8; ++k) sred[rg][h * 64 + j * 8 + k] = acc[h][k];
      asm volatile("bar.sync 1, 128;" ::: "memory");   // the four epilogue warps only
      if (et < p.nco) {
        float s = 0.f;
#pragma unroll
        for (int g = 0; g < 16; ++g) s += sred[g][et];
        const int c = cob * p.nco + et;
        if (c < p.Cg) p.partial[(long long)split * p.L + (p.L - p.Cg) + c] = s;
      }
    }
    if (ntiles > 0) {
      mbar_wait(&done_bar, 0);
      tc_fence_after();
    }
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const int ci = cib * 64 + (m & 63);
    for (int pr = pr_begin; pr < pr_end; ++pr) {
      const int tap = (m >> 6) ? p.tap_hi[pr] : p.tap_lo[pr];
      const bool row_ok = ((m >> 6) == 0 || p.tap_hi[pr] != p.tap_lo[pr]) && ci < p.Ca;
      float* dst = p.partial + (long long)split * p.L + ((long long)tap * p.Ca + ci) * p.Cg + cob * p.nco;
      for (int h = 0; h < 2 * nhalf; ++h) {
        uint32_t r[32];
        if (ntiles > 0) {
          tmem_ld32(lane_base + (uint32_t)((pr - pr_begin) * p.nco + h * 32), r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0u;
        }
        if (row_ok) {
          const int c0 = cob * p.nco + h * 32;
          if (c0 + 32 <= p.Cg && (p.Cg & 3) == 0) {
#pragma unroll
            for (int v = 0; v < 8; ++v)
              *reinterpret_cast<uint4*>(dst + h * 32 + 4 * v) = make_uint4(r[4 * v], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]);
            for (int s2 = 1; s2 < nsplit; ++s2)   // kind B: the splits folded into this CTA contribute nothing more
#pragma unroll
              for (int v = 0; v < 8; ++v)
                *reinterpret_cast<uint4*>(dst + (long long)s2 * p.L + h * 32 + 4 * v) = make_uint4(0u, 0u, 0u, 0u);
          } else {
            for (int j = 0; j < 32; ++j)
              if (c0 + j < p.Cg) {
                dst[h * 32 + j] = __uint_as_float(r[j]);
                for (int s2 = 1; s2 < nsplit; ++s2) dst[(long long)s2 * p.L + h * 32 + j] = 0.f;
              }
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

// tc_wgrad.cu calls this first; PB_ERR_UNSUPPORTED means "use the per-tap kernel"
int wgrad_tc_v2(const pb_wgrad_args* a, cudaStream_t stream) {
  const char* off = getenv("POSEB200_WGRAD_V1");
  if (off != nullptr && off[0] == '1') return PB_ERR_UNSUPPORTED;
  const int gcs = a->g_cstride ? a->g_cstride : a->Cg;
  if (a->act_dtype != PB_BF16 || a->a_nchw_f32 || a->mul_a != 1 || a->mul_g != 1 || (a->Ca & 7) != 0 || (gcs & 7) != 0 ||
      a->PH < 16 || a->PW < 8)
    return PB_ERR_UNSUPPORTED;
  for (int t = 0; t < a->ntaps; ++t)
    if (a->dyg[t] != 0 || a->dxg[t] != 0) return PB_ERR_UNSUPPORTED;
  int ymin = 127, ymax = -127, xmin = 127, xmax = -127;
  for (int t = 0; t < a->ntaps; ++t) {
    ymin = a->dya[t] < ymin ? a->dya[t] : ymin; ymax = a->dya[t] > ymax ? a->dya[t] : ymax;
    xmin = a->dxa[t] < xmin ? a->dxa[t] : xmin; xmax = a->dxa[t] > xmax ? a->dxa[t] : xmax;
  }
  Wg2P p;
  memset(&p, 0, sizeof(p));
  const int box_cols = 8 + (xmax - xmin), box_rows = 16 + (ymax - ymin);
  if (box_cols > 256 || box_rows > 256) return PB_ERR_UNSUPPORTED;
  p.N = a->N;
  p.tiles_h = cdiv(a->PH, 16);
  p.tiles_w = cdiv(a->PW, 8);
  p.total_tiles = a->N * p.tiles_h * p.tiles_w;
  p.tiles_per_split = cdiv(p.total_tiles, a->ksplit);
  p.ntaps = a->ntaps;
  p.npairs = (a->ntaps + 1) / 2;
  p.Ca = a->Ca; p.Cg = a->Cg;
  {
    const char* nw = getenv("POSEB200_WGRAD_NARROW");
    const bool wide = (a->Cg % 128) == 0 && gcs >= a->Cg && !(nw != nullptr && nw[0] == '1');
    p.nco = wide ? 128 : 64;
  }
  p.grp_pairs = 512 / p.nco;
  if (p.npairs > 2 * p.grp_pairs) return PB_ERR_UNSUPPORTED;
  p.cob_n = cdiv(a->Cg, p.nco);
  p.units = cdiv(a->Ca, 64) * p.cob_n;
  p.ksplit = a->ksplit;
  p.n_kind_a = p.units * a->ksplit;
  p.sbo_a = (uint32_t)box_cols * 128u;
  p.box_dx0 = xmin; p.box_dy0 = ymin;
  for (int t = 0; t < a->ntaps; ++t) {
    p.a_off[t] = (uint32_t)((a->dya[t] - ymin) * box_cols + (a->dxa[t] - xmin)) * 128u;
  }
  for (int pr = 0; pr < p.npairs; ++pr) {
    // the second 64-channel block of a pair must lie at a non-negative byte distance (LBO is unsigned)
    const int t1 = 2 * pr, t2 = (2 * pr + 1 < a->ntaps) ? 2 * pr + 1 : 2 * pr;
    const bool swap = p.a_off[t2] < p.a_off[t1];
    p.tap_lo[pr] = (int8_t)(swap ? t2 : t1);
    p.tap_hi[pr] = (int8_t)(swap ? t1 : t2);
  }
  p.a_tx = (uint32_t)box_cols * box_rows * 128u;
  p.a_bytes = (p.a_tx + 1023u) & ~1023u;
  p.stage_bytes = p.a_bytes + (uint32_t)(p.nco / 64) * WG2_G_BYTES;
  p.partial = a->partial;
  p.L = (long long)a->ntaps * a->Ca * a->Cg + a->Cg;
  p.want_bias = a->want_bias;

  static int dyn_max = 0;
  if (dyn_max == 0) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, tc_wgrad2_kernel);
    if (e != cudaSuccess) return cuda_fail(e, "pb_wgrad_tc(v2): func attributes");
    const int lim = 227 * 1024 - (int)fa.sharedSizeBytes;
    e = cudaFuncSetAttribute(tc_wgrad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    if (e != cudaSuccess) return cuda_fail(e, "pb_wgrad_tc(v2): smem attribute");
    dyn_max = lim;
  }
  p.stages = (int)(((uint32_t)dyn_max - 1024u) / p.stage_bytes);
  if (p.stages > WG2_MAX_STAGES) p.stages = WG2_MAX_STAGES;
  if (p.stages < 2) return PB_ERR_UNSUPPORTED;

  Wg2Maps maps;
  memset(&maps, 0, sizeof(maps));
  {
    const uint64_t C = (uint64_t)a->Ca;
    const uint64_t dims[4] = {C, (uint64_t)a->AW, (uint64_t)a->AH, (uint64_t)a->N};
    const uint64_t str[3] = {C * 2, (uint64_t)a->AW * C * 2, (uint64_t)a->AH * a->AW * C * 2};
    const uint32_t box[4] = {64, (uint32_t)box_cols, (uint32_t)box_rows, 1};
    int rc = encode_tmap_bf16(&maps.a, a->a, 4, dims, str, box);
    if (rc != PB_OK) return rc;
  }
  {
    const uint64_t C = (uint64_t)gcs;
    const uint64_t dims[4] = {C, (uint64_t)a->GW, (uint64_t)a->GH, (uint64_t)a->N};
    const uint64_t str[3] = {C * 2, (uint64_t)a->GW * C * 2, (uint64_t)a->GH * a->GW * C * 2};
    const uint32_t box[4] = {64, 8, 16, 1};
    int rc = encode_tmap_bf16(&maps.g, a->g, 4, dims, str, box);
    if (rc != PB_OK) return rc;
  }
  const size_t smem = (size_t)p.stages * p.stage_bytes + 1024;
  const int n_kind_b = p.npairs > p.grp_pairs ? p.units * cdiv(a->ksplit, WG2_B_RATIO) : 0;
  tc_wgrad2_kernel<<<p.n_kind_a + n_kind_b, WG2_THREADS, smem, stream>>>(maps, p);
  PB_LAUNCH_CHECK("tc_wgrad2_kernel");
  return PB_OK;
}

}  // namespace pb
