// tcgen05 implicit-GEMM gather-convolution, second generation: shared-memory resident input halo.
//
// tc_conv.cu re-fetches one shifted activation box per tap from L2, so a 9-tap layer pulls every
// input pixel into shared memory nine times and the L2->SM fabric (~40 B/cycle/SM), not the tensor
// pipe, sets the pace (r1 ncu: 15-47 % tensor-pipe active).  Here a group of T output tiles
// (each 16 rows x 8 columns = 128 pixels = one UMMA M) loads its input halo
// [(16+hy) x (8T+hx) pixels x 64 channels] ONCE per 64-channel K chunk; the A operand of every tap
// and tile is the same shared-memory box addressed through a shifted UMMA descriptor:
//   an 8-pixel tile row is 8 consecutive 128-byte rows of the box (one swizzle group),
//   consecutive tile rows are one box row apart  ->  SBO = box_cols * 128,
//   tap (dy,dx) / tile t only move the start address by ((dy*box_cols + dx + 8t) * 128) bytes;
//   the descriptor's base-offset field carries (start >> 7) & 7 so the 128-byte swizzle phase of
//   the (not 1024-byte aligned) start matches what TMA wrote.
// The weight tile of a (tap, K chunk) is streamed once per GROUP (reused by all T tiles), or, when
// the whole packed weight tensor fits next to the halo ring (Cin = Cout = 64, the narrow last
// layer), loaded once per CTA and kept resident.
//
// Warp roles (224 threads, one persistent CTA per SM):
//   warp 0  halo producer (TMA)      warp 1  MMA issuer        warps 2-5  epilogue
//   warp 6  weight producer (TMA)
// Geometries as in tc_conv.cu: plain, up (stride-2 transposed conv: 4 phase accumulators per tile),
// down (its input gradient: one halo box per input phase).
#include <stdlib.h>
#include <string.h>

#include "tc_epilogue.cuh"
#include "tc_head.cuh"   // V2Head (host-side carrier of the head_* fields below) + the folded-parity head kernel

namespace pb {

using namespace tc;

constexpr int V2_THREADS = 384;   // 12 warps: 0 halo / 1 MMA / 2-5 epilogue A / 6 weights / 7 epilogue operand / 8-11 epilogue B
constexpr int V2_MAX_E_STAGES = 4;
constexpr int V2_E_BYTES = 128 * 128;  // one epilogue-operand box: 128 pixels x 64 channels bf16
constexpr int V2_MAX_A_STAGES = 4;
constexpr int V2_MAX_B_STAGES = 8;
constexpr int V2_MAX_BOXES = 4;
constexpr int V2_TILE_H = 16, V2_TILE_W = 8;
constexpr int V2_UNROLL_TAPS = 9;   // 3x3 layers: the MMA issue loop is unrolled over this many taps

struct V2Maps {
  CUtensorMap a[V2_MAX_BOXES];
  CUtensorMap b;
  CUtensorMap e;   // skip / residual tensor of the epilogue (e_mode)
  CUtensorMap o;   // output tensor, stored by TMA from the epilogue's shared-memory tile (e_mode): box [64 ch][8][4]
  CUtensorMap pl;  // 2x2-pooled output tensor (pool): box [64 ch][4][2]
};

struct V2Box {
  int32_t dx0, dy0;     // box origin relative to the group origin (base-grid pixels)
  uint32_t smem_off;    // byte offset inside an A stage (1024-aligned)
};

struct V2Tap {
  uint32_t a_off;       // byte offset (inside an A stage) of tile 0's first row for this tap
  uint32_t sbo;         // bytes between consecutive 8-pixel tile rows
  int32_t acc;          // accumulator inside the pass
  int32_t wtap;         // tap index inside the packed weight tensor
};

struct V2P : EpiP {
  int N, BH, BW, groups_h, groups_w, T;
  int kchunks, ntaps, nboxes;
  V2Box boxes[V2_MAX_BOXES];
  V2Tap taps[PB_MAX_TAPS];
  int n_acc, n_tile, acc_stages, a_stages, b_stages, b_resident;
  uint32_t a_stage_bytes, a_tx_bytes, b_bytes, b_ring_off;
  int up, OH, OW, out_nchw;
  int use_base_offset;
  // cluster > 1: the CTAs of a cluster each fetch 1/cluster of every streamed weight tile and TMA-multicast it to
  // all of them (L2->SM weight traffic / cluster); every CTA of the grid then runs the same number of iterations
  int cluster, iters;
  // pair: the two CTAs of a cluster run ONE tcgen05.mma.cta_group::2 per step: M = 256 = the 128-pixel tiles of both
  // CTAs, N = n_tile with each CTA holding (and fetching) only half of every weight tile.  Per CTA the tensor core then
  // reads 4 KB of A + N*16 B of B per MMA instead of 4 KB + N*32 B, and the weight stream into shared memory halves --
  // the shared-memory port (128 B/clk), not the tensor pipe, is what paces the N >= 128 layers with cta_group::1.
  // The leader (rank 0) issues; full barriers live in the leader, empty / accumulator-full barriers are multicast.
  int pair;
  int a_keep_l2;   // the epilogue's residual operand is this layer's input: halo loads carry an L2 evict_last hint
  // stride-2 transposed convs: the four output phases are computed in npass passes over the same pixel group so
  // that the phases of one pass (n_acc of them) double-buffer in TMEM; taps are sorted by pass
  int npass;
  int pass_begin[5];
  int unroll_taps;   // 1: compile-time unrolled tap loop in the MMA issuer (default); 0: rolled (POSEB200_CONV_UNROLL=0)
  uint32_t poll_ns;   // back-off of the epilogue / producer warps' barrier polls (see mbar_wait_relaxed)
  int debug;   // timing experiments only (POSEB200_CONV_DEBUG): 1 no epilogue work, 2 no weight stream, 4 no halo stream
  // e_mode 1: the epilogue's skip (add0) or residual (add1) operand is staged by TMA, one
  // [16 x 8 pixels x 64 channels] box per (tile, 64-channel block), e_stages deep
  int e_mode, e_has_add, e_stages, e_is_add1;
  uint32_t e_ring_off;
  uint32_t smask_off;   // cp.async staging of the LeakyReLU' mask words: [2][128 threads][8 words]
  // pool: the staged epilogue also emits lrelu(maxpool2x2(out)) (CNNs.py:77,82) -- each TMEM quadrant's 4 x 8 result
  // pixels, still in shared memory for the TMA store, are pooled to 2 x 4 by the quadrant's two warps and leave as a
  // second TMA store in the same bulk group; pool_only: the full-resolution store is skipped (inference)
  int pool, pool_only;
  uint32_t pool_off;    // staging of the pooled rows: [3 ring slots][4 quadrants][8 pooled pixels][128 B]
  // fused network head (out_nchw geometry): the epilogue consumes the heatmaps instead of storing them.
  //   head_mode 1: per-(image, channel) arg-max keys (pb_convT_argmax_fused)
  //   head_mode 2: MSE loss + bf16 NHWC gradient against a target tensor or Gaussian targets (pb_convT_mse_fused)
  int head_mode;
  unsigned long long* head_keys;
  const float* head_target;
  const float* head_points;
  float head_negk2, head_gscale;
  float* head_loss;
  __nv_bfloat16* head_grad;
  int head_cpad;
};

// like smem_desc_sw128 but valid for a start address that is only 128-byte aligned
__device__ __forceinline__ uint64_t smem_desc_sw128_any(uint32_t saddr, uint32_t sbo_bytes, int use_base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                            // LBO (unused for K-major swizzled)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
  if (use_base_offset) d |= (uint64_t)((saddr >> 7) & 7) << 49;  // matrix base offset: swizzle phase of the start row
  d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// one slice of a weight tile, written to the same shared-memory offset of every CTA in `mask`; each
// destination CTA's mbarrier (same offset) receives the complete_tx
__device__ __forceinline__ void tma_load_3d_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(smem_u32(dst)),
      "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(m),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

// ---- cta_group::2 ("pair") forms
__device__ __forceinline__ uint32_t mapa_rank(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// TMA load whose complete_tx lands on a barrier given by its shared::cluster address (the pair leader's)
__device__ __forceinline__ void tma_load_4d_pair(void* dst, const CUtensorMap* m, uint32_t bar_caddr, int c0, int c1,
                                                 int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(m), "r"(bar_caddr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// halo loads of a layer whose residual operand is its own input (x + lrelu(conv(x)), CNNs.py:75-86): keep the lines in
// L2 (evict_last) so that the epilogue's residual load of the same pixels, a few microseconds later, is an L2 hit
// instead of a second trip to HBM (ncu: dram read = 2 x the input without the hint)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_4d_hint(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                                 int c3, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%3, %4, %5, %6}], [%2], %7;" ::"r"(smem_u32(dst)),
      "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair_hint(void* dst, const CUtensorMap* m, uint32_t bar_caddr, int c0,
                                                      int c1, int c2, int c3, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%3, %4, %5, %6}], [%2], %7;" ::"r"(smem_u32(dst)),
      "l"(m), "r"(bar_caddr), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void* dst, const CUtensorMap* m, uint32_t bar_caddr, int c0, int c1,
                                                 int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(m), "r"(bar_caddr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}
// the issue loop keeps descriptors as (lo, hi) words: only the 14-bit start-address field in the low word moves
template <bool kPair>
__device__ __forceinline__ void umma_bf16_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t accumulate) {
  if (kPair) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 da, db;\n"
        "setp.ne.b32 p, %6, 0;\n"
        "mov.b64 da, {%1, %2};\n"
        "mov.b64 db, {%3, %4};\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 da, db;\n"
        "setp.ne.b32 p, %6, 0;\n"
        "mov.b64 da, {%1, %2};\n"
        "mov.b64 db, {%3, %4};\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// arrive on the barrier at this offset in BOTH CTAs of the pair once all previously issued MMAs have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
// accumulator-drained signal to the pair leader.  Relaxed: it orders no memory, only TMEM reads that
// tcgen05.wait::ld + tcgen05.fence::before_thread_sync have already retired (a release at cluster scope costs a
// MEMBAR.ALL.GPU-class fence per arrive: 7 % of the samples in profiles/r1e_conv5fwd_pair).
__device__ __forceinline__ void mbar_arrive_caddr(uint32_t bar_caddr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_caddr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* slot_in_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// work item -> (pixel group, pass).  Pair mode: both CTAs of a pair run the same pass on neighbouring groups.
template <bool kPair>
__device__ __forceinline__ void v2_item(const V2P& p, int itn, uint32_t crank, int& grp, int& pass) {
  // the pass of an item is rotated by its group index: the passes of a stride-2 layer carry 1, 2, 2 and 4 taps, and
  // with an even number of CTAs (pairs) a plain j % npass would hand every CTA the same pass in every iteration
  // (half the machine running the 3-tap pass, half the 6-tap one)
  if (kPair) {
    const int j = (int)(blockIdx.x >> 1) + itn * (int)(gridDim.x >> 1);
    const int g2 = j / p.npass;
    pass = (j - g2 * p.npass + g2) % p.npass;
    grp = 2 * g2 + (int)crank;
  } else {
    const int wi = (int)blockIdx.x + itn * (int)gridDim.x;
    grp = wi / p.npass;
    pass = (wi - grp * p.npass + grp) % p.npass;
  }
}

// kPair instantiations are separate kernels: code holding cta_group::2 instructions only launches in even clusters.
// kF16: operands, skip / residual tensors and outputs are IEEE half instead of bf16 (forward of the "fp16" precision)
// kEpi: 0 = every epilogue feature decided at run time.  Non-zero = a staged epilogue whose features are fixed at compile
// time (bit mask below): no per-chunk flag tests, no spills (148-158 registers against 168 + spills), and the other
// epilogue paths drop out of the kernel -- 5-18 % per layer (profiles/r2t_epilogue_specialisation.txt).  The host picks
// a specialised instantiation only when every flag matches one in V2_EPI_LIST.
constexpr int EPI_SPEC = 1, EPI_ADD = 2, EPI_ADD1 = 4, EPI_LRELU = 8, EPI_MASKMUL = 16, EPI_PRE = 32, EPI_MASKOUT = 64,
              EPI_POOL = 128, EPI_POOLONLY = 256, EPI_OUT2 = 512, EPI_UP = 1024;   // EPI_UP: the register-path epilogue of the stride-2 'up' layers
// forward of a residual layer / of the first layer of a triple, training (sign mask) and inference, with the fused
// max-pool; input gradient with / without the second output, with / without the skip add, with / without the mask
#define V2_EPI_LIST(X)                                                                                      \
  X(EPI_SPEC | EPI_ADD | EPI_ADD1 | EPI_LRELU | EPI_MASKOUT)                                                \
  X(EPI_SPEC | EPI_ADD | EPI_ADD1 | EPI_LRELU | EPI_MASKOUT | EPI_POOL)                                     \
  X(EPI_SPEC | EPI_LRELU | EPI_MASKOUT)                                                                     \
  X(EPI_SPEC | EPI_ADD | EPI_ADD1 | EPI_LRELU)                                                              \
  X(EPI_SPEC | EPI_ADD | EPI_ADD1 | EPI_LRELU | EPI_POOL | EPI_POOLONLY)                                    \
  X(EPI_SPEC | EPI_LRELU)                                                                                   \
  X(EPI_SPEC | EPI_ADD | EPI_MASKMUL | EPI_PRE)                                                             \
  X(EPI_SPEC | EPI_ADD | EPI_MASKMUL)                                                                       \
  X(EPI_SPEC | EPI_MASKMUL | EPI_PRE)                                                                       \
  X(EPI_SPEC)                                                                                               \
  X(EPI_SPEC | EPI_UP | EPI_LRELU | EPI_MASKOUT)                                                            \
  X(EPI_SPEC | EPI_UP | EPI_LRELU)
// the "fp16" precision's forward kernels (IEEE-half operands; the training forms also write the bf16 twin)
#define V2_EPI_LIST_F16(X)                                                                                  \
  X(EPI_SPEC | EPI_ADD | EPI_ADD1 | EPI_LRELU | EPI_MASKOUT | EPI_OUT2)                                     \
  X(EPI_SPEC | EPI_LRELU | EPI_MASKOUT | EPI_OUT2)                                                          \
  X(EPI_SPEC | EPI_ADD | EPI_ADD1 | EPI_LRELU)                                                              \
  X(EPI_SPEC | EPI_ADD | EPI_ADD1 | EPI_LRELU | EPI_POOL | EPI_POOLONLY)                                    \
  X(EPI_SPEC | EPI_LRELU)
template <bool kPair, bool kF16, int kEpi = 0>
__global__ void __launch_bounds__(V2_THREADS, 1)
tc_conv2_kernel(const __grid_constant__ V2Maps maps, const V2P p) {
  const int dbg = kEpi ? 0 : p.debug;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[V2_MAX_A_STAGES];
  __shared__ __align__(8) uint64_t a_empty[V2_MAX_A_STAGES];
  __shared__ __align__(8) uint64_t b_full[V2_MAX_B_STAGES];
  __shared__ __align__(8) uint64_t b_empty[V2_MAX_B_STAGES];
  __shared__ __align__(8) uint64_t bres_full;
  __shared__ __align__(8) uint64_t e_full[V2_MAX_E_STAGES];
  __shared__ __align__(8) uint64_t e_empty[V2_MAX_E_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float sbias[256];
  // per-tap issue table (A descriptor of stage 0 / tile 0, accumulator column): the MMA issuer's inner loop is then
  // two shared-memory loads and a handful of 32-bit adds per tap -- it must stay below 48 cycles per N=64 MMA
  // {A descriptor lo, hi, accumulator column | accumulator << 16, unused}: ONE 16-byte shared load per tap, issued one
  // tap ahead of its MMAs from an address computed once (the compiler otherwise re-derives the shared-window address
  // -- an S2R of the cluster CTA id -- and waits for the load inside every tap iteration: the serial chain S2R -> LDS ->
  // R2UR -> UTCHMMA paced the T = 1 (stride-2) layers at 2.5x their operand-fetch floor)
  __shared__ __align__(16) uint4 s_tap[PB_MAX_TAPS];

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int c = threadIdx.x; c < 256; c += V2_THREADS)
    sbias[c] = (p.bias != nullptr && c < p.Cout) ? __ldg(p.bias + c) : 0.f;
  if ((int)threadIdx.x < p.ntaps) {
    const uint64_t d = smem_desc_sw128_any(smem_u32(smem) + p.taps[threadIdx.x].a_off, p.taps[threadIdx.x].sbo,
                                           p.use_base_offset);
    s_tap[threadIdx.x] = make_uint4((uint32_t)d, (uint32_t)(d >> 32),
                                    (uint32_t)(p.taps[threadIdx.x].acc * p.n_tile) | ((uint32_t)p.taps[threadIdx.x].acc << 16), 0u);
  }
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.nboxes; ++i) prefetch_tmap(&maps.a[i]);
    prefetch_tmap(&maps.b);
    for (int s = 0; s < V2_MAX_A_STAGES; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < V2_MAX_B_STAGES; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], kPair ? 1u : (uint32_t)p.cluster);   // one tcgen05.commit per CTA that reads the multicast tile
    }
    mbar_init(&bres_full, 1);
    for (int s = 0; s < V2_MAX_E_STAGES; ++s) {
      mbar_init(&e_full[s], 1);
      mbar_init(&e_empty[s], 4);   // lane 0 of the four group-A epilogue warps, after the pair barrier
    }
    if (p.e_has_add) prefetch_tmap(&maps.e);
    if (p.e_mode && !p.pool_only) prefetch_tmap(&maps.o);
    if (p.pool) prefetch_tmap(&maps.pl);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      // staged epilogue: both epilogue warp groups; pair mode: the leader's barrier collects both CTAs' warps
      mbar_init(&tmem_empty_bar[s], 8u * (kPair ? 2u : 1u));   // all eight epilogue warps, in every mode
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (kPair) tmem_alloc_pair(&tmem_slot, 512);
    else tmem_alloc(&tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  if (p.cluster > 1) cluster_sync_all();   // peers' barriers are initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t crank = p.cluster > 1 ? cluster_ctarank() : 0u;
  const uint16_t cmask = (uint16_t)((1u << p.cluster) - 1u);

  const int total_groups = p.N * p.groups_h * p.groups_w;
  const int total_items = total_groups * p.npass;      // work item = (pixel group, pass)
  // clusters run a uniform number of iterations (their weight stream is shared); lone CTAs stop at their last item
  const int iters = p.cluster > 1 ? p.iters : (total_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int cols_per_tile = p.n_acc * p.n_tile;
  const bool leader = !kPair || crank == 0;

  if (warp == 0) {
    if (elect_one()) {
      // ------------------------------------------------------------------ halo producer
      int stage = 0;
      uint32_t phase = 0;
      const uint64_t pol = l2_policy_evict_last();
      for (int itn = 0; itn < iters; ++itn) {
        int r, pass_unused;
        v2_item<kPair>(p, itn, crank, r, pass_unused);
        const int gw = r % p.groups_w; r /= p.groups_w;
        const int gh = r % p.groups_h;
        const int img = r / p.groups_h;   // >= N for the padding iterations of a cluster: TMA zero-fills
        const int h0 = gh * V2_TILE_H, w0 = gw * V2_TILE_W * p.T;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          if (dbg & 4) break;
          mbar_wait_relaxed(&a_empty[stage], phase ^ 1, p.poll_ns);
          uint8_t* sa = smem + (size_t)stage * p.a_stage_bytes;
          if (kPair) {
            // the leader's barrier counts both CTAs' halos (its own producer arms it for 2x the bytes)
            if (crank == 0) mbar_expect_tx(&a_full[stage], 2u * p.a_tx_bytes);
            const uint32_t bar = mapa_rank(smem_u32(&a_full[stage]), 0);
            for (int b = 0; b < p.nboxes; ++b) {
              if (p.a_keep_l2)
                tma_load_4d_pair_hint(sa + p.boxes[b].smem_off, &maps.a[b], bar, kc * 64, w0 + p.boxes[b].dx0,
                                      h0 + p.boxes[b].dy0, img, pol);
              else
                tma_load_4d_pair(sa + p.boxes[b].smem_off, &maps.a[b], bar, kc * 64, w0 + p.boxes[b].dx0,
                                 h0 + p.boxes[b].dy0, img);
            }
          } else {
            mbar_expect_tx(&a_full[stage], p.a_tx_bytes);
            for (int b = 0; b < p.nboxes; ++b) {
              if (p.a_keep_l2)
                tma_load_4d_hint(sa + p.boxes[b].smem_off, &maps.a[b], &a_full[stage], kc * 64, w0 + p.boxes[b].dx0,
                                 h0 + p.boxes[b].dy0, img, pol);
              else
                tma_load_4d(sa + p.boxes[b].smem_off, &maps.a[b], &a_full[stage], kc * 64, w0 + p.boxes[b].dx0,
                            h0 + p.boxes[b].dy0, img);
            }
          }
          if (++stage == p.a_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 6) {
    if (elect_one()) {
      // ------------------------------------------------------------------ weight producer
      uint8_t* sb = smem + p.b_ring_off;
      if (p.b_resident && kPair) {
        if (crank == 0) mbar_expect_tx(&bres_full, 2u * (uint32_t)(p.ntaps * p.kchunks) * p.b_bytes);
        const uint32_t bar = mapa_rank(smem_u32(&bres_full), 0);
        for (int t = 0; t < p.ntaps; ++t)
          for (int kc = 0; kc < p.kchunks; ++kc)
            tma_load_3d_pair(sb + (size_t)(t * p.kchunks + kc) * p.b_bytes, &maps.b, bar, kc * 64,
                             (int)crank * (p.n_tile >> 1), p.taps[t].wtap);
      } else if (p.b_resident) {
        mbar_expect_tx(&bres_full, (uint32_t)(p.ntaps * p.kchunks) * p.b_bytes);
        for (int t = 0; t < p.ntaps; ++t)
          for (int kc = 0; kc < p.kchunks; ++kc)
            tma_load_3d(sb + (size_t)(t * p.kchunks + kc) * p.b_bytes, &maps.b, &bres_full, kc * 64, 0, p.taps[t].wtap);
      } else if (kPair) {
        // each CTA streams ITS half (rows crank*N/2 ...) of every weight tile into its own ring; the leader's
        // barrier counts both halves
        int stage = 0;
        uint32_t phase = 0;
        const int rows = p.n_tile >> 1;
        for (int itn = 0; itn < iters; ++itn) {
          int grp_unused, pass;
          v2_item<kPair>(p, itn, crank, grp_unused, pass);
          for (int kc = 0; kc < p.kchunks; ++kc) {
            for (int t = p.pass_begin[pass]; t < p.pass_begin[pass + 1]; ++t) {
              if (dbg & 2) break;
              mbar_wait(&b_empty[stage], phase ^ 1);
              if (crank == 0) mbar_expect_tx(&b_full[stage], 2u * p.b_bytes);
              tma_load_3d_pair(sb + (size_t)stage * p.b_bytes, &maps.b, mapa_rank(smem_u32(&b_full[stage]), 0), kc * 64,
                               (int)crank * rows, p.taps[t].wtap);
              if (++stage == p.b_stages) { stage = 0; phase ^= 1; }
            }
          }
        }
      } else {
        int stage = 0;
        uint32_t phase = 0;
        const int rows = p.n_tile / p.cluster;            // weight rows this CTA fetches of every tile
        const uint32_t slice_off = crank * (uint32_t)rows * 128u;
        for (int itn = 0; itn < iters; ++itn) {
          int grp_unused, pass;
          v2_item<kPair>(p, itn, crank, grp_unused, pass);
          for (int kc = 0; kc < p.kchunks; ++kc) {
            for (int t = p.pass_begin[pass]; t < p.pass_begin[pass + 1]; ++t) {
              if (dbg & 2) break;
              mbar_wait(&b_empty[stage], phase ^ 1);      // every CTA of the cluster has released this stage
              mbar_expect_tx(&b_full[stage], p.b_bytes);
              uint8_t* dst = sb + (size_t)stage * p.b_bytes + slice_off;
              const int wt = p.taps[t].wtap;
              if (p.cluster > 1) tma_load_3d_mc(dst, &maps.b, &b_full[stage], kc * 64, (int)crank * rows, wt, cmask);
              else tma_load_3d(dst, &maps.b, &b_full[stage], kc * 64, 0, wt);
              if (++stage == p.b_stages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 7) {
    if (p.e_has_add && !(dbg & 1) && elect_one()) {
      // ------------------------------------------------------------------ epilogue-operand producer
      uint8_t* se = smem + p.e_ring_off;
      const int c64n = p.n_tile >> 6;
      int stage = 0;
      uint32_t phase = 0;
      for (int itn = 0, grp = blockIdx.x; itn < iters; ++itn, grp += gridDim.x) {
        int r = grp;
        const int gw = r % p.groups_w; r /= p.groups_w;
        const int gh = r % p.groups_h;
        const int img = r / p.groups_h;
        const int h0 = gh * V2_TILE_H;
        for (int tile = 0; tile < p.T; ++tile) {
          const int w0 = (gw * p.T + tile) * V2_TILE_W;
          for (int c64 = 0; c64 < c64n; ++c64) {
            mbar_wait_relaxed(&e_empty[stage], phase ^ 1, p.poll_ns);
            mbar_expect_tx(&e_full[stage], V2_E_BYTES);
            tma_load_4d(se + (size_t)stage * V2_E_BYTES, &maps.e, &e_full[stage], c64 * 64, w0, h0, img);
            if (++stage == p.e_stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      // ------------------------------------------------------------------ MMA issuer
      const uint32_t idesc = make_idesc(kPair ? 256 : 128, p.n_tile, 0, 0, kF16, kF16);
      const uint32_t b_ring = smem_u32(smem + p.b_ring_off);
      // weight-tile descriptor of ring slot 0 (K-major, SWIZZLE_128B, 8-row groups 1024 B apart); slots / resident
      // tiles are b_bytes apart
      const uint64_t bd0 = smem_desc_sw128(b_ring, 16, 1024);
      const uint32_t bd_lo0 = (uint32_t)bd0, bd_hi = (uint32_t)(bd0 >> 32);
      const uint32_t tap_tab = smem_u32(s_tap);
      // register copy of the tap table for the unrolled issue path (3x3 layers: at most nine taps)
      const bool unrolled = p.ntaps <= V2_UNROLL_TAPS && p.unroll_taps != 0;
      uint32_t ta_lo[V2_UNROLL_TAPS], ta_hi[V2_UNROLL_TAPS], tcol[V2_UNROLL_TAPS];
#pragma unroll
      for (int u = 0; u < V2_UNROLL_TAPS; ++u) {
        const uint4 e = ld_shared_v4(tap_tab + 16u * (uint32_t)(u < p.ntaps ? u : 0));
        ta_lo[u] = e.x; ta_hi[u] = e.y; tcol[u] = e.z;
      }
      const uint32_t b_step16 = p.b_bytes >> 4;
      const uint32_t bres_step16 = ((uint32_t)p.kchunks * p.b_bytes) >> 4;
      int astage = 0, bstage = 0;
      uint32_t aphase_s = 0, bphase_s = 0;
      if (p.b_resident) {
        mbar_wait(&bres_full, 0);
        tc_fence_after();
      }
      int it = 0;
      for (; it < iters; ++it) {
        int grp_unused, pass;
        v2_item<kPair>(p, it, crank, grp_unused, pass);
        const int as = it % p.acc_stages;
        const uint32_t accphase = (uint32_t)(it / p.acc_stages) & 1u;
        mbar_wait(&tmem_empty_bar[as], accphase ^ 1);
        tc_fence_after();
        uint32_t started = 0;
        const uint32_t d_stage = tmem_base + (uint32_t)(as * p.T * cols_per_tile);
        const int t_begin = p.pass_begin[pass], t_end = p.pass_begin[pass + 1];
        if (unrolled) {
          // ---- taps unrolled at compile time over the register-resident table: one basic block per K chunk, so the
          //      R2UR / uniform-add chains of later taps are scheduled behind the UTCHMMAs of earlier ones instead of
          //      serialising with them (a lone thread issues dependent instructions ~5 cycles apart: the rolled loop's
          //      ~50 instructions per tap paced a T = 1 layer at 110 cycles per MMA against a 44-cycle floor)
          for (int kc = 0; kc < p.kchunks; ++kc) {
            if (!(dbg & 4)) mbar_wait(&a_full[astage], aphase_s);
            tc_fence_after();
            const uint32_t a_off16 = ((uint32_t)astage * p.a_stage_bytes) >> 4;
            uint32_t bres16 = bd_lo0 + (((uint32_t)(t_begin * p.kchunks + kc) * p.b_bytes) >> 4);
#pragma unroll
            for (int u = 0; u < V2_UNROLL_TAPS; ++u) {
              if (u >= t_begin && u < t_end) {
                uint32_t b_lo;
                if (p.b_resident) {
                  b_lo = bres16;
                  bres16 += bres_step16;
                } else {
                  if (!(dbg & 2)) mbar_wait(&b_full[bstage], bphase_s);
                  tc_fence_after();
                  b_lo = bd_lo0 + (uint32_t)bstage * b_step16;
                }
                const uint32_t accbit = 1u << (tcol[u] >> 16);
                const uint32_t first = (started & accbit) ? 1u : 0u;
                uint32_t a_lo = ta_lo[u] + a_off16;
                uint32_t d_tmem = d_stage + (tcol[u] & 0xFFFFu);
#pragma unroll 1
                for (int tile = 0; tile < p.T; ++tile) {
                  umma_bf16_lohi<kPair>(d_tmem, a_lo, ta_hi[u], b_lo, bd_hi, idesc, first);
                  umma_bf16_lohi<kPair>(d_tmem, a_lo + 2, ta_hi[u], b_lo + 2, bd_hi, idesc, 1u);
                  umma_bf16_lohi<kPair>(d_tmem, a_lo + 4, ta_hi[u], b_lo + 4, bd_hi, idesc, 1u);
                  umma_bf16_lohi<kPair>(d_tmem, a_lo + 6, ta_hi[u], b_lo + 6, bd_hi, idesc, 1u);
                  a_lo += 64;
                  d_tmem += (uint32_t)cols_per_tile;
                }
                started |= accbit;
                if (!p.b_resident && !(dbg & 2)) {
                  if (kPair) umma_commit_pair(&b_empty[bstage]);
                  else if (p.cluster > 1) umma_commit_mc(&b_empty[bstage], cmask);
                  else umma_commit(&b_empty[bstage]);
                  if (++bstage == p.b_stages) { bstage = 0; bphase_s ^= 1; }
                }
              }
            }
            if (!(dbg & 4)) {
              if (kPair) umma_commit_pair(&a_empty[astage]);
              else umma_commit(&a_empty[astage]);
            }
            if (++astage == p.a_stages) { astage = 0; aphase_s ^= 1; }
          }
        } else {
        uint4 tap_next = ld_shared_v4(tap_tab + 16u * (uint32_t)t_begin);
        for (int kc = 0; kc < p.kchunks; ++kc) {
          if (!(dbg & 4)) mbar_wait(&a_full[astage], aphase_s);
          tc_fence_after();
          const uint32_t a_off16 = ((uint32_t)astage * p.a_stage_bytes) >> 4;
          uint32_t bres16 = bd_lo0 + (((uint32_t)(t_begin * p.kchunks + kc) * p.b_bytes) >> 4);
          for (int t = t_begin; t < t_end; ++t) {
            const uint2 ad = make_uint2(tap_next.x, tap_next.y);
            const uint32_t colw = tap_next.z;
            tap_next = ld_shared_v4(tap_tab + 16u * (uint32_t)(t + 1 < t_end ? t + 1 : t_begin));   // in flight behind this tap's MMAs
            uint32_t b_lo;
            if (p.b_resident) {
              b_lo = bres16;
              bres16 += bres_step16;
            } else {
              if (!(dbg & 2)) mbar_wait(&b_full[bstage], bphase_s);
              tc_fence_after();
              b_lo = bd_lo0 + (uint32_t)bstage * b_step16;
            }
            const uint32_t accbit = 1u << (colw >> 16);
            const uint32_t first = (started & accbit) ? 1u : 0u;
            uint32_t a_lo = ad.x + a_off16;
            uint32_t d_tmem = d_stage + (colw & 0xFFFFu);
#pragma unroll 1
            for (int tile = 0; tile < p.T; ++tile) {
              umma_bf16_lohi<kPair>(d_tmem, a_lo, ad.y, b_lo, bd_hi, idesc, first);
              umma_bf16_lohi<kPair>(d_tmem, a_lo + 2, ad.y, b_lo + 2, bd_hi, idesc, 1u);
              umma_bf16_lohi<kPair>(d_tmem, a_lo + 4, ad.y, b_lo + 4, bd_hi, idesc, 1u);
              umma_bf16_lohi<kPair>(d_tmem, a_lo + 6, ad.y, b_lo + 6, bd_hi, idesc, 1u);
              a_lo += 64;                       // next tile: 8 pixel columns = 1024 B further in the halo
              d_tmem += (uint32_t)cols_per_tile;
            }
            started |= accbit;
            if (!p.b_resident && !(dbg & 2)) {
              if (kPair) umma_commit_pair(&b_empty[bstage]);
              else if (p.cluster > 1) umma_commit_mc(&b_empty[bstage], cmask);
              else umma_commit(&b_empty[bstage]);
              if (++bstage == p.b_stages) { bstage = 0; bphase_s ^= 1; }
            }
          }
          if (!(dbg & 4)) {
            if (kPair) umma_commit_pair(&a_empty[astage]);
            else umma_commit(&a_empty[astage]);
          }
          if (++astage == p.a_stages) { astage = 0; aphase_s ^= 1; }
        }
        }
        if (kPair) umma_commit_pair(&tmem_full_bar[as]);
        else umma_commit(&tmem_full_bar[as]);
      }
    }
  } else {
    // -------------------------------------------------------------------- epilogue warps 2..5 (+ 8..11 when staged)
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int egrp = warp >= 8 ? 1 : 0;   // staged path: group A takes channels 0-31 of every 64-block, group B 32-63
    const int ml = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const int nch = (p.n_tile + 31) >> 5;            // 32-channel chunks per accumulator (last may be 16 wide)
    const int per_tile = p.n_acc * nch;
    const int chunks = p.T * per_tile;
    int it = 0;
    int estage = 0, mbuf = 0, prev_stage = -1, pslot = 0;
    uint32_t ephase = 0;
    float head_acc = 0.f;   // head_mode 2: this thread's share of the squared-error sum
    // accumulator-drained signal: the MMA issuer's (the pair leader's) barrier
    const uint32_t tmem_empty_caddr0 = kPair ? mapa_rank(smem_u32(&tmem_empty_bar[0]), 0) : smem_u32(&tmem_empty_bar[0]);
    for (; it < iters; ++it) {
      const int as = it % p.acc_stages;
      const uint32_t accphase = (uint32_t)(it / p.acc_stages) & 1u;
      int grp, pass;
      v2_item<kPair>(p, it, crank, grp, pass);
      int r = grp;
      const int gw = r % p.groups_w; r /= p.groups_w;
      const int gh = r % p.groups_h;
      const int img = r / p.groups_h;
      const int bh = gh * V2_TILE_H + (ml >> 3);
      // timing experiments: 1 = no epilogue work at all; 256 / 512 = none in the TMEM quadrant whose warps share / do not
      // share a scheduler with the MMA-issuing warp (warp 1 -> quadrant 1; 512 idles quadrant 2 instead)
      if ((dbg & 1) || ((dbg & 256) && q == 1) || ((dbg & 512) && q == 2)) {
        mbar_wait_relaxed(&tmem_full_bar[as], accphase, p.poll_ns);
        tc_fence_after();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (kPair) mbar_arrive_caddr(tmem_empty_caddr0 + 8u * (uint32_t)as);
          else mbar_arrive(&tmem_empty_bar[as]);
        }
        continue;
      }
      if (kEpi == 0 && p.out_nchw) {
        mbar_wait_relaxed(&tmem_full_bar[as], accphase, p.poll_ns);
        tc_fence_after();
        // network head: NCHW fp32 heatmaps, bias + activation only.  Both epilogue warp groups work: the
        // (output-row parity, 16-channel chunk) units of a tile alternate between them
        float* outf = reinterpret_cast<float*>(p.out);
        const int nph = p.up ? (p.n_acc >> 1) : 1;   // output-row parities covered by this pass
        const int nchunk = p.n_tile >> 4;
        const long long plane = (long long)p.OH * p.OW;
        for (int tile = 0; tile < p.T; ++tile) {
          const int bw = (gw * p.T + tile) * V2_TILE_W + (ml & 7);
          const bool ok = bh < p.BH && bw < p.BW && img < p.N;
          const uint32_t tile_col = (uint32_t)((as * p.T + tile) * p.n_acc * p.n_tile);
          for (int u = egrp; u < nph * nchunk; u += 2) {
            const int py = u / nchunk, c0 = (u - py * nchunk) << 4;
            const int oy = p.up ? 2 * bh + pass * nph + py : bh;
            const int ox0 = p.up ? 2 * bw : bw;
            uint32_t r0[16], r1[16];
            const int a0 = p.up ? py * 2 : 0;
            tmem_ld16(lane_base + tile_col + (uint32_t)(a0 * p.n_tile + c0), r0);
            if (p.up) tmem_ld16(lane_base + tile_col + (uint32_t)((a0 + 1) * p.n_tile + c0), r1);
            tmem_ld_wait();
            if (p.head_mode == 1) {
              // ---- arg-max of every (image, channel) map without the map: this thread's two pixels, then the warp's
              //      32 x 2 pixels (redux.sync max on the order key, redux.sync min on the index among the holders:
              //      lowest flat index on ties, NaN greatest -- Augmentor.py:131), then ONE atomicMax per channel
              //      from lane j of the warp for channel c0 + j
              const uint32_t idx0 = (uint32_t)(oy * p.OW + ox0);
              uint32_t k_hi = 0u, k_lo = 0u;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                if (c0 + j < p.Cout) {
                  const float b = sbias[c0 + j];
                  float v0 = __uint_as_float(r0[j]) + b, v1 = __uint_as_float(r1[j]) + b;
                  if (p.act == PB_ACT_LRELU) {
                    v0 = v0 > 0.f ? v0 : p.slope * v0;
                    v1 = v1 > 0.f ? v1 : p.slope * v1;
                  }
                  const uint32_t k0 = order_key(v0), k1 = p.up ? order_key(v1) : 0u;
                  const uint32_t kb = ok ? (k1 > k0 ? k1 : k0) : 0u;
                  const uint32_t ib = k1 > k0 ? idx0 + 1u : idx0;
                  const uint32_t m = __reduce_max_sync(0xffffffffu, kb);
                  const uint32_t mi = __reduce_min_sync(0xffffffffu, (ok && kb == m) ? ib : 0xFFFFFFFFu);
                  if (lane == j) { k_hi = m; k_lo = 0xFFFFFFFFu - mi; }
                }
              }
              if (lane < 16 && c0 + lane < p.Cout && img < p.N && k_hi != 0u)
                atomicMax(p.head_keys + (long long)img * p.Cout + c0 + lane, ((unsigned long long)k_hi << 32) | k_lo);
            } else if (p.head_mode == 2) {
              // ---- MSE against the target (read, or rendered: exp(-r^2 / 2 sigma^2) as one ex2.approx, exactly as
              //      mse_nhwc_bf16_kernel does) and the gradient w.r.t. this layer's pre-activation, bf16 NHWC
              float g0[16], g1[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                g0[j] = g1[j] = 0.f;
                if (c0 + j < p.Cout && ok) {
                  const float b = sbias[c0 + j];
                  float v0 = __uint_as_float(r0[j]) + b, v1 = __uint_as_float(r1[j]) + b;
                  if (p.act == PB_ACT_LRELU) {
                    v0 = v0 > 0.f ? v0 : p.slope * v0;
                    v1 = v1 > 0.f ? v1 : p.slope * v1;
                  }
                  float t0, t1;
                  if (p.head_target != nullptr) {
                    const float2 tt = __ldcs(reinterpret_cast<const float2*>(
                        p.head_target + ((long long)img * p.Cout + c0 + j) * plane + (long long)oy * p.OW + ox0));
                    t0 = tt.x; t1 = tt.y;
                  } else {
                    const float2 mxy = __ldg(reinterpret_cast<const float2*>(p.head_points) + ((long long)img * p.Cout + c0 + j));
                    const float dx0 = (float)ox0 - mxy.x, dx1 = (float)(ox0 + 1) - mxy.x, dy = (float)oy - mxy.y;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"((dx0 * dx0 + dy * dy) * p.head_negk2));
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"((dx1 * dx1 + dy * dy) * p.head_negk2));
                  }
                  const float d0 = v0 - t0, d1 = v1 - t1;
                  head_acc += d0 * d0 + d1 * d1;
                  g0[j] = d0 * (v0 > 0.f ? p.head_gscale : p.head_gscale * p.slope);
                  g1[j] = d1 * (v1 > 0.f ? p.head_gscale : p.head_gscale * p.slope);
                }
              }
              if (ok) {
                __nv_bfloat16* gd = p.head_grad + ((long long)img * plane + (long long)oy * p.OW + ox0) * p.head_cpad + c0;
                st_global_256(gd, pack16x8<false>(g0), pack16x8<false>(g0 + 8));
                st_global_256(gd + p.head_cpad, pack16x8<false>(g1), pack16x8<false>(g1 + 8));
              }
            } else if (ok) {
              float* dst = outf + ((long long)img * p.Cout + c0) * plane + (long long)oy * p.OW + ox0;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                if (c0 + j < p.Cout) {
                  const float b = sbias[c0 + j];
                  float v0 = __uint_as_float(r0[j]) + b, v1 = __uint_as_float(r1[j]) + b;
                  if (p.act == PB_ACT_LRELU) {
                    v0 = v0 > 0.f ? v0 : p.slope * v0;
                    v1 = v1 > 0.f ? v1 : p.slope * v1;
                  }
                  if (p.up) *reinterpret_cast<float2*>(dst) = make_float2(v0, v1);
                  else *dst = v0;
                }
                dst += plane;
              }
            }
          }
        }
      } else if (kEpi != 0 ? (kEpi & EPI_UP) == 0 : p.e_mode != 0) {
        // ---- staged epilogue (plain geometry, Cout % 64 == 0): skip/residual tile arrives by TMA,
        //      LeakyReLU' mask words by cp.async one tile ahead; no global-load latency on this path
        const int words = p.Cout >> 5;
        const bool masks = kEpi ? (kEpi & EPI_MASKMUL) != 0 : p.act == PB_ACT_MASKMUL;
        const bool has_add = kEpi ? (kEpi & EPI_ADD) != 0 : p.e_has_add != 0;
        const bool is_add1 = kEpi ? (kEpi & EPI_ADD1) != 0 : p.e_is_add1 != 0;
        const bool act_lrelu = kEpi ? (kEpi & EPI_LRELU) != 0 : p.act == PB_ACT_LRELU;
        const bool act_gelu = kEpi ? false : p.act == PB_ACT_GELU;
        const bool has_pre = kEpi ? (kEpi & EPI_PRE) != 0 : p.pre_out != nullptr;
        const bool has_mask_out = kEpi ? (kEpi & EPI_MASKOUT) != 0 : p.mask_out != nullptr;
        const bool has_out2 = kEpi ? (kF16 && (kEpi & EPI_OUT2) != 0) : (kF16 && p.out2 != nullptr);
        const bool do_pool = kEpi ? (kEpi & EPI_POOL) != 0 : p.pool != 0;
        const bool pool_only = kEpi ? (kEpi & EPI_POOLONLY) != 0 : p.pool_only != 0;
        uint32_t* smask = reinterpret_cast<uint32_t*>(smem + p.smask_off);
        auto mask_issue = [&](int g, int tile, int buf) {
          int rr = g;
          const int gw2 = rr % p.groups_w; rr /= p.groups_w;
          const int gh2 = rr % p.groups_h;
          const int img2 = rr / p.groups_h;
          const int bh2 = gh2 * V2_TILE_H + (ml >> 3);
          const int bw2 = (gw2 * p.T + tile) * V2_TILE_W + (ml & 7);
          if (bh2 < p.BH && bw2 < p.BW && img2 < p.N) {
            const uint32_t* src = p.mask_in + (((long long)img2 * p.OH + bh2) * p.OW + bw2) * words;
            for (int w = 0; w < words; ++w) {
              const uint32_t dst = smem_u32(smask + (buf * 8 + w) * 128 + ml);
              asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src + w) : "memory");
            }
          }
        };
        if (it == 0) {
          if (masks) mask_issue(grp, 0, 0);
          asm volatile("cp.async.commit_group;" ::: "memory");
        }
        uint8_t* se = smem + p.e_ring_off;
        for (int tile = 0; tile < p.T; ++tile) {
          if (masks) {
            if (tile + 1 < p.T) mask_issue(grp, tile + 1, mbuf ^ 1);
            else if (it + 1 < iters) mask_issue(grp + gridDim.x, 0, mbuf ^ 1);
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
          asm volatile("cp.async.wait_group 1;" ::: "memory");
          if (tile == 0) {
            mbar_wait_relaxed(&tmem_full_bar[as], accphase, p.poll_ns);
            tc_fence_after();
          }
          const int bw = (gw * p.T + tile) * V2_TILE_W + (ml & 7);
          const bool ok = bh < p.BH && bw < p.BW && img < p.N;
          const long long pix = ((long long)img * p.OH + bh) * p.OW + bw;
          const uint32_t tile_col = (uint32_t)((as * p.T + tile) * p.n_tile);
          for (int c64 = 0; c64 < (p.n_tile >> 6); ++c64) {
            if (has_add) mbar_wait_relaxed(&e_full[estage], ephase, p.poll_ns);
            uint8_t* erow = se + (size_t)estage * V2_E_BYTES + ml * 128;
            {
              const int half = egrp;
              const int c0 = c64 * 64 + half * 32;
              uint32_t rr[32];
              if (dbg & 16) {   // timing experiment: no TMEM read
#pragma unroll
                for (int j = 0; j < 32; ++j) rr[j] = 0x3f000000u + (uint32_t)j;
              } else {
                tmem_ld32(lane_base + tile_col + (uint32_t)c0, rr);
              }
              uint4 ev[4];
              if (has_add) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  ev[k] = *reinterpret_cast<const uint4*>(erow + (((half * 4 + k) ^ (ml & 7)) << 4));
              }
              tmem_ld_wait();
              float v[32];
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const float4 b = *reinterpret_cast<const float4*>(sbias + c0 + 4 * k);
                v[4 * k + 0] = __uint_as_float(rr[4 * k + 0]) + b.x;
                v[4 * k + 1] = __uint_as_float(rr[4 * k + 1]) + b.y;
                v[4 * k + 2] = __uint_as_float(rr[4 * k + 2]) + b.z;
                v[4 * k + 3] = __uint_as_float(rr[4 * k + 3]) + b.w;
              }
              const long long base = pix * p.Cout + c0;
              if (has_add && !is_add1) {
#pragma unroll
                for (int k = 0; k < 4; ++k) add16x8<kF16>(v + 8 * k, ev[k]);
              }
              if (has_pre && ok) {
                // second output (the unmasked gradient): 256-bit stores, one full 32-byte sector per instruction
                // (four 16-byte stores at a 128-byte lane stride fill every sector in two partial writes)
                if (dbg & 8) {
#pragma unroll
                  for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(p.pre_out + base + k * 8) = pack16x8<kF16>(v + 8 * k);
                } else {
#pragma unroll
                  for (int k = 0; k < 2; ++k) {
                    const uint4 lo = pack16x8<kF16>(v + 16 * k), hi = pack16x8<kF16>(v + 16 * k + 8);
                    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p.pre_out + base + k * 16),
                                 "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w)
                                 : "memory");
                  }
                }
              }
              if (act_lrelu) {
                // sign bits: (x > 0) is the sign of -x as an integer; the funnel shift appends it (2 ops / channel)
                uint32_t bits = 0;
#pragma unroll
                for (int j = 31; j >= 0; --j) {
                  bits = __funnelshift_l((uint32_t)(-(int)__float_as_uint(v[j])), bits, 1);
                  v[j] = fmaxf(v[j], p.slope * v[j]);      // LeakyReLU for 0 < slope < 1
                }
                if (has_mask_out && ok && !(dbg & 128)) p.mask_out[pix * words + (c0 >> 5)] = bits;
              } else if (masks) {
                const uint32_t m = smask[(mbuf * 8 + (c0 >> 5)) * 128 + ml];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] *= ((m >> j) & 1u) ? 1.f : p.slope;
              } else if (act_gelu) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0.5f * v[j] * (1.f + erff(v[j] * 0.70710678118654752440f));
              }
              if (has_add && is_add1) {
#pragma unroll
                for (int k = 0; k < 4; ++k) add16x8<kF16>(v + 8 * k, ev[k]);
              }
              if (has_out2 && ok) {
                // training in the "fp16" precision: the bf16 twin the weight gradient reads (see poseb200.h, PB_F16)
#pragma unroll
                for (int k = 0; k < 2; ++k)
                  st_global_256(p.out2 + base + k * 16, pack16x8<false>(v + 16 * k), pack16x8<false>(v + 16 * k + 8));
              }
              // the result row replaces the operand row this thread just consumed (same slot, same swizzle)
              if (!(dbg & 32)) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                *reinterpret_cast<uint4*>(erow + (((half * 4 + k) ^ (ml & 7)) << 4)) = pack16x8<kF16>(v + 8 * k);
              }
            }
            // both warps of this TMEM quadrant have written their halves: the quadrant's 32 pixels x 64 channels
            // leave as one coalesced TMA store
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
            uint8_t* pstage = nullptr;
            if (do_pool) {
              // 64 threads, 8 pooled pixels x 8 sixteen-byte channel chunks: one (pixel, chunk) each.  Source rows are
              // the quadrant's pixels (r, c) = rows 8r + c of its 4 KB slot, chunk j at position j ^ c (the swizzle key
              // of a row is its low three bits = c)
              const int t64 = egrp * 32 + lane;
              const int pp = t64 >> 3, j = t64 & 7;
              const int prow = pp >> 2, pcol = pp & 3;
              const uint8_t* qs = se + (size_t)estage * V2_E_BYTES + q * 4096;
              uint4 s4[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int r = 2 * prow + (k >> 1), c = 2 * pcol + (k & 1);
                s4[k] = *reinterpret_cast<const uint4*>(qs + (r * 8 + c) * 128 + ((j ^ c) << 4));
              }
              float m[8];
              {
                const uint32_t w0[4] = {s4[0].x, s4[0].y, s4[0].z, s4[0].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) { m[2 * e] = lo16<kF16>(w0[e]); m[2 * e + 1] = hi16<kF16>(w0[e]); }
              }
#pragma unroll
              for (int k = 1; k < 4; ++k) {
                const uint32_t wk[4] = {s4[k].x, s4[k].y, s4[k].z, s4[k].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  m[2 * e] = fmaxf(m[2 * e], lo16<kF16>(wk[e]));
                  m[2 * e + 1] = fmaxf(m[2 * e + 1], hi16<kF16>(wk[e]));
                }
              }
#pragma unroll
              for (int e = 0; e < 8; ++e) m[e] = m[e] > 0.f ? m[e] : p.slope * m[e];
              pstage = smem + p.pool_off + (size_t)((pslot * 4 + q) * 1024);
              *reinterpret_cast<uint4*>(pstage + pp * 128 + ((j ^ (pp & 7)) << 4)) = pack16x8<kF16>(m);
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
              asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
              if (++pslot == 3) pslot = 0;
            }
            if (egrp == 0) {
              if (elect_one()) {
                if (!pool_only && !(dbg & 64))
                  tma_store_4d(&maps.o, se + (size_t)estage * V2_E_BYTES + q * 4096, c64 * 64,
                               (gw * p.T + tile) * V2_TILE_W, gh * V2_TILE_H + 4 * q, img);
                if (do_pool)
                  tma_store_4d(&maps.pl, pstage, c64 * 64, (gw * p.T + tile) * (V2_TILE_W / 2), gh * (V2_TILE_H / 2) + 2 * q, img);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                // retire the PREVIOUS chunk's store (its shared-memory read), not this one: the store latency then
                // overlaps the next chunk's TMEM load and arithmetic.  With >= 3 slots the slot written next was
                // retired one chunk earlier, before this warp reached the pair barrier above, so both warps of the
                // quadrant see it free.
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                if (has_add && prev_stage >= 0) mbar_arrive(&e_empty[prev_stage]);
              }
            }
            __syncwarp();
            prev_stage = estage;
            if (++estage == p.e_stages) { estage = 0; ephase ^= 1; }
          }
          mbuf ^= 1;
        }
      } else {
        // the register-path epilogue's view of the parameters: a specialised 'up' instantiation is bias + LeakyReLU
        // (+ sign mask) only -- every other operand is a compile-time null
        EpiP ep = static_cast<const EpiP&>(p);
        if (kEpi & EPI_UP) {
          ep.add0 = nullptr; ep.add1 = nullptr; ep.pre_out = nullptr; ep.mask_in = nullptr; ep.out2 = nullptr;
          ep.act = PB_ACT_LRELU;
          if (!(kEpi & EPI_MASKOUT)) ep.mask_out = nullptr;
          else __builtin_assume(ep.mask_out != nullptr);
        }
        auto decode = [&](int idx, EpiPre& e) {
          const int tile = idx / per_tile;
          const int rem = idx - tile * per_tile;
          const int a = rem / nch;
          const int ci = rem - a * nch;
          e.c0 = ci * 32;
          e.width = (p.n_tile - e.c0) >= 32 ? 32 : 16;
          const int bw = (gw * p.T + tile) * V2_TILE_W + (ml & 7);
          e.ok = bh < p.BH && bw < p.BW && img < p.N;
          const int phase = pass * p.n_acc + a;
          const bool up = (kEpi & EPI_UP) ? true : p.up != 0;
          const int oy = up ? 2 * bh + (phase >> 1) : bh;
          const int ox = up ? 2 * bw + (phase & 1) : bw;
          e.pix = ((long long)img * p.OH + oy) * p.OW + ox;
          e.col = (uint32_t)((as * p.T + tile) * p.n_acc * p.n_tile + a * p.n_tile + e.c0);
          epi_prefetch(ep, e);
        };
        // (no operand prefetch here: this path serves the stride-2 transposed convs, whose epilogue is
        //  bias + LeakyReLU + sign mask only; layers with skip / residual operands take the staged path)
        mbar_wait_relaxed(&tmem_full_bar[as], accphase, p.poll_ns);
        tc_fence_after();
        for (int idx = egrp; idx < chunks; idx += 2) {   // the two warp groups alternate 32-channel chunks
          EpiPre cur;
          decode(idx, cur);
          if (cur.width == 32) {
            uint32_t rr[32];
            tmem_ld32(lane_base + cur.col, rr);
            tmem_ld_wait();
            if (cur.fast) {
              epi32_fast<kF16>(ep, sbias, rr, cur);
            } else {
              epilogue_chunk<32, kF16>(ep, rr, cur.pix, cur.c0, cur.ok);
              store_nhwc<32, kF16>(ep, rr, cur.pix, cur.c0, cur.ok);
            }
          } else {
            uint32_t rr[16];
            tmem_ld16(lane_base + cur.col, rr);
            tmem_ld_wait();
            epilogue_chunk<16, kF16>(ep, rr, cur.pix, cur.c0, cur.ok);
            store_nhwc<16, kF16>(ep, rr, cur.pix, cur.c0, cur.ok);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kPair) mbar_arrive_caddr(tmem_empty_caddr0 + 8u * (uint32_t)as);
        else mbar_arrive(&tmem_empty_bar[as]);
      }
    }
    if (p.head_mode == 2) {
      head_acc = warp_sum(head_acc);
      if (lane == 0) atomicAdd(p.head_loss, head_acc);
    }
  }
  // TMA stores landed (bulk groups are per thread: every lane waits, only the elected ones own groups)
  if (warp >= 2 && warp <= 5) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (p.cluster > 1) cluster_sync_all();   // no CTA exits while a peer may still signal its barriers
  if (warp == 1) {
    tc_fence_after();
    if (kPair) tmem_dealloc_pair(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

int conv_args_check(const pb_conv_args* a, const char* fn);

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v != nullptr && *v) ? atoi(v) : dflt;
}

static inline uint32_t round1k(uint32_t v) { return (v + 1023u) & ~1023u; }

// returns PB_OK and a filled (p, maps), or PB_ERR_UNSUPPORTED when the shape is outside this kernel
static int v2_plan(const pb_conv_args* a, V2P& p, V2Maps& maps, uint32_t budget, const V2Head* head) {
  const pb_taps& tp = a->taps;
  const bool plain = tp.out_mul == 1 && tp.in_div == 1;
  const bool up = tp.out_mul == 1 && tp.in_div == 2;
  const bool down = tp.out_mul == 2 && tp.in_div == 1;
  if (!(plain || up || down) || (a->Cin & 7) != 0) return PB_ERR_UNSUPPORTED;
  if (tp.ntaps < 2 && (a->OH < V2_TILE_H || a->Cout > 256)) return PB_ERR_UNSUPPORTED;  // 1 x rows "images" of nn.Linear
  memset((void*)&p, 0, sizeof(p));
  memset(&maps, 0, sizeof(maps));
  p.N = a->N;
  p.BH = up ? a->IH : a->OH;
  p.BW = up ? a->IW : a->OW;
  p.kchunks = cdiv(a->Cin, 64);
  p.ntaps = tp.ntaps;
  p.n_tile = cdiv(a->Cout, 16) * 16;
  if (p.n_tile > 256) return PB_ERR_UNSUPPORTED;
  // stride-2 transposed conv: as many passes over the pixel group as it takes for one pass's output phases to
  // double-buffer in TMEM (2 * n_acc * n_tile <= 512 columns)
  int npass = 1;
  if (up) {
    npass = (8 * p.n_tile <= 512) ? 1 : (4 * p.n_tile <= 512) ? 2 : 4;
    const int forced = env_int("POSEB200_CONV_NPASS", 0);
    if (forced == 1 || forced == 2 || forced == 4) npass = forced;
    if (a->out_nchw_f32 && npass == 4) return PB_ERR_UNSUPPORTED;   // the head stores x-parity pairs
  }
  p.npass = npass;
  p.n_acc = up ? 4 / npass : 1;
  if (p.n_acc * p.n_tile > 512) return PB_ERR_UNSUPPORTED;
  // cta_group::2 pairs for the wide layers (see V2P::pair); each CTA then holds N/2 rows of every weight tile
  p.pair = (p.n_tile >= env_int("POSEB200_CONV_PAIR_MIN_N", 48) && (p.n_tile % 16) == 0 && tp.ntaps >= 2 && env_int("POSEB200_CONV_PAIR", 1) != 0) ? 1 : 0;
  p.b_bytes = (uint32_t)(p.pair ? p.n_tile / 2 : p.n_tile) * 128u;
  const int cols_per_tile = p.n_acc * p.n_tile;
  const bool strips = env_int("POSEB200_CONV_PLAN_HALO", 1) == 0;  // default: one halo box per phase
  const bool cols8 = env_int("POSEB200_CONV_COLS8", 0) != 0;
  if (strips && down) return PB_ERR_UNSUPPORTED;

  // taps in pass order; per tap: shifted offsets, input phase (down), accumulator inside its pass (up)
  int tdy[PB_MAX_TAPS], tdx[PB_MAX_TAPS];
  {
    int n = 0;
    for (int pass = 0; pass < npass; ++pass) {
      p.pass_begin[pass] = n;
      for (int t = 0; t < tp.ntaps; ++t) {
        const int phase = up ? (tp.dy[t] & 1) * 2 + (tp.dx[t] & 1) : 0;
        if (phase / p.n_acc != pass) continue;
        tdy[n] = tp.dy[t]; tdx[n] = tp.dx[t];
        p.taps[n].wtap = t;
        p.taps[n].acc = phase % p.n_acc;
        ++n;
      }
    }
    for (int pass = npass; pass <= 4; ++pass) p.pass_begin[pass] = n;
  }
  int sdy[PB_MAX_TAPS], sdx[PB_MAX_TAPS], ph[PB_MAX_TAPS];
  for (int t = 0; t < tp.ntaps; ++t) {
    const int dy = tdy[t], dx = tdx[t];
    if (plain) {
      sdy[t] = dy; sdx[t] = dx; ph[t] = 0;
    } else if (up) {
      const int py = dy & 1, px = dx & 1;
      sdy[t] = (dy + py) / 2; sdx[t] = (dx + px) / 2; ph[t] = 0;
    } else {
      const int py = dy & 1, px = dx & 1;
      sdy[t] = (dy - py) / 2; sdx[t] = (dx - px) / 2; ph[t] = py * 2 + px;
    }
  }
  // staged epilogue (see kernel): reserve its shared memory before the operand rings are sized
  const bool staged = !up && !a->out_nchw_f32 && (a->Cout % 64) == 0 && !(a->add0 && a->add1) &&
                      !(a->act == PB_ACT_LRELU && !(a->slope > 0.f && a->slope < 1.f)) &&
                      env_int("POSEB200_TC_NO_STAGED_EPI", 0) == 0;
  const bool e_has_add = staged && (a->add0 || a->add1);
  const bool e_masks = staged && a->act == PB_ACT_MASKMUL;
  const uint32_t e_stages = (uint32_t)env_int("POSEB200_CONV_ESTAGES", 3);   // >= 3: see the deferred store retire in the kernel
  if (e_stages < 3 || e_stages > (uint32_t)V2_MAX_E_STAGES) return PB_ERR_INVALID;
  const bool pool = a->pool_out != nullptr;
  if (pool && !(staged && plain && (a->OH & 1) == 0 && (a->OW & 1) == 0 && !(a->act_dtype == PB_F16 && a->out2 != nullptr) &&
                a->slope > 0.f))
    return PB_ERR_UNSUPPORTED;
  const uint32_t epi_bytes = (staged ? e_stages * V2_E_BYTES : 0u) + (e_masks ? 8192u : 0u) + (pool ? 3u * 4096u : 0u);
  if (epi_bytes + 65536u > budget) return PB_ERR_UNSUPPORTED;
  budget -= epi_bytes;
  // default: as many tiles per group (up to 4) as still double-buffer their accumulators in TMEM; the placement loop
  // below steps down until halo ring + weights fit.  Round 2 (profiles/r2p_conv_T_sweep.txt): with N = 64 the per-group
  // hand-offs (accumulator full / drained, halo stage turn-over) are amortised over too few MMAs at T = 2 --
  // starting from T = 4 (in practice T = 3 with streamed weights next to the 48 KB epilogue ring) is 10-15 % faster on
  // the 64-channel forward layers and conv4's input gradient, neutral on the others.
  int T = env_int("POSEB200_TC_T", (8 * cols_per_tile <= 512) ? 4 : (4 * cols_per_tile <= 512) ? 2 : 1);
  if (T < 1) T = 1;
  if (T > 4) T = 4;
  while (T > 1 && (T * cols_per_tile > 512 || (T - 1) * V2_TILE_W >= p.BW)) --T;
  for (;; --T) {
    if (T < 1) return PB_ERR_UNSUPPORTED;
    p.T = T;
    const int tile_cols = V2_TILE_W * T;
    // ---- boxes
    p.nboxes = 0;
    uint32_t off = 0, tx = 0;
    int box_cols[V2_MAX_BOXES], box_rows[V2_MAX_BOXES], box_phase[V2_MAX_BOXES];
    bool fail = false;
    if (!strips) {
      // one halo box per input phase
      for (int phase = 0; phase < 4 && !fail; ++phase) {
        int ymin = 127, ymax = -127, xmin = 127, xmax = -127, cnt = 0;
        for (int t = 0; t < tp.ntaps; ++t)
          if (ph[t] == phase) {
            ++cnt;
            ymin = sdy[t] < ymin ? sdy[t] : ymin; ymax = sdy[t] > ymax ? sdy[t] : ymax;
            xmin = sdx[t] < xmin ? sdx[t] : xmin; xmax = sdx[t] > xmax ? sdx[t] : xmax;
          }
        if (cnt == 0) continue;
        const int b = p.nboxes++;
        box_cols[b] = tile_cols + (xmax - xmin);
        if (cols8) box_cols[b] = (box_cols[b] + 7) & ~7;  // keeps SBO a multiple of the 1024-byte swizzle period
        box_rows[b] = V2_TILE_H + (ymax - ymin);
        box_phase[b] = phase;
        p.boxes[b].dx0 = xmin; p.boxes[b].dy0 = ymin; p.boxes[b].smem_off = off;
        for (int t = 0; t < tp.ntaps; ++t)
          if (ph[t] == phase) {
            p.taps[t].a_off = off + (uint32_t)((sdy[t] - ymin) * box_cols[b] + (sdx[t] - xmin)) * 128u;
            p.taps[t].sbo = (uint32_t)box_cols[b] * 128u;
          }
        const uint32_t bytes = (uint32_t)box_cols[b] * box_rows[b] * 128u;
        tx += bytes;
        off += round1k(bytes);
      }
    } else {
      // one full-height strip per distinct x offset: every start stays 1024-byte aligned
      int ymin = 127, ymax = -127;
      for (int t = 0; t < tp.ntaps; ++t) {
        ymin = sdy[t] < ymin ? sdy[t] : ymin; ymax = sdy[t] > ymax ? sdy[t] : ymax;
      }
      for (int t = 0; t < tp.ntaps && !fail; ++t) {
        int b = -1;
        for (int i = 0; i < p.nboxes; ++i)
          if (p.boxes[i].dx0 == sdx[t]) b = i;
        if (b < 0) {
          if (p.nboxes == V2_MAX_BOXES) { fail = true; break; }
          b = p.nboxes++;
          box_cols[b] = tile_cols; box_rows[b] = V2_TILE_H + (ymax - ymin); box_phase[b] = 0;
          p.boxes[b].dx0 = sdx[t]; p.boxes[b].dy0 = ymin; p.boxes[b].smem_off = off;
          const uint32_t bytes = (uint32_t)box_cols[b] * box_rows[b] * 128u;
          tx += bytes;
          off += round1k(bytes);
        }
        p.taps[t].a_off = p.boxes[b].smem_off + (uint32_t)((sdy[t] - ymin) * box_cols[b]) * 128u;
        p.taps[t].sbo = (uint32_t)box_cols[b] * 128u;
      }
    }
    if (fail) return PB_ERR_UNSUPPORTED;
    p.a_stage_bytes = off;
    p.a_tx_bytes = tx;
    bool box_ok = true;
    for (int b = 0; b < p.nboxes; ++b)
      if (box_cols[b] > 256 || box_rows[b] > 256) box_ok = false;
    // ---- shared-memory split between the halo ring and the weights
    const uint32_t w_all = (uint32_t)(p.ntaps * p.kchunks) * p.b_bytes;
    bool placed = false;
    if (box_ok && w_all + 2u * p.a_stage_bytes <= budget && env_int("POSEB200_TC_NO_BRES", 0) == 0) {
      p.b_resident = 1;
      p.a_stages = (int)((budget - w_all) / p.a_stage_bytes);
      p.b_stages = 1;
      placed = true;
    } else if (box_ok && 2u * p.a_stage_bytes + 2u * p.b_bytes <= budget) {
      p.b_resident = 0;
      p.a_stages = 2;
      p.b_stages = (int)((budget - 2u * p.a_stage_bytes) / p.b_bytes);
      placed = true;
    }
    if (!placed) {
      if (T == 1) return PB_ERR_UNSUPPORTED;
      continue;
    }
    if (p.a_stages > V2_MAX_A_STAGES) p.a_stages = V2_MAX_A_STAGES;
    if (p.b_stages > V2_MAX_B_STAGES) p.b_stages = V2_MAX_B_STAGES;
    p.b_ring_off = (uint32_t)p.a_stages * p.a_stage_bytes;
    // ---- tensor maps of the boxes
    const uint64_t C = (uint64_t)a->Cin;
    for (int b = 0; b < p.nboxes; ++b) {
      const uint32_t box[4] = {64, (uint32_t)box_cols[b], (uint32_t)box_rows[b], 1};
      int rc;
      if (!down) {
        const uint64_t dims[4] = {C, (uint64_t)a->IW, (uint64_t)a->IH, (uint64_t)a->N};
        const uint64_t str[3] = {C * 2, (uint64_t)a->IW * C * 2, (uint64_t)a->IH * a->IW * C * 2};
        rc = encode_tmap_bf16(&maps.a[b], a->in, 4, dims, str, box);
      } else {
        const int py = box_phase[b] >> 1, px = box_phase[b] & 1;
        const __nv_bfloat16* base = (const __nv_bfloat16*)a->in + ((size_t)py * a->IW + px) * C;
        const uint64_t dims[4] = {C, (uint64_t)a->IW / 2, (uint64_t)a->IH / 2, (uint64_t)a->N};
        const uint64_t str[3] = {2 * C * 2, 2 * (uint64_t)a->IW * C * 2, (uint64_t)a->IH * a->IW * C * 2};
        rc = encode_tmap_bf16(&maps.a[b], base, 4, dims, str, box);
      }
      if (rc != PB_OK) return rc;
    }
    break;
  }
  {
    const uint32_t ab_end = p.b_ring_off + (uint32_t)(p.b_resident ? p.ntaps * p.kchunks : p.b_stages) * p.b_bytes;
    p.e_mode = staged ? 1 : 0;
    p.e_has_add = e_has_add ? 1 : 0;
    p.e_is_add1 = a->add1 != nullptr ? 1 : 0;
    p.e_stages = (int)e_stages;
    p.e_ring_off = ab_end;
    p.smask_off = ab_end + (staged ? e_stages * V2_E_BYTES : 0u);
    p.pool = pool ? 1 : 0;
    p.pool_only = (pool && a->pool_only) ? 1 : 0;
    p.pool_off = p.smask_off + (e_masks ? 8192u : 0u);
    if (pool) {
      const uint64_t C = (uint64_t)a->Cout;
      const uint64_t dims[4] = {C, (uint64_t)a->OW / 2, (uint64_t)a->OH / 2, (uint64_t)a->N};
      const uint64_t str[3] = {C * 2, (uint64_t)(a->OW / 2) * C * 2, (uint64_t)(a->OH / 2) * (a->OW / 2) * C * 2};
      const uint32_t box[4] = {64, V2_TILE_W / 2, 2, 1};
      int rc = encode_tmap_bf16(&maps.pl, a->pool_out, 4, dims, str, box);
      if (rc != PB_OK) return rc;
    }
    if (staged && !p.pool_only) {
      const uint64_t C = (uint64_t)a->Cout;
      const uint64_t dims[4] = {C, (uint64_t)a->OW, (uint64_t)a->OH, (uint64_t)a->N};
      const uint64_t str[3] = {C * 2, (uint64_t)a->OW * C * 2, (uint64_t)a->OH * a->OW * C * 2};
      const uint32_t box[4] = {64, V2_TILE_W, 4, 1};
      int rc = encode_tmap_bf16(&maps.o, a->out, 4, dims, str, box);
      if (rc != PB_OK) return rc;
    }
    if (e_has_add) {
      const void* src = a->add1 != nullptr ? a->add1 : a->add0;
      const uint64_t C = (uint64_t)a->Cout;
      const uint64_t dims[4] = {C, (uint64_t)a->OW, (uint64_t)a->OH, (uint64_t)a->N};
      const uint64_t str[3] = {C * 2, (uint64_t)a->OW * C * 2, (uint64_t)a->OH * a->OW * C * 2};
      const uint32_t box[4] = {64, V2_TILE_W, V2_TILE_H, 1};
      int rc = encode_tmap_bf16(&maps.e, src, 4, dims, str, box);
      if (rc != PB_OK) return rc;
    }
  }
  p.acc_stages = (2 * p.T * cols_per_tile <= 512) ? 2 : 1;
  p.groups_h = cdiv(p.BH, V2_TILE_H);
  p.groups_w = cdiv(p.BW, V2_TILE_W * p.T);
  {
    // weights: bf16 [ntaps][n_tile rows][Cin], K contiguous
    const uint64_t C = (uint64_t)a->Cin;
    const uint64_t dims[3] = {C, (uint64_t)p.n_tile, (uint64_t)tp.ntaps};
    const uint64_t str[2] = {C * 2, (uint64_t)p.n_tile * C * 2};
    // streamed weight tiles are fetched in `cluster` row slices (one per CTA, multicast to all)
    p.cluster = p.pair ? 2 : 1;
    if (!p.b_resident && !p.pair) {
      // measured: only 'down' gains (r1 notes).  Never for the stride-2 'up' layers: their passes carry different tap
      // counts per CTA, and a multicast cluster needs every CTA to release every weight stage (round 2: did not terminate)
      const int want = up ? 1 : env_int("POSEB200_CONV_CLUSTER", down ? 2 : 1);
      if ((want == 2 || want == 4) && (p.n_tile % (8 * want)) == 0) p.cluster = want;
    }
    const uint32_t box[3] = {64, (uint32_t)(p.n_tile / p.cluster), 1};
    int rc = encode_tmap_bf16(&maps.b, a->w, 3, dims, str, box);
    if (rc != PB_OK) return rc;
  }
  p.Cout = a->Cout; p.up = up ? 1 : 0; p.OH = a->OH; p.OW = a->OW; p.out_nchw = a->out_nchw_f32;
  p.bias = a->bias;
  p.add0 = (const __nv_bfloat16*)a->add0; p.add1 = (const __nv_bfloat16*)a->add1;
  p.pre_out = (__nv_bfloat16*)a->pre_out; p.out = a->out;
  p.out2 = a->act_dtype == PB_F16 ? (__nv_bfloat16*)a->out2 : nullptr;
  p.mask_out = a->mask_out; p.mask_in = a->mask_in; p.act = a->act; p.slope = a->slope;
  p.a_keep_l2 = (e_has_add && (a->add1 == a->in || a->add0 == a->in) && env_int("POSEB200_CONV_KEEP_L2", 0) != 0) ? 1 : 0;
  p.use_base_offset = env_int("POSEB200_CONV_BASEOFF", 0);
  p.debug = env_int("POSEB200_CONV_DEBUG", 0);
  // measured: 2-3 % on the 64-channel layers at 192^2, neutral elsewhere
  p.poll_ns = (uint32_t)env_int("POSEB200_CONV_POLL_NS", 100);
  // measured (profiles/r2g_unrolled_issue_loop.txt): the unrolled issue loop is 6-15 % faster on every layer except the
  // 64 -> 64 ones at 192^2.  With the compile-time specialised epilogues (lighter, so the MMA stream matters again) the
  // forward of those layers gains 6 % from it too (176 -> 166 us); only their input gradient, whose heavier epilogue
  // still competes with the issuing thread, keeps the rolled loop (250 vs 253 us)
  p.unroll_taps = env_int("POSEB200_CONV_UNROLL",
                          (plain && p.kchunks == 1 && p.n_tile == 64 && tp.ntaps > 1 && a->act == PB_ACT_MASKMUL) ? 0 : 1);
  if (head != nullptr) {
    p.head_mode = head->mode;
    p.head_keys = head->keys;
    p.head_target = head->target; p.head_points = head->points;
    p.head_negk2 = head->negk2; p.head_gscale = head->gscale;
    p.head_loss = head->loss; p.head_grad = (__nv_bfloat16*)head->grad; p.head_cpad = head->cpad;
  }
  return PB_OK;
}

static int conv_tc_v2_ex(const pb_conv_args* a, const V2Head* head, cudaStream_t stream);

// tc_conv.cu calls this first; PB_ERR_UNSUPPORTED means "use the per-tap kernel"
int conv_tc_v2(const pb_conv_args* a, cudaStream_t stream) {
  if (env_int("POSEB200_CONV_V1", 0) != 0) return PB_ERR_UNSUPPORTED;
  return conv_tc_v2_ex(a, nullptr, stream);
}

static int conv_tc_v2_ex(const pb_conv_args* a, const V2Head* head, cudaStream_t stream) {
  if (a->out_nchw_f32 && (a->add0 || a->add1 || a->pre_out || a->mask_out ||
                          (a->act != PB_ACT_NONE && a->act != PB_ACT_LRELU))) return PB_ERR_UNSUPPORTED;
  if (a->taps.out_mul == 1 && a->taps.in_div == 2 && (a->OH != 2 * a->IH || a->OW != 2 * a->IW)) return PB_ERR_UNSUPPORTED;
  if (a->taps.out_mul == 2 && a->taps.in_div == 1 && (a->IH != 2 * a->OH || a->IW != 2 * a->OW)) return PB_ERR_UNSUPPORTED;
  if (a->out_nchw_f32 && a->taps.in_div == 2) {
    // the network head: its own kernel (tc_head.cu, output parities folded into N) when the shape fits it
    const int rc_head = head_tc(a, head, stream);
    if (rc_head != PB_ERR_UNSUPPORTED) return rc_head;
  }
  if (head != nullptr && head->dbias != nullptr) return PB_ERR_UNSUPPORTED;   // only tc_head.cu folds the bias gradient
  // dynamic shared memory available to this kernel: the 227 KB per-CTA limit minus its static part
  static int dyn_max = 0;
  if (dyn_max == 0) {
    int lim = 0;
    const void* fns[] = {(const void*)tc_conv2_kernel<false, false>, (const void*)tc_conv2_kernel<true, false>,
                         (const void*)tc_conv2_kernel<false, true>, (const void*)tc_conv2_kernel<true, true>,
#define V2_EPI_FN(E) (const void*)tc_conv2_kernel<true, false, (E)>,
                         V2_EPI_LIST(V2_EPI_FN)
#undef V2_EPI_FN
#define V2_EPI_FN(E) (const void*)tc_conv2_kernel<true, true, (E)>,
                         V2_EPI_LIST_F16(V2_EPI_FN)
#undef V2_EPI_FN
    };
    const int nfn = (int)(sizeof(fns) / sizeof(fns[0]));
    for (int i = 0; i < nfn; ++i) {
      cudaFuncAttributes fa;
      cudaError_t e = cudaFuncGetAttributes(&fa, fns[i]);
      if (e != cudaSuccess) return cuda_fail(e, "pb_conv_tc(v2): func attributes");
      const int l = 227 * 1024 - (int)fa.sharedSizeBytes;
      lim = (i == 0 || l < lim) ? l : lim;
    }
    for (int i = 0; i < nfn; ++i) {
      cudaError_t e = cudaFuncSetAttribute(fns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
      if (e != cudaSuccess) return cuda_fail(e, "pb_conv_tc(v2): smem attribute");
    }
    dyn_max = lim;
  }
  V2P p;
  V2Maps maps;
  int rc = v2_plan(a, p, maps, (uint32_t)dyn_max - 1024u, head);  // 1024: alignment slack of the dynamic base
  if (rc != PB_OK) return rc;
  const size_t smem = (size_t)p.smask_off + (p.e_mode && p.act == PB_ACT_MASKMUL ? 8192 : 0) + (p.pool ? 3 * 4096 : 0) + 1024;
  typedef void (*V2Kernel)(const V2Maps, const V2P);
  const bool f16 = a->act_dtype == PB_F16;
  // compile-time specialised staged epilogues (see the kernel's kEpi): bf16 cta pairs only -- every hot layer
  V2Kernel kern = p.pair ? (f16 ? tc_conv2_kernel<true, true> : tc_conv2_kernel<true, false>)
                         : (f16 ? tc_conv2_kernel<false, true> : tc_conv2_kernel<false, false>);
  if (p.pair && !p.e_mode && p.up && !f16 && p.debug == 0 && p.head_mode == 0 && !p.out_nchw && p.act == PB_ACT_LRELU &&
      !a->add0 && !a->add1 && !a->pre_out && p.out2 == nullptr && (a->Cout & 31) == 0 &&
      env_int("POSEB200_CONV_EPI_SPEC", 1) != 0) {
    kern = p.mask_out != nullptr ? tc_conv2_kernel<true, false, (EPI_SPEC | EPI_UP | EPI_LRELU | EPI_MASKOUT)>
                                 : tc_conv2_kernel<true, false, (EPI_SPEC | EPI_UP | EPI_LRELU)>;
  }
  if (p.pair && p.e_mode && p.debug == 0 && p.head_mode == 0 && !p.out_nchw &&
      (p.act == PB_ACT_LRELU || p.act == PB_ACT_MASKMUL || p.act == PB_ACT_NONE) && !(a->add0 && a->add1) &&
      env_int("POSEB200_CONV_EPI_SPEC", 1) != 0) {
    const int want = EPI_SPEC | (p.e_has_add ? EPI_ADD : 0) | ((p.e_has_add && p.e_is_add1) ? EPI_ADD1 : 0) |
                     (p.act == PB_ACT_LRELU ? EPI_LRELU : 0) | (p.act == PB_ACT_MASKMUL ? EPI_MASKMUL : 0) |
                     (p.pre_out != nullptr ? EPI_PRE : 0) | ((p.act == PB_ACT_LRELU && p.mask_out != nullptr) ? EPI_MASKOUT : 0) |
                     (p.pool ? EPI_POOL : 0) | (p.pool_only ? EPI_POOLONLY : 0) | (p.out2 != nullptr ? EPI_OUT2 : 0);
    if (!f16) {
      switch (want) {
#define V2_EPI_CASE(E) case (E): kern = tc_conv2_kernel<true, false, (E)>; break;
        V2_EPI_LIST(V2_EPI_CASE)
#undef V2_EPI_CASE
        default: break;
      }
    } else {
      switch (want) {
#define V2_EPI_CASE(E) case (E): kern = tc_conv2_kernel<true, true, (E)>; break;
        V2_EPI_LIST_F16(V2_EPI_CASE)
#undef V2_EPI_CASE
        default: break;
      }
    }
  }
  // work items; pair mode: one item = two neighbouring pixel groups, one per CTA of the pair
  const int total = p.pair ? cdiv(p.N * p.groups_h * p.groups_w, 2) * p.npass * 2 : p.N * p.groups_h * p.groups_w * p.npass;
  int grid = total < sm_count() ? total : sm_count();
  if (p.cluster > 1) {
    // co-schedulable clusters of this size (GPC boundaries can strand a few SMs); cached per (size, smem)
    static int cached_cs[3] = {0, 0, 0};      // index: pair -> 0, multicast cluster 2 -> 1, 4 -> 2
    static size_t cached_smem[3] = {0, 0, 0};
    const int ci = p.pair ? 0 : (p.cluster == 2 ? 1 : 2);
    if (cached_cs[ci] == 0 || cached_smem[ci] != smem) {
      cudaLaunchConfig_t qc;
      memset(&qc, 0, sizeof(qc));
      cudaLaunchAttribute qa;
      qa.id = cudaLaunchAttributeClusterDimension;
      qa.val.clusterDim.x = (unsigned)p.cluster; qa.val.clusterDim.y = 1; qa.val.clusterDim.z = 1;
      qc.gridDim = dim3((unsigned)(sm_count() / p.cluster * p.cluster));
      qc.blockDim = dim3(V2_THREADS);
      qc.dynamicSmemBytes = smem;
      qc.attrs = &qa; qc.numAttrs = 1;
      int ncl = 0;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, kern, &qc);
      if (e != cudaSuccess) { cudaGetLastError(); ncl = 0; }
      cached_cs[ci] = ncl > 0 ? ncl : -1;
      cached_smem[ci] = smem;
    }
    if (cached_cs[ci] > 0) {
      int ctas = cached_cs[ci] * p.cluster;
      const int need = cdiv(total, p.cluster) * p.cluster;
      grid = ctas < need ? ctas : need;
    } else {
      set_error("pb_conv_tc(v2): no co-schedulable cluster of %d CTAs (set POSEB200_CONV_CLUSTER=1)", p.cluster);
      return PB_ERR_CUDA;
    }
  }
  p.iters = cdiv(total, grid);
  if (p.cluster > 1) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = (unsigned)p.cluster; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(V2_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, maps, p);
    if (e != cudaSuccess) return cuda_fail(e, "pb_conv_tc(v2): cluster launch");
  } else {
    kern<<<grid, V2_THREADS, smem, stream>>>(maps, p);
  }
  PB_LAUNCH_CHECK("tc_conv2_kernel");
  return PB_OK;
}

int launch_zero_u64(unsigned long long* p, int n, cudaStream_t st);                          // bandwidth.cu
int launch_argmax_finalize(float* peaks, float* values, int maps, int W, cudaStream_t st);     // bandwidth.cu

// shared argument checks of the two fused-head entry points; fills the conv descriptor the kernel plans from
static int head_prepare(const pb_head_fused_args* h, const char* fn, pb_conv_args& c) {
  if (h == nullptr) { set_error("%s: null args", fn); return PB_ERR_INVALID; }
  c = h->conv;
  const pb_taps& tp = c.taps;
  if (!(tp.out_mul == 1 && tp.in_div == 2) || c.OH != 2 * c.IH || c.OW != 2 * c.IW) {
    set_error("%s: the fused head is a stride-2 transposed convolution (CNNs.py:125-128)", fn);
    return PB_ERR_INVALID;
  }
  if (c.add0 || c.add1 || c.pre_out || c.mask_out || (c.act != PB_ACT_NONE && c.act != PB_ACT_LRELU)) {
    set_error("%s: bias + LeakyReLU epilogue only", fn);
    return PB_ERR_INVALID;
  }
  if (c.act_dtype != PB_BF16 && c.act_dtype != PB_F16) { set_error("%s: bf16 / fp16 activations only", fn); return PB_ERR_UNSUPPORTED; }
  c.out_nchw_f32 = 1;
  c.out2 = nullptr;
  if (c.out == nullptr) c.out = const_cast<void*>(c.in);   // never written in the fused modes; keeps the shared checks happy
  return conv_args_check(&c, fn);
}

}  // namespace pb

using namespace pb;

extern "C" int pb_convT_argmax_fused(const pb_head_fused_args* h, void* stream) {
  pb_conv_args c;
  int rc = head_prepare(h, "pb_convT_argmax_fused", c);
  if (rc != PB_OK) return rc;
  PB_REQUIRE(h->peaks != nullptr, "pb_convT_argmax_fused: peaks is required");
  PB_REQUIRE_DEV(h->peaks, "peaks");
  PB_REQUIRE_DEV(h->values, "values");
  PB_REQUIRE((long long)c.OH * c.OW < 0xFFFFFFFFll, "pb_convT_argmax_fused: map too large for 32-bit flat indices");
  cudaStream_t st = (cudaStream_t)stream;
  const int maps = c.N * c.Cout;
  rc = launch_zero_u64(reinterpret_cast<unsigned long long*>(h->peaks), maps, st);
  if (rc != PB_OK) return rc;
  V2Head hd;
  memset(&hd, 0, sizeof(hd));
  hd.mode = 1;
  hd.keys = reinterpret_cast<unsigned long long*>(h->peaks);
  rc = conv_tc_v2_ex(&c, &hd, st);
  if (rc != PB_OK) {
    if (rc == PB_ERR_UNSUPPORTED) set_error("pb_convT_argmax_fused: shape outside the halo kernel's tiling");
    return rc;
  }
  return launch_argmax_finalize(h->peaks, h->values, maps, c.OW, st);
}

extern "C" int pb_convT_mse_fused(const pb_head_fused_args* h, void* stream) {
  pb_conv_args c;
  int rc = head_prepare(h, "pb_convT_mse_fused", c);
  if (rc != PB_OK) return rc;
  PB_REQUIRE(h->loss_sum != nullptr && h->grad_nhwc != nullptr, "pb_convT_mse_fused: loss_sum and grad_nhwc are required");
  PB_REQUIRE((h->target != nullptr) != (h->points != nullptr), "pb_convT_mse_fused: exactly one of target / points");
  PB_REQUIRE(h->Cpad >= c.Cout && h->Cpad % 16 == 0 && h->Cpad == (c.Cout + 15) / 16 * 16,
             "pb_convT_mse_fused: Cpad must be Cout rounded up to 16");
  PB_REQUIRE(h->points == nullptr || h->sigma > 0.f, "pb_convT_mse_fused: sigma must be positive");
  PB_REQUIRE_DEV(h->loss_sum, "loss_sum");
  PB_REQUIRE_DEV(h->grad_nhwc, "grad_nhwc");
  PB_REQUIRE_DEV(h->target, "target");
  PB_REQUIRE_DEV(h->points, "points");
  V2Head hd;
  memset(&hd, 0, sizeof(hd));
  hd.mode = 2;
  hd.target = h->target; hd.points = h->points;
  hd.negk2 = h->points != nullptr ? -(1.0f / (2.0f * h->sigma * h->sigma)) * 1.4426950408889634f : 0.f;
  hd.gscale = h->grad_scale;
  hd.loss = h->loss_sum; hd.grad = h->grad_nhwc; hd.cpad = h->Cpad;
  PB_REQUIRE_DEV(h->dbias, "dbias");
  PB_REQUIRE(h->dbias == nullptr || h->dbias_rows >= sm_count(), "pb_convT_mse_fused: dbias needs one row per SM");
  hd.dbias = h->dbias;
  rc = conv_tc_v2_ex(&c, &hd, (cudaStream_t)stream);
  if (rc == PB_ERR_UNSUPPORTED) set_error("pb_convT_mse_fused: shape outside the halo kernel's tiling");
  return rc;
}
