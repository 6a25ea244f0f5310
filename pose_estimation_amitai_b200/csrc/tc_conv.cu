// tcgen05 implicit-GEMM gather-convolution (bf16 operands, fp32 accumulation in TMEM).
//
// One persistent CTA per SM, warp-specialised:
//   warp 0 (1 lane)  TMA producer : per K-step one NHWC activation box [TH x TW pixels x 64 ch]
//                                   (OOB zero fill = conv padding) + one weight box [n_tile x 64 ch]
//   warp 1 (1 lane)  MMA issuer   : 4 x tcgen05.mma (M=128, N=n_tile, K=16) per K-step into TMEM
//   warps 2..5       epilogue     : tcgen05.ld -> bias / skip add / LeakyReLU (+sign mask) /
//                                   LeakyReLU' / residual -> 16-byte NHWC stores (or NCHW fp32)
// K-steps = taps x (Cin/64).  A tile of M=128 output pixels is a TH x TW patch of one image so the
// activation operand of every tap is a plain shifted TMA box (no im2col buffer in HBM).
//
// Three geometries share the kernel (see pb_conv_args):
//   plain  : conv / convT stride 1 / linear          -- 1 accumulator
//   up     : convT stride 2 forward                   -- 4 accumulators (output parity phases)
//   down   : convT stride 2 input gradient            -- 4 phase-strided tensor maps over the input
#include <stdlib.h>
#include <string.h>

#include "tc_epilogue.cuh"

namespace pb {

using namespace tc;

constexpr int TC_THREADS = 320;   // TMA producer, MMA issuer, 8 epilogue warps (two per TMEM lane quadrant)
constexpr int TC_BLOCK_K = 64;
constexpr int TC_A_BYTES = 128 * TC_BLOCK_K * 2;  // 16 KB
constexpr int TC_MAX_STAGES = 8;

struct TmapPack {
  CUtensorMap a[4];
  CUtensorMap b;
  CUtensorMap o;   // TE_TMA: the output matrix [rows][Cout], box 32 rows x 64 columns
};

struct TcTap {
  int8_t dy, dx, acc, map;
};

struct TcConvP : EpiP {
  int N, TH, TW, tiles_h, tiles_w, BH, BW;
  int kchunks, ntaps;
  TcTap taps[PB_MAX_TAPS];
  int n_acc, n_tile, n_tiles, acc_stages, stages;
  uint32_t stage_bytes;
  int up, OH, OW, out_nchw;
  uint32_t epi_off;   // TE_TMA: byte offset of the output staging slots behind the operand ring
  uint32_t b_off, b_chunk_bytes;   // TE_BRES: the resident weight tile, kchunks x [n_tile rows x 128 B]
};

// kEpi != 0: the epilogue's features are fixed at compile time (bit mask) -- the flag tests, the absent operands' code and
// the NCHW / stride-2 output paths drop out (nn.Linear of the ViT: pytorch_vit_encoder.py:20-23,52,55,122).  The host
// picks a specialised instantiation only when every flag matches one in TC_EPI_LIST.
//
// TE_TMA (token matrices, no skip operand): the result leaves through shared memory as TMA stores.  A lane of the
// register path owns one ROW, so a warp-wide store touches 32 rows 2*Cout bytes apart -- 64-byte pieces of 32 different
// DRAM pages; with K = 256 and N = 9216 (to_qkv, pytorch_vit_encoder.py:52) the 170 MB result then paces the kernel at
// 2.6 TB/s.  Staged, every quadrant (32 rows) x 64 columns is one 4 KB swizzled slot that leaves as full 128-byte lines.
//
// TE_BRES (with TE_TMA, K <= 256): the weight tile stays RESIDENT in shared memory and the CTA walks a contiguous range
// of M tiles under it (tiles ordered n-major), so that only the 16 KB activation chunks stream.  With K = 256 a 128 x 256
// tile needs 192 KB of operands for 16.8 MFLOP -- to_qkv moved 498 MB of operands + 170 MB of results per launch and
// ran at the ~12.5 TB/s the SMs get from L2 (the same rate its K = 9216 input gradient reaches), not at the tensor pipe.
constexpr int TE_SPEC = 1, TE_BIAS = 2, TE_ADD0 = 4, TE_ADD1 = 8, TE_PRE = 16, TE_GELU = 32, TE_TMA = 64, TE_BRES = 128;
constexpr int TC_EPI_SLOTS = 3;                                  // staging slots per TMEM lane quadrant
constexpr int TC_EPI_BYTES = 4 * TC_EPI_SLOTS * 4096;            // 48 KB
#define TC_EPI_LIST(X)                            \
  X(TE_SPEC)                                      \
  X(TE_SPEC | TE_BIAS)                            \
  X(TE_SPEC | TE_BIAS | TE_ADD1)                  \
  X(TE_SPEC | TE_BIAS | TE_GELU | TE_PRE)         \
  X(TE_SPEC | TE_ADD0)                            \
  X(TE_SPEC | TE_TMA)                             \
  X(TE_SPEC | TE_BIAS | TE_TMA)                   \
  X(TE_SPEC | TE_TMA | TE_BRES)                   \
  X(TE_SPEC | TE_BIAS | TE_TMA | TE_BRES)

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(m),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}

template <bool kF16, int kEpi = 0>   // kF16: operands / skip tensors / outputs in IEEE half instead of bf16 ("fp16" precision forward)
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_conv_kernel(const __grid_constant__ TmapPack maps, const TcConvP p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ __align__(8) uint64_t b_full_bar, b_empty_bar;   // TE_BRES: the resident weight tile
  __shared__ uint32_t tmem_slot;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) prefetch_tmap(&maps.a[i]);
    prefetch_tmap(&maps.b);
    if (kEpi & TE_TMA) prefetch_tmap(&maps.o);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 8);
    }
    mbar_init(&b_full_bar, 1);
    mbar_init(&b_empty_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  const int m_tiles = p.N * p.tiles_h * p.tiles_w;
  const int total_tiles = m_tiles * p.n_tiles;
  const uint32_t idesc = make_idesc(128, p.n_tile, 0, 0, kF16, kF16);
  constexpr bool kBres = (kEpi & TE_BRES) != 0;
  // TE_BRES: tiles ordered n-major (tile = nt * m_tiles + mt), every CTA a contiguous range of them
  const int r_begin = kBres ? (int)(((long long)total_tiles * blockIdx.x) / gridDim.x) : 0;
  const int r_end = kBres ? (int)(((long long)total_tiles * (blockIdx.x + 1)) / gridDim.x) : 0;

  if (kBres && warp == 0) {
    if (elect_one()) {
      // ------------------------------------------------------------------ TMA producer, resident weight tile
      int stage = 0, cur_nt = -1, switches = 0;
      uint32_t phase = 0;
      for (int tile = r_begin; tile < r_end; ++tile) {
        const int nt = tile / m_tiles, tw = tile % m_tiles;
        if (nt != cur_nt) {
          if (cur_nt >= 0) { mbar_wait(&b_empty_bar, (uint32_t)(switches & 1)); ++switches; }   // every MMA on the old tile retired
          mbar_expect_tx(&b_full_bar, (uint32_t)p.kchunks * p.b_chunk_bytes);
          for (int kc = 0; kc < p.kchunks; ++kc)
            tma_load_3d(smem + p.b_off + (size_t)kc * p.b_chunk_bytes, &maps.b, &b_full_bar, kc * TC_BLOCK_K, nt * p.n_tile, 0);
          cur_nt = nt;
        }
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], TC_A_BYTES);
          tma_load_4d(smem + (size_t)stage * TC_A_BYTES, &maps.a[0], &full_bar[stage], kc * TC_BLOCK_K, tw * 128, 0, 0);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (kBres && warp == 1) {
    if (elect_one()) {
      // ------------------------------------------------------------------ MMA issuer, resident weight tile
      int stage = 0, cur_nt = -1, loads = 0, it = 0;
      uint32_t phase = 0;
      for (int tile = r_begin; tile < r_end; ++tile, ++it) {
        const int nt = tile / m_tiles;
        if (nt != cur_nt) {
          mbar_wait(&b_full_bar, (uint32_t)(loads & 1));
          ++loads;
          cur_nt = nt;
        }
        const int as = it % p.acc_stages;
        const uint32_t aphase = (uint32_t)(it / p.acc_stages) & 1u;
        mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.n_tile);
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + (size_t)stage * TC_A_BYTES);
          const uint32_t b_addr = smem_u32(smem + p.b_off + (size_t)kc * p.b_chunk_bytes);
#pragma unroll
          for (int j = 0; j < TC_BLOCK_K / 16; ++j) {
            const uint64_t ad = smem_desc_sw128(a_addr + j * 32, 16, 1024);
            const uint64_t bd = smem_desc_sw128(b_addr + j * 32, 16, 1024);
            umma_bf16(d_tmem, ad, bd, idesc, (kc > 0 || j > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full_bar[as]);
        if (tile + 1 < r_end && (tile + 1) / m_tiles != nt) umma_commit(&b_empty_bar);   // the producer may overwrite the tile
      }
    }
  } else if (warp == 0) {
    if (elect_one()) {
      // ------------------------------------------------------------------ TMA producer
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles;
        int mt = tile / p.n_tiles;
        const int tw = mt % p.tiles_w; mt /= p.tiles_w;
        const int th = mt % p.tiles_h;
        const int img = mt / p.tiles_h;
        const int h0 = th * p.TH, w0 = tw * p.TW, n0 = nt * p.n_tile;
        for (int t = 0; t < p.ntaps; ++t) {
          const TcTap tap = p.taps[t];
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + (size_t)stage * p.stage_bytes;
            mbar_expect_tx(&full_bar[stage], p.stage_bytes);
            tma_load_4d(sa, &maps.a[tap.map], &full_bar[stage], kc * TC_BLOCK_K, w0 + tap.dx, h0 + tap.dy, img);
            tma_load_3d(sa + TC_A_BYTES, &maps.b, &full_bar[stage], kc * TC_BLOCK_K, n0, t);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ------------------------------------------------------------------ MMA issuer
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int as = it % p.acc_stages;
        const uint32_t aphase = (uint32_t)(it / p.acc_stages) & 1u;
        mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
        tc_fence_after();
        uint32_t started = 0;
        for (int t = 0; t < p.ntaps; ++t) {
          const int acc = p.taps[t].acc;
          const uint32_t d_tmem = tmem_base + (uint32_t)((as * p.n_acc + acc) * p.n_tile);
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + (size_t)stage * p.stage_bytes);
            const uint32_t b_addr = a_addr + TC_A_BYTES;
#pragma unroll
            for (int j = 0; j < TC_BLOCK_K / 16; ++j) {
              const uint64_t ad = smem_desc_sw128(a_addr + j * 32, 16, 1024);
              const uint64_t bd = smem_desc_sw128(b_addr + j * 32, 16, 1024);
              umma_bf16(d_tmem, ad, bd, idesc, ((started >> acc) & 1u) | (j > 0 ? 1u : 0u));
            }
            started |= 1u << acc;
            umma_commit(&empty_bar[stage]);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
        umma_commit(&tmem_full_bar[as]);
      }
    }
  } else {
    // -------------------------------------------------------------------- epilogue warps
    const int q = warp & 3;          // TMEM lane quadrant this warp may access
    const int hsel = (warp - 2) >> 2;  // the two warps of a quadrant take alternate 32-channel chunks
    // the epilogue's view of the parameters; a specialised instantiation overrides every feature with its constant
    EpiP ep = static_cast<const EpiP&>(p);
    if (kEpi) {
      if (!(kEpi & TE_BIAS)) ep.bias = nullptr;
      if (!(kEpi & TE_ADD0)) ep.add0 = nullptr;
      if (!(kEpi & TE_ADD1)) ep.add1 = nullptr;
      if (!(kEpi & TE_PRE)) ep.pre_out = nullptr;
      ep.mask_out = nullptr; ep.mask_in = nullptr; ep.out2 = nullptr;
      ep.act = (kEpi & TE_GELU) ? PB_ACT_GELU : PB_ACT_NONE;
      if (kEpi & TE_BIAS) __builtin_assume(ep.bias != nullptr);
      if (kEpi & TE_ADD0) __builtin_assume(ep.add0 != nullptr);
      if (kEpi & TE_ADD1) __builtin_assume(ep.add1 != nullptr);
      if (kEpi & TE_PRE) __builtin_assume(ep.pre_out != nullptr);
    }
    const bool out_nchw = kEpi ? false : p.out_nchw != 0;
    const bool up = kEpi ? false : p.up != 0;
    const int n_acc = kEpi ? 1 : p.n_acc;
    int it = 0;
    int eslot = 0;
    const int t_begin = kBres ? r_begin : (int)blockIdx.x, t_end = kBres ? r_end : total_tiles;
    const int t_step = kBres ? 1 : (int)gridDim.x;
    for (int tile = t_begin; tile < t_end; tile += t_step, ++it) {
      const int as = it % p.acc_stages;
      const uint32_t aphase = (uint32_t)(it / p.acc_stages) & 1u;
      const int nt = kBres ? tile / m_tiles : tile % p.n_tiles;
      int mt = kBres ? tile % m_tiles : tile / p.n_tiles;
      const int tw = mt % p.tiles_w; mt /= p.tiles_w;
      const int th = mt % p.tiles_h;
      const int img = mt / p.tiles_h;
      const int ml = q * 32 + lane;
      const int bh = th * p.TH + ml / p.TW, bw = tw * p.TW + ml % p.TW;
      const bool ok = bh < p.BH && bw < p.BW;
      const int n0 = nt * p.n_tile;
      mbar_wait(&tmem_full_bar[as], aphase);
      tc_fence_after();
      const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
      if constexpr ((kEpi & TE_TMA) != 0) {
        // token matrix (one image row of BW tokens, TH = 1, TW = 128): this quadrant's 32 rows start at row0.  The two
        // warps of the quadrant fill the halves of a 4 KB slot (row = lane, 16-byte piece j at j ^ (row & 7): the layout
        // a SWIZZLE_128B box has in shared memory); rows / columns past the matrix are clipped by the store itself.
        const int row0 = tw * 128 + q * 32;
        const int nb = p.n_tile / 64;
        const uint32_t tcol = lane_base + (uint32_t)(as * p.n_tile + hsel * 32);
        // the accumulator chunk of box b + 1 is in flight while box b is converted, staged and stored
        uint32_t r[2][32];
        tmem_ld32(tcol, r[0]);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          if (b < nb) {
            tmem_ld_wait();
            if (b + 1 < nb) tmem_ld32(tcol + (uint32_t)((b + 1) * 64), r[(b + 1) & 1]);
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[b & 1][j]);
            if constexpr ((kEpi & TE_BIAS) != 0) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += __ldg(ep.bias + n0 + b * 64 + hsel * 32 + j);
            }
            uint8_t* slot = smem + p.epi_off + (size_t)((q * TC_EPI_SLOTS + eslot) * 4096);
            uint8_t* erow = slot + lane * 128;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              *reinterpret_cast<uint4*>(erow + (((hsel * 4 + k) ^ (lane & 7)) << 4)) = pack16x8<kF16>(v + 8 * k);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
            if (hsel == 0) {
              if (elect_one()) {
                tma_store_2d(&maps.o, slot, n0 + b * 64, row0);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                // retire the PREVIOUS slot's read, not this one; with three slots the slot written next was retired one
                // step earlier, before this warp reached the pair barrier above, so both warps see it free
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
              }
            }
            __syncwarp();
            if (++eslot == TC_EPI_SLOTS) eslot = 0;
          }
        }
      } else if (out_nchw) {
        if (hsel == 0) {
        float* outf = reinterpret_cast<float*>(p.out);
        const int nph = up ? 2 : 1;
        for (int py = 0; py < nph; ++py) {
          const int oy = up ? 2 * bh + py : bh;
          for (int c0 = 0; c0 < p.n_tile; c0 += 16) {
            uint32_t r0[16], r1[16];
            const int a0 = up ? py * 2 : 0;
            tmem_ld16(lane_base + (uint32_t)((as * n_acc + a0) * p.n_tile + c0), r0);
            if (up) tmem_ld16(lane_base + (uint32_t)((as * n_acc + a0 + 1) * p.n_tile + c0), r1);
            tmem_ld_wait();
            const int ox0 = up ? 2 * bw : bw;
            const long long pix0 = ((long long)img * p.OH + oy) * p.OW + ox0;
            epilogue_chunk<16, kF16>(ep, r0, pix0, n0 + c0, ok);
            if (up) epilogue_chunk<16, kF16>(ep, r1, pix0 + 1, n0 + c0, ok);
            if (ok) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int c = n0 + c0 + j;
                if (c < p.Cout) {
                  float* dst = outf + (((long long)img * p.Cout + c) * p.OH + oy) * p.OW + ox0;
                  if (up) *reinterpret_cast<float2*>(dst) = make_float2(__uint_as_float(r0[j]), __uint_as_float(r1[j]));
                  else *dst = __uint_as_float(r0[j]);
                }
              }
            }
          }
        }
        }
      } else {
        for (int a = 0; a < n_acc; ++a) {
          const int oy = up ? 2 * bh + (a >> 1) : bh;
          const int ox = up ? 2 * bw + (a & 1) : bw;
          const long long pix = ((long long)img * p.OH + oy) * p.OW + ox;
          const uint32_t col0 = (uint32_t)((as * n_acc + a) * p.n_tile);
          int c0 = 0, k = 0;
          for (; c0 + 32 <= p.n_tile; c0 += 32, ++k) {
            if ((k & 1) != hsel) continue;
            uint32_t r[32];
            tmem_ld32(lane_base + col0 + c0, r);
            EpiPre e;
            e.pix = pix; e.c0 = n0 + c0; e.width = 32; e.ok = ok;
            epi_prefetch(ep, e);            // skip / residual rows are in flight while the accumulator chunk arrives
            tmem_ld_wait();
            if (e.fast) {
              epi32_fast_gbias<kF16>(ep, r, e);   // 256-bit stores: one full sector per lane and instruction
            } else {
              epilogue_chunk<32, kF16>(ep, r, pix, n0 + c0, ok);
              store_nhwc<32, kF16>(ep, r, pix, n0 + c0, ok);
            }
          }
          if (c0 < p.n_tile && (k & 1) == hsel) {
            uint32_t r[16];
            tmem_ld16(lane_base + col0 + c0, r);
            tmem_ld_wait();
            epilogue_chunk<16, kF16>(ep, r, pix, n0 + c0, ok);
            store_nhwc<16, kF16>(ep, r, pix, n0 + c0, ok);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
    }
  }
  if ((kEpi & TE_TMA) && warp >= 2 && warp <= 5) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ----------------------------------------------------------------------------- host side
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode_fn() {
  static EncodeFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || sym == nullptr) {
      cudaGetLastError();
      return nullptr;
    }
    fn = reinterpret_cast<EncodeFn>(sym);
  }
  return fn;
}

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box) {
  EncodeFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (driver too old / no GPU)");
    return PB_ERR_CUDA;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu] box [%u %u %u %u]", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
              rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return PB_ERR_CUDA;
  }
  return PB_OK;
}

int conv_args_check(const pb_conv_args* a, const char* fn);
int conv_tc_v2(const pb_conv_args* a, cudaStream_t stream);  // tc_conv2.cu

static void pick_tile(int bh, int bw, int* th, int* tw) {
  // 128 pixels per tile; prefer the shape that wastes the fewest out-of-range pixels
  const int cand[][2] = {{8, 16}, {16, 8}, {4, 32}, {2, 64}, {1, 128}, {32, 4}};
  long long best = -1;
  for (auto& c : cand) {
    const long long cover = (long long)cdiv(bh, c[0]) * c[0] * cdiv(bw, c[1]) * c[1];
    if (best < 0 || cover < best) { best = cover; *th = c[0]; *tw = c[1]; }
  }
}

}  // namespace pb

using namespace pb;

extern "C" {

int pb_conv_tc(const pb_conv_args* a, void* stream) {
  int rc = conv_args_check(a, "pb_conv_tc");
  if (rc != PB_OK) return rc;
  if ((a->act_dtype != PB_BF16 && a->act_dtype != PB_F16) || a->in_nchw_f32) {
    set_error("pb_conv_tc: bf16 / fp16 NHWC activations only");
    return PB_ERR_UNSUPPORTED;
  }
  rc = conv_tc_v2(a, (cudaStream_t)stream);  // halo-resident kernel; falls through when it does not tile the shape
  if (rc != PB_ERR_UNSUPPORTED) return rc;
  if (a->pool_out != nullptr) {
    set_error("pb_conv_tc: pool_out needs the halo kernel's staged epilogue (stride-1 layer, Cout %% 64 == 0, even OH / OW)");
    return PB_ERR_UNSUPPORTED;
  }
  const pb_taps& tp = a->taps;
  const bool plain = tp.out_mul == 1 && tp.in_div == 1;
  const bool up = tp.out_mul == 1 && tp.in_div == 2;
  const bool down = tp.out_mul == 2 && tp.in_div == 1;
  if (!(plain || up || down)) {
    set_error("pb_conv_tc: unsupported (out_mul, in_div) = (%d, %d)", tp.out_mul, tp.in_div);
    return PB_ERR_UNSUPPORTED;
  }
  if ((a->Cin & 7) != 0) {
    set_error("pb_conv_tc: Cin (%d) must be a multiple of 8 (16-byte TMA rows)", a->Cin);
    return PB_ERR_UNSUPPORTED;
  }
  if (up && (a->OH != 2 * a->IH || a->OW != 2 * a->IW)) {
    set_error("pb_conv_tc: stride-2 transposed conv needs OH=2*IH");
    return PB_ERR_INVALID;
  }
  if (down && (a->IH != 2 * a->OH || a->IW != 2 * a->OW)) {
    set_error("pb_conv_tc: stride-2 gather needs IH=2*OH");
    return PB_ERR_INVALID;
  }
  if (a->out_nchw_f32 && (a->add0 || a->add1 || a->pre_out || a->mask_out)) {
    set_error("pb_conv_tc: NCHW output supports bias+activation only");
    return PB_ERR_UNSUPPORTED;
  }

  TcConvP p;
  memset((void*)&p, 0, sizeof(p));
  p.N = a->N;
  p.BH = up ? a->IH : a->OH;
  p.BW = up ? a->IW : a->OW;
  pick_tile(p.BH, p.BW, &p.TH, &p.TW);
  p.tiles_h = cdiv(p.BH, p.TH);
  p.tiles_w = cdiv(p.BW, p.TW);
  p.kchunks = cdiv(a->Cin, TC_BLOCK_K);
  p.ntaps = tp.ntaps;
  p.n_acc = up ? 4 : 1;
  const int cout_pad = cdiv(a->Cout, 16) * 16;
  const int max_tile = up ? 128 : 256;
  if (cout_pad <= max_tile) {
    p.n_tile = cout_pad;
    p.n_tiles = 1;
  } else {
    p.n_tile = up ? 128 : 256;
    while (cout_pad % p.n_tile != 0) p.n_tile -= 32;
    if (p.n_tile < 32) {
      set_error("pb_conv_tc: cannot tile Cout=%d", a->Cout);
      return PB_ERR_UNSUPPORTED;
    }
    p.n_tiles = cout_pad / p.n_tile;
  }
  // few M tiles (token matrices with 256 output columns: 72 tiles on 148 SMs): split N so that every SM gets a tile
  {
    const int m_tiles = p.N * p.tiles_h * p.tiles_w;
    while (m_tiles * p.n_tiles * 2 <= sm_count() + sm_count() / 8 && p.n_tile >= 128 && (p.n_tile / 2) % 32 == 0) {   // 32: sign-mask words stay within one tile
      p.n_tile /= 2;
      p.n_tiles *= 2;
    }
  }
  p.acc_stages = (2 * p.n_acc * p.n_tile <= 512) ? 2 : 1;
  for (int t = 0; t < tp.ntaps; ++t) {
    const int dy = tp.dy[t], dx = tp.dx[t];
    TcTap& k = p.taps[t];
    if (plain) {
      k.dy = (int8_t)dy; k.dx = (int8_t)dx; k.acc = 0; k.map = 0;
    } else if (up) {
      const int py = dy & 1, px = dx & 1;
      k.dy = (int8_t)((dy + py) / 2); k.dx = (int8_t)((dx + px) / 2);
      k.acc = (int8_t)(py * 2 + px); k.map = 0;
    } else {
      const int py = dy & 1, px = dx & 1;
      k.dy = (int8_t)((dy - py) / 2); k.dx = (int8_t)((dx - px) / 2);
      k.acc = 0; k.map = (int8_t)(py * 2 + px);
    }
  }
  p.stage_bytes = (uint32_t)(TC_A_BYTES + p.n_tile * TC_BLOCK_K * 2);
  p.stages = (int)((220 * 1024) / p.stage_bytes);
  if (p.stages > TC_MAX_STAGES) p.stages = TC_MAX_STAGES;
  p.Cout = a->Cout; p.up = up ? 1 : 0; p.OH = a->OH; p.OW = a->OW; p.out_nchw = a->out_nchw_f32;
  p.bias = a->bias;
  p.add0 = (const __nv_bfloat16*)a->add0; p.add1 = (const __nv_bfloat16*)a->add1;
  p.pre_out = (__nv_bfloat16*)a->pre_out; p.out = a->out;
  p.out2 = a->act_dtype == PB_F16 ? (__nv_bfloat16*)a->out2 : nullptr;
  p.mask_out = a->mask_out; p.mask_in = a->mask_in; p.act = a->act; p.slope = a->slope;

  // token matrices without a skip operand (nn.Linear forward without residual, its input gradient): staged TMA-store
  // epilogue (TE_TMA) and, for K <= 256, the resident weight tile (TE_BRES).  Decided here because both change the
  // shared-memory plan and TE_BRES the N tiling the weight map is encoded with.
  int tma_bits = 0;
  size_t smem_tma = 0;
  {
    static int tma_on = -1, bres_on = -1;
    if (tma_on < 0) {
      const char* v = getenv("POSEB200_LINEAR_TMA_EPI");
      const char* sp = getenv("POSEB200_CONV_EPI_SPEC");   // both forms are specialised instantiations
      tma_on = ((v != nullptr && v[0] == '0') || (sp != nullptr && sp[0] == '0')) ? 0 : 1;
    }
    if (bres_on < 0) { const char* v = getenv("POSEB200_LINEAR_BRES"); bres_on = (v != nullptr && v[0] == '0') ? 0 : 1; }
    const bool no_skip = a->add0 == nullptr && a->add1 == nullptr && a->pre_out == nullptr && a->mask_out == nullptr &&
                         a->mask_in == nullptr && a->out2 == nullptr && a->act == PB_ACT_NONE;
    if (tma_on && no_skip && plain && a->act_dtype == PB_BF16 && !a->out_nchw_f32 && tp.ntaps == 1 && p.N == 1 &&
        p.BH == 1 && p.TH == 1 && p.TW == 128 && a->Cout % 64 == 0 && p.kchunks <= 16) {
      const int m_tiles = p.tiles_w;
      // N tile of the resident form: <= 96 KB of weights, a multiple of 64 that divides Cout.  Measured at 9216 rows, K =
      // 256 (tools/lin_bench.py, profiles/r2x_linear_epilogue.txt): it pays for wide results (N = 9216: 51.2 -> 49.2 us,
      // N = 3072: 23.3 -> 21.3) and costs at N = 1024 (13.1 -> 15.1), hence the >= 16 N tiles.
      int nt_res = 0;
      if (bres_on && p.kchunks <= 4 && m_tiles >= 16)
        for (int c = 256; c >= 64; c -= 64)
          if (a->Cout % c == 0 && (size_t)c * p.kchunks * 128 <= 96 * 1024) {
            if (a->Cout / c >= 16 && m_tiles * (a->Cout / c) >= 2 * sm_count()) nt_res = c;
            break;
          }
      if (nt_res > 0) {
        p.n_tile = nt_res;
        p.n_tiles = a->Cout / nt_res;
        p.acc_stages = 2;
        p.b_chunk_bytes = (uint32_t)nt_res * 128;
        const uint32_t b_bytes = (uint32_t)p.kchunks * p.b_chunk_bytes;
        const int st = (int)((224 * 1024 - TC_EPI_BYTES - b_bytes) / TC_A_BYTES);
        p.stages = st > TC_MAX_STAGES ? TC_MAX_STAGES : st;
        p.b_off = (uint32_t)p.stages * TC_A_BYTES;
        p.epi_off = p.b_off + b_bytes;
        smem_tma = (size_t)p.epi_off + TC_EPI_BYTES + 1024;
        tma_bits = TE_TMA | TE_BRES;
      } else if (p.n_tile % 64 == 0) {
        const int st = (int)((220 * 1024 - TC_EPI_BYTES) / p.stage_bytes);
        if (st >= 2) {
          p.stages = st > TC_MAX_STAGES ? TC_MAX_STAGES : st;
          p.epi_off = (uint32_t)p.stages * p.stage_bytes;
          smem_tma = (size_t)p.epi_off + TC_EPI_BYTES + 1024;
          tma_bits = TE_TMA;
        }
      }
    }
  }

  TmapPack maps;
  memset(&maps, 0, sizeof(maps));
  const uint64_t C = (uint64_t)a->Cin;
  if (tma_bits) {
    const uint64_t odims[2] = {(uint64_t)a->Cout, (uint64_t)p.BW};
    const uint64_t ostr[1] = {(uint64_t)a->Cout * 2};
    const uint32_t obox[2] = {64, 32};
    rc = encode_tmap_bf16(&maps.o, a->out, 2, odims, ostr, obox);
    if (rc != PB_OK) return rc;
  }
  if (!down) {
    const uint64_t dims[4] = {C, (uint64_t)a->IW, (uint64_t)a->IH, (uint64_t)a->N};
    const uint64_t str[3] = {C * 2, (uint64_t)a->IW * C * 2, (uint64_t)a->IH * a->IW * C * 2};
    const uint32_t box[4] = {TC_BLOCK_K, (uint32_t)p.TW, (uint32_t)p.TH, 1};
    rc = encode_tmap_bf16(&maps.a[0], a->in, 4, dims, str, box);
    if (rc != PB_OK) return rc;
    for (int i = 1; i < 4; ++i) maps.a[i] = maps.a[0];
  } else {
    for (int ph = 0; ph < 4; ++ph) {
      const int py = ph >> 1, px = ph & 1;
      const __nv_bfloat16* base = (const __nv_bfloat16*)a->in + ((size_t)py * a->IW + px) * C;
      const uint64_t dims[4] = {C, (uint64_t)a->IW / 2, (uint64_t)a->IH / 2, (uint64_t)a->N};
      const uint64_t str[3] = {2 * C * 2, 2 * (uint64_t)a->IW * C * 2, (uint64_t)a->IH * a->IW * C * 2};
      const uint32_t box[4] = {TC_BLOCK_K, (uint32_t)p.TW, (uint32_t)p.TH, 1};
      rc = encode_tmap_bf16(&maps.a[ph], base, 4, dims, str, box);
      if (rc != PB_OK) return rc;
    }
  }
  {
    // weights: bf16 [ntaps][n_tiles*n_tile rows][Cin], K contiguous
    const uint64_t rows = (uint64_t)p.n_tiles * p.n_tile;
    const uint64_t dims[3] = {C, rows, (uint64_t)tp.ntaps};
    const uint64_t str[2] = {C * 2, rows * C * 2};
    const uint32_t box[3] = {TC_BLOCK_K, (uint32_t)p.n_tile, 1};
    rc = encode_tmap_bf16(&maps.b, a->w, 3, dims, str, box);
    if (rc != PB_OK) return rc;
  }
  const size_t smem = tma_bits ? smem_tma : (size_t)p.stages * p.stage_bytes + 1024;
  typedef void (*TcKernel)(const TmapPack, const TcConvP);
  static bool attr_set = false;
  if (!attr_set) {
    const TcKernel fns[] = {tc_conv_kernel<false>, tc_conv_kernel<true>,
#define TC_EPI_FN(E) tc_conv_kernel<false, (E)>,
                            TC_EPI_LIST(TC_EPI_FN)
#undef TC_EPI_FN
    };
    for (const TcKernel f : fns) {
      cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
      if (e != cudaSuccess) return cuda_fail(e, "pb_conv_tc: smem attribute");
    }
    attr_set = true;
  }
  const int total = p.N * p.tiles_h * p.tiles_w * p.n_tiles;
  const int grid = total < sm_count() ? total : sm_count();
  TcKernel kern = a->act_dtype == PB_F16 ? tc_conv_kernel<true> : tc_conv_kernel<false>;
  {
    // compile-time specialised epilogues (bf16, NHWC, stride-1 geometries): see TC_EPI_LIST
    static int spec_on = -1;
    if (spec_on < 0) { const char* v = getenv("POSEB200_CONV_EPI_SPEC"); spec_on = (v != nullptr && v[0] == '0') ? 0 : 1; }
    if (spec_on && a->act_dtype == PB_BF16 && !p.out_nchw && !p.up && p.n_acc == 1 && a->mask_out == nullptr &&
        a->mask_in == nullptr && a->out2 == nullptr && (a->act == PB_ACT_NONE || a->act == PB_ACT_GELU)) {
      int want = TE_SPEC | (a->bias ? TE_BIAS : 0) | (a->add0 ? TE_ADD0 : 0) | (a->add1 ? TE_ADD1 : 0) |
                 (a->pre_out ? TE_PRE : 0) | (a->act == PB_ACT_GELU ? TE_GELU : 0);
      want |= tma_bits;
      switch (want) {
#define TC_EPI_CASE(E) case (E): kern = tc_conv_kernel<false, (E)>; break;
        TC_EPI_LIST(TC_EPI_CASE)
#undef TC_EPI_CASE
        default: break;
      }
    }
  }
  kern<<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(maps, p);
  PB_LAUNCH_CHECK("tc_conv_kernel");
  return PB_OK;
}

}  // extern "C"
