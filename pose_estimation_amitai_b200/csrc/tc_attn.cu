// Fused tcgen05 attention for the ViT encoder (pytorch_vit_encoder.py:59-78 and its autograd): ONE CTA owns one
// (sample, head) problem and runs TWO chained products on operands it loads once.
//
//   forward     S = Q K^T -> row softmax -> P (bf16: to HBM for the backward AND into shared memory) ; O = P V
//   backward A  dP = dO V^T -> dS = P o (dP - rowsum(dP o P)) * scale (bf16 to HBM) ;  dV = P^T dO
//   backward B  dQ = dS K ;  dK = dS^T Q
//
// against six single-product launches of tc_bgemm.cu, each of which re-loaded both of its operands and left
// through HBM: three launches, q / k / v / dO / P / dS each loaded once per launch that uses them.
//
// Canonical operand image.  Every matrix X[token][feature] (Q, K, V, dO: 144 x 256; P, dS: 144 x 144) sits in
// shared memory as [64-feature chunk][token row x 128 B], SWIZZLE_128B -- one TMA box {64, S} per chunk.  The same
// bytes serve BOTH operand roles: "K-major" when the product contracts over features (S = Q K^T, dP = dO V^T,
// O = P V and dQ = dS K for the P / dS side), "MN-major" when it contracts over tokens (V in P V, dO and P in
// P^T dO, K in dS K, dS and Q in dS^T Q): rows are then the K index (16 rows = 2048 B per K step) and consecutive
// 64-element M/N blocks are one chunk (S * 128 B) apart -- the descriptor's leading byte offset.
//
// M = S = 144 tokens is two 128-row tiles; the second one's operand rows 144..255 are whatever follows in shared
// memory (rows of an MMA are independent; those accumulator rows are never read).  TMEM: product 1 uses columns
// [0, 2*N1), product 2 re-uses [0, 2*N2) after the epilogue of product 1 has drained (one __syncthreads).
// The kernel is persistent (one CTA per SM) and requests the next problem's images behind the last epilogue.
#include <string.h>

#include "tc_common.cuh"

namespace pb {

using namespace tc;

struct AtMaps {
  CUtensorMap x[3];
};

struct AtGemm {
  int a_img, b_img, a_mn, b_mn;   // operand images and roles
  int n, ksteps;                  // output columns (multiple of 16), K / 16
  int epi;                        // 0 store alpha*acc | 1 row softmax of alpha*acc | 2 softmax backward (needs P)
  void* C;                        // bf16 output, element (zb, zh, row, col) at C + zb*c_zb + zh*c_zh + row*c_m + col
  long long c_zb, c_zh, c_m;
  float alpha;
};

struct AtP {
  int S, ZH;
  int chunks[3];                  // 64-feature chunks per image
  uint32_t CH, IMG;               // bytes per chunk (S * 128), bytes per image slot
  AtGemm g[2];
  const __nv_bfloat16* P;         // epi 2: probabilities, indexed like g[0].C
  int p_to_img0;                  // forward: the softmax epilogue also writes P as image 0 (over Q)
  int early_img1;                 // request the next problem's image 1 when product 1 retires (POSEB200_ATTN_EARLY=0: off)
};

constexpr int AT_THREADS = 256;

// 32 bytes per lane and instruction: a row-per-thread epilogue writes at a row-pitch lane stride, where a 16-byte store
// fills each 32-byte sector in two partial writes and costs the same 32 LSU wavefronts
__device__ __forceinline__ void st_global_256(void* dst, const uint4& lo, const uint4& hi) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(lo.x), "r"(lo.y), "r"(lo.z),
               "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w)
               : "memory");
}

__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void at_issue(const AtP& p, const AtGemm& g, uint32_t img_base, uint32_t tmem_base, int ntile,
                                         uint64_t* bars) {
  const uint32_t idesc = make_idesc(128, g.n, g.a_mn, g.b_mn);
  const uint32_t a0 = img_base + (uint32_t)g.a_img * p.IMG, b0 = img_base + (uint32_t)g.b_img * p.IMG;
  for (int tile = 0; tile < ntile; ++tile) {
    for (int j = 0; j < g.ksteps; ++j) {
      const uint32_t kmaj = (uint32_t)(j >> 2) * p.CH + (uint32_t)(j & 3) * 32u;   // feature chunk j/4, 16 features
      const uint32_t mnmaj = (uint32_t)j * 2048u;                                   // 16 token rows
      const uint64_t ad = g.a_mn ? smem_desc_sw128(a0 + (uint32_t)(tile * 2) * p.CH + mnmaj, p.CH, 1024)
                                 : smem_desc_sw128(a0 + kmaj + (uint32_t)tile * 16384u, 16, 1024);
      const uint64_t bd = g.b_mn ? smem_desc_sw128(b0 + mnmaj, p.CH, 1024) : smem_desc_sw128(b0 + kmaj, 16, 1024);
      umma_bf16(tmem_base + (uint32_t)(tile * g.n), ad, bd, idesc, j > 0 ? 1u : 0u);
    }
    umma_commit(&bars[tile]);   // tile 0's epilogue starts while tile 1's MMAs run
  }
}

// Epilogue of one product.  A thread owns one accumulator row and HALF of its columns: warps 0-3 (TMEM lane quadrant
// = warp) take the low 16-column chunks, warps 4-7 (quadrant = warp - 4) the high ones, so all eight warps work on the
// 128 rows of tile 0 (tile 1 of a 144-token problem is 16 rows: one more pass by warps 0 and 4).  The row is read
// from TMEM once and kept in registers; softmax's max / sum and the softmax-backward dot product are combined across
// the two halves through `red`.  Every warp executes every __syncthreads (inactive ones just skip the work).
constexpr int AT_MAXC = 6;   // 16-column chunks per half held in registers: N = S <= 192 (host check)

// `early`: called once by warp 1's elected lane when the last tile's MMAs have retired and warp 1 has no rows in that
// tile (S = 144: tile 1 is 16 rows) -- the kernel uses it to request the NEXT problem's image 1, which only product 1
// reads, behind this epilogue and the whole of product 2.
template <int EPI, class Early>
__device__ __forceinline__ void at_epilogue_rows(const AtP& p, const AtGemm& g, uint8_t* img0, uint32_t tmem_base, int ntile,
                                            uint64_t* bars, uint32_t ph, int zb, int zh, bool p_to_img,
                                            float (*red)[2][128], const bool early_on, Early&& early) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = warp >> 2, q = warp & 3;
  const int nchunks = g.n >> 4;
  const int c_lo = half == 0 ? 0 : (nchunks + 1) >> 1;
  const int c_hi = half == 0 ? (nchunks + 1) >> 1 : nchunks;
  const int rit = q * 32 + lane;                        // row inside the tile
  for (int tile = 0; tile < ntile; ++tile) {
    const bool active = tile * 128 + q * 32 < p.S;      // warp-uniform
    const int row = tile * 128 + rit;
    const bool row_ok = row < p.S;
    const long long coff = (long long)zb * g.c_zb + (long long)zh * g.c_zh + (long long)row * g.c_m;
    uint32_t xr[AT_MAXC][16];
    if (early_on && tile == ntile - 1 && warp == 1) {   // warp 1 is idle in this tile (early_on implies it)
      if (elect_one()) {
        mbar_wait(&bars[tile], ph);
        early();
      }
      __syncwarp();
    }
    if (active) {
      mbar_wait(&bars[tile], ph);
      tc_fence_after();
      const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tile * g.n);
#pragma unroll
      for (int k = 0; k < AT_MAXC; ++k)
        if (c_lo + k < c_hi) tmem_ld16(lane_base + (uint32_t)((c_lo + k) * 16), xr[k]);
      tmem_ld_wait();
    }
    float r_max = 0.f, inv_sum = 0.f, r_dot = 0.f;
    if (EPI == 1) {
      float m = -INFINITY;
      if (active) {
#pragma unroll
        for (int k = 0; k < AT_MAXC; ++k)
          if (c_lo + k < c_hi) {
#pragma unroll
            for (int j = 0; j < 16; ++j) m = fmaxf(m, g.alpha * __uint_as_float(xr[k][j]));
          }
        red[0][half][rit] = m;
      }
      __syncthreads();
      float sum = 0.f;
      if (active) {
        r_max = fmaxf(red[0][0][rit], red[0][1][rit]);
#pragma unroll
        for (int k = 0; k < AT_MAXC; ++k)
          if (c_lo + k < c_hi) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float e = expf(g.alpha * __uint_as_float(xr[k][j]) - r_max);
              xr[k][j] = __float_as_uint(e);
              sum += e;
            }
          }
        red[1][half][rit] = sum;
      }
      __syncthreads();
      if (active) inv_sum = 1.f / (red[1][0][rit] + red[1][1][rit]);
    } else {
      float d = 0.f;
      if (active && row_ok) {
#pragma unroll
        for (int k = 0; k < AT_MAXC; ++k)
          if (c_lo + k < c_hi) {
            const uint4* pp = reinterpret_cast<const uint4*>(p.P + coff + (c_lo + k) * 16);
            const uint4 q0 = __ldg(pp), q1 = __ldg(pp + 1);
            const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              d += __uint_as_float(xr[k][2 * j]) * bf16lo(w[j]);
              d += __uint_as_float(xr[k][2 * j + 1]) * bf16hi(w[j]);
            }
          }
      }
      if (active) red[0][half][rit] = d;
      __syncthreads();
      if (active) r_dot = red[0][0][rit] + red[0][1][rit];
    }
    if (active && p_to_img && ntile == 2) {
      // P overwrites Q: every MMA of product 1 (both tiles) must have finished reading Q first
      mbar_wait(&bars[1], ph);
      tc_fence_after();
    }
    if (active && row_ok) {
#pragma unroll
      for (int k = 0; k < AT_MAXC; ++k)
        if (c_lo + k < c_hi) {
          const int c = c_lo + k;
          float v[16];
          if (EPI == 2) {
            const uint4* pp = reinterpret_cast<const uint4*>(p.P + coff + c * 16);
            const uint4 q0 = __ldg(pp), q1 = __ldg(pp + 1);
            const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              v[2 * j] = bf16lo(w[j]) * (__uint_as_float(xr[k][2 * j]) - r_dot) * g.alpha;
              v[2 * j + 1] = bf16hi(w[j]) * (__uint_as_float(xr[k][2 * j + 1]) - r_dot) * g.alpha;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(xr[k][j]) * inv_sum;   // xr holds exp(alpha*acc - max)
          }
          uint4 t0, t1;
          t0.x = pack_bf16x2(v[0], v[1]); t0.y = pack_bf16x2(v[2], v[3]);
          t0.z = pack_bf16x2(v[4], v[5]); t0.w = pack_bf16x2(v[6], v[7]);
          t1.x = pack_bf16x2(v[8], v[9]); t1.y = pack_bf16x2(v[10], v[11]);
          t1.z = pack_bf16x2(v[12], v[13]); t1.w = pack_bf16x2(v[14], v[15]);
          st_global_256(reinterpret_cast<__nv_bfloat16*>(g.C) + coff + c * 16, t0, t1);
          if (p_to_img) {
            // columns c*16 .. c*16+15 of row `row` in the canonical image: chunk (c/4), 16-byte units u, u+1 of the
            // 128-byte row, XOR-swizzled with the row's position in its 8-row group (chunk bases are 1024-aligned)
            uint8_t* rowp = img0 + (size_t)(c >> 2) * p.CH + (size_t)row * 128;
            const int u = (c & 3) * 2, sw = row & 7;
            *reinterpret_cast<uint4*>(rowp + ((u ^ sw) << 4)) = t0;
            *reinterpret_cast<uint4*>(rowp + (((u + 1) ^ sw) << 4)) = t1;
          }
        }
    }
    if (tile + 1 < ntile) __syncthreads();   // `red` is re-used by the next tile
  }
}

// Plain store epilogue (alpha * acc -> bf16): same row / column-half ownership, 32 columns at a time, no exchange.
template <class Early>
__device__ __forceinline__ void at_epilogue_plain(const AtP& p, const AtGemm& g, uint32_t tmem_base, int ntile,
                                                  uint64_t* bars, uint32_t ph, int zb, int zh, const bool early_on,
                                                  Early&& early) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = warp >> 2, q = warp & 3;
  const int nchunks = g.n >> 4;
  const int c_lo = half == 0 ? 0 : (nchunks + 1) >> 1;
  const int c_hi = half == 0 ? (nchunks + 1) >> 1 : nchunks;
  for (int tile = 0; tile < ntile; ++tile) {
    if (early_on && tile == ntile - 1 && warp == 1) {   // see at_epilogue_rows
      if (elect_one()) {
        mbar_wait(&bars[tile], ph);
        early();
      }
      __syncwarp();
    }
    if (tile * 128 + q * 32 >= p.S) continue;           // warp-uniform
    const int row = tile * 128 + q * 32 + lane;
    const bool row_ok = row < p.S;
    __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(g.C) + (long long)zb * g.c_zb + (long long)zh * g.c_zh +
                          (long long)row * g.c_m;
    mbar_wait(&bars[tile], ph);
    tc_fence_after();
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tile * g.n);
    for (int c = c_lo; c < c_hi; c += 2) {
      uint32_t r0[16], r1[16];
      const bool two = c + 1 < c_hi;
      tmem_ld16(lane_base + (uint32_t)(c * 16), r0);
      if (two) tmem_ld16(lane_base + (uint32_t)((c + 1) * 16), r1);
      tmem_ld_wait();
      if (!row_ok) continue;
      uint4 t[4];
      uint32_t* tw = reinterpret_cast<uint32_t*>(t);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        tw[j] = pack_bf16x2(g.alpha * __uint_as_float(r0[2 * j]), g.alpha * __uint_as_float(r0[2 * j + 1]));
        tw[8 + j] = pack_bf16x2(g.alpha * __uint_as_float(r1[2 * j]), g.alpha * __uint_as_float(r1[2 * j + 1]));
      }
      st_global_256(crow + c * 16, t[0], t[1]);
      if (two) st_global_256(crow + (c + 1) * 16, t[2], t[3]);
    }
  }
}

// Persistent: one CTA per SM walks the (sample, head) problems z = blockIdx.x, + gridDim.x, ...; barriers and the
// TMEM allocation are set up once, every barrier completes exactly one phase per problem (parity = iteration & 1),
// and the next problem's operand images are requested as soon as product 2's MMAs have retired -- they land behind
// its epilogue.
template <int EPI0>
__global__ void __launch_bounds__(AT_THREADS, 1)
tc_attn_kernel(const __grid_constant__ AtMaps maps, const AtP p, const int Z) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t load_bar[3], mma_bar[4];   // load_bar: image 0 / image 2 / image 1
  __shared__ uint32_t tmem_slot;
  __shared__ float red[2][2][128];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5;
  const int ntile = p.S > 128 ? 2 : 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) prefetch_tmap(&maps.x[i]);
    for (int i = 0; i < 3; ++i) mbar_init(&load_bar[i], 1);
    for (int i = 0; i < 4; ++i) mbar_init(&mma_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t img_base = smem_u32(smem);

  // images 0 and 1 feed product 1; image 2 is only needed by product 2 and lands behind product 1.  Image 1 (K / V / K)
  // is read by product 1 ALONE in all three launches: with an idle warp in the last tile (S = 144) the next problem's
  // image 1 is requested as soon as product 1's MMAs have retired, the other two when product 2's have.
  const bool early_on = ntile == 2 && p.S - 128 <= 32 && p.early_img1 != 0;
  auto request_image = [&](int z, int i, uint64_t* bar) {
    const int zb = z / p.ZH, zh = z - zb * p.ZH;
    mbar_expect_tx(bar, (uint32_t)p.chunks[i] * p.CH);
    for (int c = 0; c < p.chunks[i]; ++c)
      tma_load_4d(smem + (size_t)i * p.IMG + (size_t)c * p.CH, &maps.x[i], bar, c * 64, 0, zh, zb);
  };
  auto request_images = [&](int z, bool with_img1) {
    request_image(z, 0, &load_bar[0]);
    if (with_img1) request_image(z, 1, &load_bar[2]);
    request_image(z, 2, &load_bar[1]);
  };
  if (warp == 0 && elect_one() && (int)blockIdx.x < Z) request_images(blockIdx.x, true);
  __syncwarp();

  uint32_t it = 0;
  for (int z = blockIdx.x; z < Z; z += gridDim.x, ++it) {
    const uint32_t ph = it & 1u;
    const int zb = z / p.ZH, zh = z - zb * p.ZH;
    if (warp == 0 && elect_one()) {
      mbar_wait(&load_bar[0], ph);
      mbar_wait(&load_bar[2], ph);
      if (p.g[0].a_img == 2 || p.g[0].b_img == 2) mbar_wait(&load_bar[1], ph);
      tc_fence_after();
      at_issue(p, p.g[0], img_base, tmem_base, ntile, &mma_bar[0]);
    }
    __syncwarp();
    const int z_next = z + (int)gridDim.x;
    auto early = [&]() { if (z_next < Z) request_image(z_next, 1, &load_bar[2]); };
    if (EPI0 == 0) at_epilogue_plain(p, p.g[0], tmem_base, ntile, &mma_bar[0], ph, zb, zh, early_on, early);
    else at_epilogue_rows<EPI0 == 0 ? 1 : EPI0>(p, p.g[0], smem, tmem_base, ntile, &mma_bar[0], ph, zb, zh,
                                                 p.p_to_img0 != 0, red, early_on, early);

    // product 2 re-uses the accumulator columns (and, in the forward, reads the P image the epilogue just wrote
    // with ordinary stores: make them visible to the tensor core's async proxy)
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      if (elect_one()) {
        tc_fence_after();
        mbar_wait(&load_bar[1], ph);
        tc_fence_after();
        at_issue(p, p.g[1], img_base, tmem_base, ntile, &mma_bar[2]);
      }
      __syncwarp();
      // all three images are free once product 2's MMAs have retired: fetch the next problem behind this epilogue
      for (int t = 0; t < ntile; ++t) mbar_wait(&mma_bar[2 + t], ph);
      if (elect_one() && z_next < Z) request_images(z_next, !early_on);
      __syncwarp();
    }
    at_epilogue_plain(p, p.g[1], tmem_base, ntile, &mma_bar[2], ph, zb, zh, false, [] {});

    tc_fence_before();
    __syncthreads();      // accumulators drained, `red` free: the next problem may overwrite both
    tc_fence_after();
  }

  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ----------------------------------------------------------------------------- host side
struct AtOperand {
  const void* ptr;                // element (zb, zh, token, feature) at ptr + zb*s_zb + zh*s_zh + token*ld + feature
  long long s_zb, s_zh, ld;
  int features;
};

static bool out_aligned32(const AtGemm& g) {
  return (((uintptr_t)g.C) & 31) == 0 && (g.c_zb & 15) == 0 && (g.c_zh & 15) == 0 && (g.c_m & 15) == 0;
}

static int encode_image(CUtensorMap* map, const AtOperand& o, int S, int ZH, int ZB) {
  if ((o.ld & 7) || (o.s_zh & 7) || (o.s_zb & 7) || (((uintptr_t)o.ptr) & 15)) return PB_ERR_UNSUPPORTED;
  const uint64_t dims[4] = {(uint64_t)o.features, (uint64_t)S, (uint64_t)ZH, (uint64_t)ZB};
  const uint64_t str[3] = {(uint64_t)o.ld * 2, (uint64_t)o.s_zh * 2, (uint64_t)o.s_zb * 2};
  const uint32_t box[4] = {64, (uint32_t)S, 1, 1};
  return encode_tmap_bf16(map, o.ptr, 4, dims, str, box);
}

static bool attn_fused_shape_ok(int S, int D) {
  const char* off = getenv("POSEB200_ATTN_UNFUSED");
  if (off != nullptr && off[0] == '1') return false;
  // two 128-row tiles; 16-token K steps; whole 64-feature chunks; three image slots within 227 KB
  if (S < 16 || S > 256 || (S & 15) != 0 || D < 64 || D > 256 || (D & 63) != 0) return false;
  const size_t img = (size_t)(D / 64 > (S + 63) / 64 ? D / 64 : (S + 63) / 64) * S * 128;
  if ((S / 16 + 1) / 2 > AT_MAXC) return false;                    // a softmax half-row lives in registers
  return 3 * img + 1024 <= (size_t)227 * 1024 - 4096;              // + ~2 KB of static shared memory
}

static int attn_launch(AtP& p, const AtOperand (&ops)[3], int S, int D, int ZH, int ZB, cudaStream_t st) {
  AtMaps maps;
  memset(&maps, 0, sizeof(maps));
  p.S = S;
  p.ZH = ZH;
  p.CH = (uint32_t)S * 128u;
  const int max_chunks = D / 64 > (S + 63) / 64 ? D / 64 : (S + 63) / 64;
  p.IMG = (uint32_t)max_chunks * p.CH;
  for (int i = 0; i < 3; ++i) {
    p.chunks[i] = (ops[i].features + 63) / 64;
    const int rc = encode_image(&maps.x[i], ops[i], S, ZH, ZB);
    if (rc != PB_OK) return rc;
  }
  const size_t smem = (size_t)3 * p.IMG + 1024;
  {
    static int early = -1;
    if (early < 0) { const char* v = getenv("POSEB200_ATTN_EARLY"); early = (v != nullptr && v[0] == '0') ? 0 : 1; }
    // image 1 must be product 1's alone
    p.early_img1 = early && p.g[1].a_img != 1 && p.g[1].b_img != 1 && (p.g[0].a_img == 1 || p.g[0].b_img == 1);
  }
  if (p.g[1].epi != 0) return PB_ERR_INVALID;
  if (!out_aligned32(p.g[0]) || !out_aligned32(p.g[1])) return PB_ERR_UNSUPPORTED;   // 256-bit row stores
  void (*kern)(const AtMaps, const AtP, const int) =
      p.g[0].epi == 1 ? tc_attn_kernel<1> : p.g[0].epi == 2 ? tc_attn_kernel<2> : tc_attn_kernel<0>;
  static size_t attr[3] = {0, 0, 0};
  if (smem > attr[p.g[0].epi]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "tc_attn: smem attribute");
    attr[p.g[0].epi] = smem;
  }
  const int Z = ZH * ZB;
  kern<<<Z < sm_count() ? Z : sm_count(), AT_THREADS, smem, st>>>(maps, p, Z);
  PB_LAUNCH_CHECK("tc_attn_kernel");
  return PB_OK;
}

static AtGemm at_gemm(int a_img, int a_mn, int b_img, int b_mn, int n, int k, int epi, void* C, long long c_zb,
                      long long c_zh, long long c_m, float alpha) {
  AtGemm g;
  memset(&g, 0, sizeof(g));
  g.a_img = a_img; g.a_mn = a_mn; g.b_img = b_img; g.b_mn = b_mn;
  g.n = n; g.ksteps = k / 16; g.epi = epi;
  g.C = C; g.c_zb = c_zb; g.c_zh = c_zh; g.c_m = c_m; g.alpha = alpha;
  return g;
}

// softmax(scale * Q K^T) V; probs (bf16 [B][H][S][S]) kept for the backward.  PB_ERR_UNSUPPORTED: use tc_bgemm.
int attn_fwd_fused(const void* qkv_, void* probs_, void* out, int B, int S, int H, int D, float scale, cudaStream_t st) {
  if (!attn_fused_shape_ok(S, D)) return PB_ERR_UNSUPPORTED;
  const __nv_bfloat16* qkv = (const __nv_bfloat16*)qkv_;
  const long long HD = (long long)H * D, qkv_b = (long long)S * 3 * HD;
  const long long pr_b = (long long)H * S * S, pr_h = (long long)S * S;
  AtP p;
  memset(&p, 0, sizeof(p));
  const AtOperand ops[3] = {{qkv, qkv_b, D, 3 * HD, D}, {qkv + HD, qkv_b, D, 3 * HD, D}, {qkv + 2 * HD, qkv_b, D, 3 * HD, D}};
  p.g[0] = at_gemm(0, 0, 1, 0, S, D, 1, probs_, pr_b, pr_h, S, scale);                       // P = softmax(Q K^T)
  p.g[1] = at_gemm(0, 0, 2, 1, D, S, 0, out, (long long)S * HD, D, HD, 1.f);                 // O = P V (P: image 0)
  p.p_to_img0 = 1;
  return attn_launch(p, ops, S, D, H, B, st);
}

int attn_bwd_fused(const void* qkv_, const void* probs_, const void* gout_, void* gqkv_, void* ds_, int B, int S, int H,
                   int D, float scale, cudaStream_t st) {
  if (!attn_fused_shape_ok(S, D)) return PB_ERR_UNSUPPORTED;
  const __nv_bfloat16* qkv = (const __nv_bfloat16*)qkv_;
  const __nv_bfloat16* go = (const __nv_bfloat16*)gout_;
  const __nv_bfloat16* probs = (const __nv_bfloat16*)probs_;
  __nv_bfloat16* gqkv = (__nv_bfloat16*)gqkv_;
  __nv_bfloat16* ds = (__nv_bfloat16*)ds_;
  const long long HD = (long long)H * D, qkv_b = (long long)S * 3 * HD, go_b = (long long)S * HD;
  const long long pr_b = (long long)H * S * S, pr_h = (long long)S * S;
  {
    // A: dS = softmax'(dO V^T) * scale ; dV = P^T dO            images: dO, V, P
    AtP p;
    memset(&p, 0, sizeof(p));
    const AtOperand ops[3] = {{go, go_b, D, HD, D}, {qkv + 2 * HD, qkv_b, D, 3 * HD, D}, {probs, pr_b, pr_h, S, S}};
    p.g[0] = at_gemm(0, 0, 1, 0, S, D, 2, ds, pr_b, pr_h, S, scale);
    p.g[1] = at_gemm(2, 1, 0, 1, D, S, 0, gqkv + 2 * HD, qkv_b, D, 3 * HD, 1.f);
    p.P = probs;
    const int rc = attn_launch(p, ops, S, D, H, B, st);
    if (rc != PB_OK) return rc;
  }
  // B: dQ = dS K ; dK = dS^T Q                                   images: dS, K, Q
  AtP p;
  memset(&p, 0, sizeof(p));
  const AtOperand ops[3] = {{ds, pr_b, pr_h, S, S}, {qkv + HD, qkv_b, D, 3 * HD, D}, {qkv, qkv_b, D, 3 * HD, D}};
  p.g[0] = at_gemm(0, 0, 1, 1, D, S, 0, gqkv, qkv_b, D, 3 * HD, 1.f);
  p.g[1] = at_gemm(0, 1, 2, 1, D, S, 0, gqkv + HD, qkv_b, D, 3 * HD, 1.f);
  return attn_launch(p, ops, S, D, H, B, st);
}

}  // namespace pb
