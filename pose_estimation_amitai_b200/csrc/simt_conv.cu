// CUDA-core gather-convolution (fp32 accumulate) and its weight gradient.
//
// This is the "fp32 mode" contraction path of the hot path (north_star: heatmaps within 1e-4 of
// the reference in fp32 mode) and the path for shapes the tcgen05 kernels do not tile
// (first layer, Cin = 4).  It is a GPU kernel, not a fallback to the CPU: the same descriptor
// (pb_conv_args) drives the tensor-core kernels in tc_conv.cu.
//
// Tiling: 64 output pixels x 64 output channels per CTA, 256 threads, 4x4 register micro-tile
// with channels strided by 16 across lanes (conflict-free shared reads, warp-ballot mask packing).
#include "common.cuh"

namespace pb {

constexpr int ST_BM = 64, ST_BN = 64, ST_BK = 16, ST_THREADS = 256;

struct TapTable {
  int ntaps, out_mul, in_div;
  int dy[PB_MAX_TAPS], dx[PB_MAX_TAPS];
};

struct ConvP {
  const void* in; const float* w; const float* bias; const void* add0; const void* add1;
  void* pre_out; void* out; uint32_t* mask_out; const uint32_t* mask_in;
  int N, IH, IW, Cin, OH, OW, Cout;
  int act; float slope; int out_nchw;
  TapTable taps;
};

template <typename T, bool IN_NCHW>
__global__ void __launch_bounds__(ST_THREADS)
conv_simt_kernel(const ConvP p) {
  __shared__ float As[ST_BK][ST_BM + 4];
  __shared__ float Bs[ST_BK][ST_BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long M = (long long)p.N * p.OH * p.OW;
  const long long m0 = (long long)blockIdx.x * ST_BM;
  const int co0 = blockIdx.y * ST_BN;

  // loader roles
  const int lp = tid >> 2;         // pixel within tile loaded by this thread (A)
  const int lq = (tid & 3) * 4;    // first of 4 channels (A)
  const int bk = tid >> 4;         // ci row (B)
  const int bc = (tid & 15) * 4;   // first of 4 co (B)
  const long long lm = m0 + lp;
  const bool lvalid = lm < M;
  int ln = 0, loy = 0, lox = 0;
  if (lvalid) {
    lox = (int)(lm % p.OW);
    loy = (int)((lm / p.OW) % p.OH);
    ln = (int)(lm / ((long long)p.OW * p.OH));
  }
  const T* in = reinterpret_cast<const T*>(p.in);
  const float* inf = reinterpret_cast<const float*>(p.in);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int t = 0; t < p.taps.ntaps; ++t) {
    // source pixel of this thread's A row for tap t
    bool ok = lvalid;
    int iy = loy * p.taps.out_mul + p.taps.dy[t];
    int ix = lox * p.taps.out_mul + p.taps.dx[t];
    if (p.taps.in_div > 1) {
      ok = ok && (iy % p.taps.in_div == 0) && (ix % p.taps.in_div == 0) && iy >= 0 && ix >= 0;
      iy /= p.taps.in_div;
      ix /= p.taps.in_div;
    }
    ok = ok && iy >= 0 && iy < p.IH && ix >= 0 && ix < p.IW;
    const long long apix = ((long long)ln * p.IH + iy) * p.IW + ix;
    const float* wt = p.w + (long long)t * p.Cin * p.Cout;
    for (int c0 = 0; c0 < p.Cin; c0 += ST_BK) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int ci = c0 + lq + e;
        float v = 0.f;
        if (ok && ci < p.Cin) {
          if (IN_NCHW) v = inf[(((long long)ln * p.Cin + ci) * p.IH + iy) * p.IW + ix];
          else v = ldf<T>(in, apix * p.Cin + ci);
        }
        As[lq + e][lp] = v;
      }
      {
        const int ci = c0 + bk;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int co = co0 + bc + e;
          Bs[bk][bc + e] = (ci < p.Cin && co < p.Cout) ? wt[(long long)ci * p.Cout + co] : 0.f;
        }
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < ST_BK; ++k) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[k][ty + 16 * i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx + 16 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // ---- epilogue: v = acc + bias + add0 ; pre_out ; act ; + add1 ; out
  const int words = (p.Cout + 31) / 32;
  const T* add0 = reinterpret_cast<const T*>(p.add0);
  const T* add1 = reinterpret_cast<const T*>(p.add1);
  T* pre_out = reinterpret_cast<T*>(p.pre_out);
  T* out = reinterpret_cast<T*>(p.out);
  float* outf = reinterpret_cast<float*>(p.out);
  const int lane = tid & 31;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty + 16 * i;
    const bool mv = m < M;
    uint32_t bal[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + tx + 16 * j;
      const bool v_ok = mv && co < p.Cout;
      float v = acc[i][j];
      const long long idx = m * p.Cout + co;
      if (v_ok) {
        if (p.bias) v += p.bias[co];
        if (add0) v += ldf<T>(add0, idx);
        if (pre_out) stf<T>(pre_out, idx, v);
      }
      const bool pos = v_ok && v > 0.f;
      bal[j] = __ballot_sync(0xffffffffu, pos);
      if (v_ok) {
        if (p.act == PB_ACT_LRELU) v = lrelu(v, p.slope);
        else if (p.act == PB_ACT_MASKMUL) {
          const uint32_t bit = (p.mask_in[m * words + (co >> 5)] >> (co & 31)) & 1u;
          v *= bit ? 1.f : p.slope;
        } else if (p.act == PB_ACT_GELU) {
          v = 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
        }
        if (add1) v += ldf<T>(add1, idx);
        if (p.out_nchw) {
          const int ox = (int)(m % p.OW);
          const int oy = (int)((m / p.OW) % p.OH);
          const long long n = m / ((long long)p.OW * p.OH);
          outf[((n * p.Cout + co) * p.OH + oy) * p.OW + ox] = v;
        } else {
          stf<T>(out, idx, v);
        }
      }
    }
    if (p.mask_out != nullptr) {
      // lanes 0-15: pixel row ty (even warp half), lanes 16-31: pixel row ty of the upper half
      const bool upper = lane >= 16;
      if ((lane & 15) == 0 && mv) {
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int wi = (co0 >> 5) + jj;
          if (wi < words) {
            const uint32_t lo = upper ? (bal[2 * jj] >> 16) : (bal[2 * jj] & 0xFFFFu);
            const uint32_t hi = upper ? (bal[2 * jj + 1] >> 16) : (bal[2 * jj + 1] & 0xFFFFu);
            p.mask_out[m * words + wi] = lo | (hi << 16);
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// weight gradient: partial[k][t][ca][cg] = sum over the CTA's pixel range of a(tap) x g(tap)
// ---------------------------------------------------------------------------------------
struct WgradP {
  const void* a; const void* g; float* partial;
  int N, PH, PW, AH, AW, Ca, GH, GW, Cg, mul_a, mul_g, ntaps, ksplit, gcs;
  int dya[PB_MAX_TAPS], dxa[PB_MAX_TAPS], dyg[PB_MAX_TAPS], dxg[PB_MAX_TAPS];
  long long px_per_split;
};

template <typename T, bool A_NCHW>
__global__ void __launch_bounds__(ST_THREADS)
wgrad_simt_kernel(const WgradP p) {
  __shared__ float As[ST_BK][ST_BM + 4];
  __shared__ float Gs[ST_BK][ST_BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int ca_tiles = (p.Ca + ST_BM - 1) / ST_BM, cg_tiles = (p.Cg + ST_BN - 1) / ST_BN;
  int bid = blockIdx.x;
  const int cgt = bid % cg_tiles; bid /= cg_tiles;
  const int cat = bid % ca_tiles; bid /= ca_tiles;
  const int t = bid;
  const int ca0 = cat * ST_BM, cg0 = cgt * ST_BN;
  const long long P = (long long)p.N * p.PH * p.PW;
  const long long q0 = (long long)blockIdx.y * p.px_per_split;
  const long long q1 = min(P, q0 + p.px_per_split);
  const T* a = reinterpret_cast<const T*>(p.a);
  const float* af = reinterpret_cast<const float*>(p.a);
  const T* g = reinterpret_cast<const T*>(p.g);
  const int lpx = tid >> 4;          // pixel row in the chunk (0..15)
  const int lch = (tid & 15) * 4;    // 4 consecutive channels

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (long long q = q0; q < q1; q += ST_BK) {
    const long long pq = q + lpx;
    bool oka = pq < q1, okg = oka;
    int n = 0, py = 0, px = 0;
    if (oka) {
      px = (int)(pq % p.PW);
      py = (int)((pq / p.PW) % p.PH);
      n = (int)(pq / ((long long)p.PW * p.PH));
    }
    const int ay = py * p.mul_a + p.dya[t], ax = px * p.mul_a + p.dxa[t];
    const int gy = py * p.mul_g + p.dyg[t], gx = px * p.mul_g + p.dxg[t];
    oka = oka && ay >= 0 && ay < p.AH && ax >= 0 && ax < p.AW;
    okg = okg && gy >= 0 && gy < p.GH && gx >= 0 && gx < p.GW;
    const bool both = oka && okg;
    const long long apix = ((long long)n * p.AH + ay) * p.AW + ax;
    const long long gpix = ((long long)n * p.GH + gy) * p.GW + gx;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int ca = ca0 + lch + e;
      float v = 0.f;
      if (both && ca < p.Ca) {
        if (A_NCHW) v = af[(((long long)n * p.Ca + ca) * p.AH + ay) * p.AW + ax];
        else v = ldf<T>(a, apix * p.Ca + ca);
      }
      As[lpx][lch + e] = v;
      const int cg = cg0 + lch + e;
      Gs[lpx][lch + e] = (both && cg < p.Cg) ? ldf<T>(g, gpix * p.gcs + cg) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < ST_BK; ++k) {
      float av[4], gv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[k][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) gv[j] = Gs[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], gv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const long long L = (long long)p.ntaps * p.Ca * p.Cg + p.Cg;
  float* dst = p.partial + (long long)blockIdx.y * L + (long long)t * p.Ca * p.Cg;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ca = ca0 + ty + 16 * i;
    if (ca >= p.Ca) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cg = cg0 + tx + 16 * j;
      if (cg < p.Cg) dst[(long long)ca * p.Cg + cg] = acc[i][j];
    }
  }
}

// bias gradient: column sums of g over all pixels (HBM-bound: g is read exactly once with 16-byte
// loads).  grid (pixel blocks, 256-channel column blocks); every block writes its per-channel partial
// to a scratch row, the LAST block to finish (ticket) folds the rows in block order -- a fixed
// summation order, so the result is deterministic -- into split 0's bias slot of `partial` and
// zeroes the other splits' slots (pb_wgrad_reduce sums the slots).  Not re-entrant across streams
// of one device (one scratch buffer); the engines issue every wgrad on a single stream.
constexpr int BIAS_SCRATCH_FLOATS = 592 * 256;
__device__ float g_bias_scratch[BIAS_SCRATCH_FLOATS];
__device__ unsigned int g_bias_ticket;

template <typename T>
__global__ void __launch_bounds__(256)
bias_partial_kernel(const T* __restrict__ g, float* __restrict__ partial, long long Pg, int Cg, int gcs, long long L,
                    long long off, int ksplit) {
  __shared__ float red[256 * 8];
  __shared__ int is_last;
  const int tid = threadIdx.x;
  const int col0 = blockIdx.y * 256;
  const long long per = (Pg + gridDim.x - 1) / gridDim.x;
  const long long q0 = (long long)blockIdx.x * per;
  const long long q1 = min(Pg, q0 + per);
  float* my_row = g_bias_scratch + (long long)blockIdx.x * Cg;
  constexpr bool kBf16 = sizeof(T) == 2;
  if (kBf16 && (gcs & 7) == 0) {
    // 8 channels (16 B) per thread, tpr threads per pixel, rpi pixels per pass
    const int cb = min(256, gcs - col0);
    const int tpr = cb >> 3;
    const int rpi = 256 / tpr;
    const int r = tid / tpr, v = tid - r * tpr;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (r < rpi) {
      const T* base = g + col0 + v * 8;
      long long q = q0 + r;
      for (; q + 3LL * rpi < q1; q += 4LL * rpi) {
        uint4 t[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) t[u] = ld_stream16(base + (q + (long long)u * rpi) * gcs);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t w[4] = {t[u].x, t[u].y, t[u].z, t[u].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            acc[2 * e] += bf16lo(w[e]);
            acc[2 * e + 1] += bf16hi(w[e]);
          }
        }
      }
      for (; q < q1; q += rpi) {
        const uint4 t = ld_stream16(base + q * gcs);
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc[2 * e] += bf16lo(w[e]);
          acc[2 * e + 1] += bf16hi(w[e]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[tid * 8 + j] = acc[j];
    __syncthreads();
    if (tid < cb && col0 + tid < Cg) {
      float s = 0.f;
      for (int rr = 0; rr < rpi; ++rr) s += red[(rr * tpr + (tid >> 3)) * 8 + (tid & 7)];
      my_row[col0 + tid] = s;
    }
  } else {
    // generic: 64 channels x 4 pixel rows per pass
    for (int cc = 0; cc < 256; cc += 64) {
      const int c = col0 + cc + (tid & 63);
      const int r = tid >> 6;
      float s = 0.f;
      if (c < Cg)
        for (long long q = q0 + r; q < q1; q += 4) s += ldf<T>(g, q * gcs + c);
      red[tid] = s;
      __syncthreads();
      if (r == 0 && c < Cg) my_row[c] = red[tid] + red[tid + 64] + red[tid + 128] + red[tid + 192];
      __syncthreads();
    }
  }
  // ---- last block folds the scratch rows
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int total = gridDim.x * gridDim.y;
    is_last = (atomicAdd(&g_bias_ticket, 1u) == total - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int c = tid; c < Cg; c += 256) {
    // four independent chains (rows b = 0, 1, 2, 3 mod 4): a fixed order, a quarter of the load latency
    float s4[4] = {0.f, 0.f, 0.f, 0.f};
    unsigned int b = 0;
    for (; b + 4 <= gridDim.x; b += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) s4[u] += __ldcg(g_bias_scratch + (long long)(b + u) * Cg + c);
    }
    for (; b < gridDim.x; ++b) s4[0] += __ldcg(g_bias_scratch + (long long)b * Cg + c);
    const float s = (s4[0] + s4[1]) + (s4[2] + s4[3]);
    partial[off + c] = s;
    for (int k = 1; k < ksplit; ++k) partial[(long long)k * L + off + c] = 0.f;
  }
  if (tid == 0) g_bias_ticket = 0u;
}

int launch_bias_partial(const pb_wgrad_args* a, cudaStream_t st) {
  const long long Pg = (long long)a->N * a->GH * a->GW;
  const long long L = (long long)a->ntaps * a->Ca * a->Cg + a->Cg;
  const int gcs = a->g_cstride ? a->g_cstride : a->Cg;
  if (a->Cg > BIAS_SCRATCH_FLOATS) {
    set_error("bias gradient: Cg=%d exceeds the scratch row", a->Cg);
    return PB_ERR_UNSUPPORTED;
  }
  const int colblocks = cdiv(a->Cg, 256);
  long long nblk = (4LL * sm_count()) / colblocks;          // ~4 blocks per SM in total
  nblk = nblk < BIAS_SCRATCH_FLOATS / a->Cg ? nblk : BIAS_SCRATCH_FLOATS / a->Cg;
  const long long by_work = (Pg + 255) / 256;               // at least 256 pixels per block: the last block's fold over
                                                            // the scratch rows is the latency chain of small problems
  if (nblk > by_work) nblk = by_work;
  if (nblk < 1) nblk = 1;
  dim3 g2((unsigned)nblk, (unsigned)colblocks);
  if (a->act_dtype == PB_BF16)
    bias_partial_kernel<__nv_bfloat16><<<g2, 256, 0, st>>>((const __nv_bfloat16*)a->g, a->partial, Pg, a->Cg, gcs, L,
                                                         L - a->Cg, a->ksplit);
  else
    bias_partial_kernel<float><<<g2, 256, 0, st>>>((const float*)a->g, a->partial, Pg, a->Cg, gcs, L, L - a->Cg,
                                                  a->ksplit);
  PB_LAUNCH_CHECK("bias_partial_kernel");
  return PB_OK;
}

int conv_args_check(const pb_conv_args* a, const char* fn) {
  if (a == nullptr || a->in == nullptr || a->w == nullptr || a->out == nullptr) {
    set_error("%s: null in/w/out", fn);
    return PB_ERR_INVALID;
  }
  if (a->N <= 0 || a->IH <= 0 || a->IW <= 0 || a->Cin <= 0 || a->OH <= 0 || a->OW <= 0 || a->Cout <= 0) {
    set_error("%s: empty shape", fn);
    return PB_ERR_INVALID;
  }
  if (a->taps.ntaps < 1 || a->taps.ntaps > PB_MAX_TAPS || a->taps.out_mul < 1 || a->taps.in_div < 1) {
    set_error("%s: bad tap table", fn);
    return PB_ERR_INVALID;
  }
  if (a->act == PB_ACT_MASKMUL && a->mask_in == nullptr) {
    set_error("%s: PB_ACT_MASKMUL needs mask_in", fn);
    return PB_ERR_INVALID;
  }
  const void* ptrs[] = {a->in, a->w, a->bias, a->add0, a->add1, a->pre_out, a->out, a->out2, a->mask_out, a->mask_in};
  for (const void* q : ptrs)
    if (q != nullptr && !is_device_ptr(q)) {
      set_error("%s: host pointer passed (no CPU fallback)", fn);
      return PB_ERR_NOT_DEVICE;
    }
  return PB_OK;
}

}  // namespace pb

using namespace pb;

extern "C" {

int pb_conv_simt(const pb_conv_args* a, void* stream) {
  int rc = conv_args_check(a, "pb_conv_simt");
  if (rc != PB_OK) return rc;
  ConvP p;
  p.in = a->in; p.w = (const float*)a->w; p.bias = a->bias; p.add0 = a->add0; p.add1 = a->add1;
  p.pre_out = a->pre_out; p.out = a->out; p.mask_out = a->mask_out; p.mask_in = a->mask_in;
  p.N = a->N; p.IH = a->IH; p.IW = a->IW; p.Cin = a->Cin; p.OH = a->OH; p.OW = a->OW; p.Cout = a->Cout;
  p.act = a->act; p.slope = a->slope; p.out_nchw = a->out_nchw_f32;
  p.taps.ntaps = a->taps.ntaps; p.taps.out_mul = a->taps.out_mul; p.taps.in_div = a->taps.in_div;
  for (int t = 0; t < PB_MAX_TAPS; ++t) { p.taps.dy[t] = a->taps.dy[t]; p.taps.dx[t] = a->taps.dx[t]; }
  const long long M = (long long)a->N * a->OH * a->OW;
  dim3 grid(cdiv(M, ST_BM), cdiv(a->Cout, ST_BN));
  cudaStream_t st = (cudaStream_t)stream;
  if (a->in_nchw_f32) {
    if (a->act_dtype == PB_BF16) conv_simt_kernel<__nv_bfloat16, true><<<grid, ST_THREADS, 0, st>>>(p);
    else conv_simt_kernel<float, true><<<grid, ST_THREADS, 0, st>>>(p);
  } else {
    if (a->act_dtype == PB_BF16) conv_simt_kernel<__nv_bfloat16, false><<<grid, ST_THREADS, 0, st>>>(p);
    else conv_simt_kernel<float, false><<<grid, ST_THREADS, 0, st>>>(p);
  }
  PB_LAUNCH_CHECK("conv_simt_kernel");
  return PB_OK;
}

int pb_wgrad_simt(const pb_wgrad_args* a, void* stream) {
  PB_REQUIRE(a != nullptr && a->a && a->g && a->partial, "pb_wgrad_simt: null args");
  PB_REQUIRE(a->N > 0 && a->PH > 0 && a->PW > 0 && a->Ca > 0 && a->Cg > 0 && a->ksplit >= 1 && a->ntaps >= 1 &&
                 a->ntaps <= PB_MAX_TAPS,
             "pb_wgrad_simt: bad shape");
  PB_REQUIRE_DEV(a->a, "a");
  PB_REQUIRE_DEV(a->g, "g");
  PB_REQUIRE_DEV(a->partial, "partial");
  WgradP p;
  p.a = a->a; p.g = a->g; p.partial = a->partial;
  p.N = a->N; p.PH = a->PH; p.PW = a->PW; p.AH = a->AH; p.AW = a->AW; p.Ca = a->Ca; p.GH = a->GH; p.GW = a->GW;
  p.Cg = a->Cg; p.gcs = a->g_cstride ? a->g_cstride : a->Cg; p.mul_a = a->mul_a; p.mul_g = a->mul_g; p.ntaps = a->ntaps; p.ksplit = a->ksplit;
  for (int t = 0; t < PB_MAX_TAPS; ++t) {
    p.dya[t] = a->dya[t]; p.dxa[t] = a->dxa[t]; p.dyg[t] = a->dyg[t]; p.dxg[t] = a->dxg[t];
  }
  const long long P = (long long)a->N * a->PH * a->PW;
  long long per = (P + a->ksplit - 1) / a->ksplit;
  per = ((per + ST_BK - 1) / ST_BK) * ST_BK;
  p.px_per_split = per;
  const int ca_tiles = cdiv(a->Ca, ST_BM), cg_tiles = cdiv(a->Cg, ST_BN);
  dim3 grid(a->ntaps * ca_tiles * cg_tiles, a->ksplit);
  cudaStream_t st = (cudaStream_t)stream;
  if (a->a_nchw_f32) {
    if (a->act_dtype == PB_BF16) wgrad_simt_kernel<__nv_bfloat16, true><<<grid, ST_THREADS, 0, st>>>(p);
    else wgrad_simt_kernel<float, true><<<grid, ST_THREADS, 0, st>>>(p);
  } else {
    if (a->act_dtype == PB_BF16) wgrad_simt_kernel<__nv_bfloat16, false><<<grid, ST_THREADS, 0, st>>>(p);
    else wgrad_simt_kernel<float, false><<<grid, ST_THREADS, 0, st>>>(p);
  }
  PB_LAUNCH_CHECK("wgrad_simt_kernel");
  if (a->want_bias) return launch_bias_partial(a, st);
  return PB_OK;
}

}  // extern "C"
