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

// bias gradient partials: column sums of g over a pixel range.  grid (ceil(Cg/64), ksplit), 256 threads
template <typename T>
__global__ void __launch_bounds__(256)
bias_partial_kernel(const T* __restrict__ g, float* __restrict__ partial, long long Pg, int Cg, int gcs, long long L,
                    long long off, long long px_per_split) {
  __shared__ float red[4][64];
  const int c = blockIdx.x * 64 + (threadIdx.x & 63);
  const int r = threadIdx.x >> 6;
  const long long q0 = (long long)blockIdx.y * px_per_split;
  const long long q1 = min(Pg, q0 + px_per_split);
  float s = 0.f;
  if (c < Cg)
    for (long long q = q0 + r; q < q1; q += 4) s += ldf<T>(g, q * gcs + c);
  red[r][threadIdx.x & 63] = s;
  __syncthreads();
  if (r == 0 && c < Cg)
    partial[(long long)blockIdx.y * L + off + c] = red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] +
                                                   red[3][threadIdx.x];
}

int launch_bias_partial(const pb_wgrad_args* a, cudaStream_t st) {
  const long long Pg = (long long)a->N * a->GH * a->GW;
  const long long L = (long long)a->ntaps * a->Ca * a->Cg + a->Cg;
  const long long perg = (Pg + a->ksplit - 1) / a->ksplit;
  dim3 g2(cdiv(a->Cg, 64), a->ksplit);
  if (a->act_dtype == PB_BF16)
    bias_partial_kernel<__nv_bfloat16><<<g2, 256, 0, st>>>((const __nv_bfloat16*)a->g, a->partial, Pg, a->Cg, (a->g_cstride ? a->g_cstride : a->Cg), L,
                                                         L - a->Cg, perg);
  else
    bias_partial_kernel<float><<<g2, 256, 0, st>>>((const float*)a->g, a->partial, Pg, a->Cg, (a->g_cstride ? a->g_cstride : a->Cg), L, L - a->Cg, perg);
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
  const void* ptrs[] = {a->in, a->w, a->bias, a->add0, a->add1, a->pre_out, a->out, a->mask_out, a->mask_in};
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
