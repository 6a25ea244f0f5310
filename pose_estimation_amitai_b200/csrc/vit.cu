// ViT-encoder pieces of the hot path (pytorch/pytorch_vit_encoder.py, pytorch/VITs.py):
// patchify, LayerNorm fwd/bwd, softmax attention fwd/bwd (strided batched GEMMs + row softmax),
// GELU backward, and the batch-global min/max normalisation of CNN_Decoder (VITs.py:55-58).
// The Linear layers run through the gather-convolution kernels (1 tap, H = 1).
#include <math.h>

#include <stdlib.h>

#include "common.cuh"

namespace pb {

// ------------------------------------------------------------------------------ patchify
template <typename T>
__global__ void __launch_bounds__(256)
patchify_kernel(const float* __restrict__ img, T* __restrict__ out, int C, int H, int W, int P, long long total) {
  const int gw = W / P, gh = H / P;
  const int F = C * P * P;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(e % F);
    long long row = e / F;
    const int pw = f % P, ph = (f / P) % P, c = f / (P * P);
    const int px = (int)(row % gw); row /= gw;
    const int py = (int)(row % gh);
    const long long b = row / gh;
    stf<T>(out, e, img[((b * C + c) * H + py * P + ph) * (long long)W + px * P + pw]);
  }
}

// ------------------------------------------------------------------------------ batched transpose
template <typename T>
__global__ void __launch_bounds__(256)
transpose_kernel(const T* __restrict__ x, T* __restrict__ y, int rows, int cols) {
  __shared__ float tile[32][33];
  const long long base = (long long)blockIdx.z * rows * cols;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int k = ty; k < 32; k += 8)
    if (r0 + k < rows && c0 + tx < cols) tile[k][tx] = ldf<T>(x, base + (long long)(r0 + k) * cols + c0 + tx);
  __syncthreads();
  for (int k = ty; k < 32; k += 8)
    if (c0 + k < cols && r0 + tx < rows) stf<T>(y, base + (long long)(c0 + k) * rows + r0 + tx, tile[tx][k]);
}

// ------------------------------------------------------------------------------ LayerNorm
// one warp per row; each lane owns columns lane, lane+32, ...  (dim <= 32 * LN_MAX_PER_LANE)
constexpr int LN_MAX_PER_LANE = 40;   // dim <= 1280: the cross-attention blocks of VIT4CamerasBaseLine (VITs.py:262)

template <typename T>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     const float* __restrict__ add, T* __restrict__ y, float* __restrict__ mean_out,
                     float* __restrict__ rstd_out, int rows, int dim, int add_rows, float eps) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int per = (dim + 31) / 32;
  float v[LN_MAX_PER_LANE];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < LN_MAX_PER_LANE; ++k) {
    if (k < per) {
      const int j = lane + 32 * k;
      v[k] = j < dim ? ldf<T>(x, (long long)row * dim + j) : 0.f;
      s += v[k];
    }
  }
  const float mean = warp_sum(s) / (float)dim;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < LN_MAX_PER_LANE; ++k) {
    if (k < per) {
      const int j = lane + 32 * k;
      const float d = j < dim ? v[k] - mean : 0.f;
      q += d * d;
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)dim + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
#pragma unroll
  for (int k = 0; k < LN_MAX_PER_LANE; ++k) {
    if (k < per) {
      const int j = lane + 32 * k;
      if (j < dim) {
        float o = (v[k] - mean) * rstd * gamma[j] + beta[j];
        if (add) o += add[(long long)(row % add_rows) * dim + j];
        stf<T>(y, (long long)row * dim + j, o);
      }
    }
  }
}

// dim == 256, bf16 (the ViT's token width): a lane owns 8 contiguous channels = one 16-byte load / store per tensor
// and row, 8-float accumulators instead of the generic kernel's 32-entry register arrays.
__device__ __forceinline__ void unpack8_bf16(const uint4& q, float* v) {
  v[0] = bf16lo(q.x); v[1] = bf16hi(q.x); v[2] = bf16lo(q.y); v[3] = bf16hi(q.y);
  v[4] = bf16lo(q.z); v[5] = bf16hi(q.z); v[6] = bf16lo(q.w); v[7] = bf16hi(q.w);
}
__device__ __forceinline__ uint4 pack8_bf16(const float* v) {
  return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

__global__ void __launch_bounds__(256)
layernorm_fwd256_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                        const float* __restrict__ beta, const float* __restrict__ add, __nv_bfloat16* __restrict__ y,
                        float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows, int add_rows, float eps) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float v[8];
  unpack8_bf16(*reinterpret_cast<const uint4*>(x + (long long)row * 256 + lane * 8), v);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += v[k];
  const float mean = warp_sum(s) * (1.f / 256.f);
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) { const float d = v[k] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / 256.f) + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  float o[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int j = lane * 8 + k;
    o[k] = (v[k] - mean) * rstd * __ldg(gamma + j) + __ldg(beta + j);
    if (add) o[k] += __ldg(add + (long long)(row % add_rows) * 256 + j);
  }
  *reinterpret_cast<uint4*>(y + (long long)row * 256 + lane * 8) = pack8_bf16(o);
}

// kColsum: also emit the column sums of the OUTPUT gx (one partial row per block).  In the Transformer's backward gx is
// the residual-stream gradient, i.e. the output gradient of the nn.Linear that closes the previous block (to_out / the
// MLP's second Linear, pytorch_vit_encoder.py:20-23,55): its column sums are that layer's bias gradient, which then
// needs no pass of its own over the tensor.
template <bool kColsum>
__global__ void __launch_bounds__(256)
layernorm_bwd256_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ gy,
                        const float* __restrict__ gamma, const float* __restrict__ mean,
                        const float* __restrict__ rstd, const __nv_bfloat16* __restrict__ gx_add,
                        __nv_bfloat16* __restrict__ gx, float* __restrict__ dgamma_partial,
                        float* __restrict__ dbeta_partial, float* __restrict__ gx_colsum_partial, int rows) {
  __shared__ float sg[8][256], sb[8][256];
  __shared__ float sx[kColsum ? 8 : 1][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float gam[8], dg[8], db[8], dx[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { gam[k] = __ldg(gamma + lane * 8 + k); dg[k] = db[k] = dx[k] = 0.f; }
  for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
    const long long off = (long long)row * 256 + lane * 8;
    const uint4 qx = *reinterpret_cast<const uint4*>(x + off);
    const uint4 qg = *reinterpret_cast<const uint4*>(gy + off);
    uint4 qa = make_uint4(0u, 0u, 0u, 0u);
    if (gx_add) qa = *reinterpret_cast<const uint4*>(gx_add + off);
    const float mu = __ldg(mean + row), rs = __ldg(rstd + row);
    float xv[8], dy[8], av[8], g[8];
    unpack8_bf16(qx, xv);
    unpack8_bf16(qg, dy);
    unpack8_bf16(qa, av);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      xv[k] = (xv[k] - mu) * rs;
      g[k] = dy[k] * gam[k];
      dg[k] += dy[k] * xv[k];
      db[k] += dy[k];
      s1 += g[k];
      s2 += g[k] * xv[k];
    }
    s1 = warp_sum(s1) * (1.f / 256.f);
    s2 = warp_sum(s2) * (1.f / 256.f);
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = rs * (g[k] - s1 - xv[k] * s2) + av[k];
    const uint4 packed = pack8_bf16(o);
    *reinterpret_cast<uint4*>(gx + off) = packed;
    if (kColsum) {   // sums of the STORED (bf16-rounded) values: what a separate pass over gx would add up
      float ov[8];
      unpack8_bf16(packed, ov);
#pragma unroll
      for (int k = 0; k < 8; ++k) dx[k] += ov[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sg[warp][lane * 8 + k] = dg[k];
    sb[warp][lane * 8 + k] = db[k];
    if (kColsum) sx[warp][lane * 8 + k] = dx[k];
  }
  __syncthreads();
  {
    const int j = threadIdx.x;
    float a = 0.f, b = 0.f, c = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      a += sg[w][j];
      b += sb[w][j];
      if (kColsum) c += sx[w][j];
    }
    dgamma_partial[(long long)blockIdx.x * 256 + j] = a;
    dbeta_partial[(long long)blockIdx.x * 256 + j] = b;
    if (kColsum) gx_colsum_partial[(long long)blockIdx.x * 256 + j] = c;
  }
}

// each block (8 warps) walks rows blockIdx.x, +gridDim.x, ... and emits one dgamma/dbeta partial row
template <typename T>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const T* __restrict__ x, const T* __restrict__ gy, const float* __restrict__ gamma,
                     const float* __restrict__ mean, const float* __restrict__ rstd, const T* __restrict__ gx_add,
                     T* __restrict__ gx, float* __restrict__ dgamma_partial, float* __restrict__ dbeta_partial,
                     int rows, int dim) {
  extern __shared__ float sm[];  // [8][dim] x 2
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per = (dim + 31) / 32;
  float dg[LN_MAX_PER_LANE], db[LN_MAX_PER_LANE];
#pragma unroll
  for (int k = 0; k < LN_MAX_PER_LANE; ++k) dg[k] = db[k] = 0.f;
  for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
    const float mu = mean[row], rs = rstd[row];
    float xh[LN_MAX_PER_LANE], g[LN_MAX_PER_LANE];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAX_PER_LANE; ++k) {
      if (k < per) {
        const int j = lane + 32 * k;
        if (j < dim) {
          const float dy = ldf<T>(gy, (long long)row * dim + j);
          xh[k] = (ldf<T>(x, (long long)row * dim + j) - mu) * rs;
          g[k] = dy * gamma[j];
          dg[k] += dy * xh[k];
          db[k] += dy;
          s1 += g[k];
          s2 += g[k] * xh[k];
        } else {
          xh[k] = g[k] = 0.f;
        }
      }
    }
    s1 = warp_sum(s1) / (float)dim;
    s2 = warp_sum(s2) / (float)dim;
#pragma unroll
    for (int k = 0; k < LN_MAX_PER_LANE; ++k) {
      if (k < per) {
        const int j = lane + 32 * k;
        if (j < dim) {
          float o = rs * (g[k] - s1 - xh[k] * s2);
          if (gx_add) o += ldf<T>(gx_add, (long long)row * dim + j);
          stf<T>(gx, (long long)row * dim + j, o);
        }
      }
    }
  }
  float* sg = sm;
  float* sb = sm + 8 * dim;
#pragma unroll
  for (int k = 0; k < LN_MAX_PER_LANE; ++k) {
    if (k < per) {
      const int j = lane + 32 * k;
      if (j < dim) { sg[warp * dim + j] = dg[k]; sb[warp * dim + j] = db[k]; }
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < dim; j += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < 8; ++w) { a += sg[w * dim + j]; b += sb[w * dim + j]; }
    dgamma_partial[(long long)blockIdx.x * dim + j] = a;
    dbeta_partial[(long long)blockIdx.x * dim + j] = b;
  }
}

// ------------------------------------------------------------------------------ strided batched GEMM
// C[z](m,n) = alpha * sum_k A[z](m,k) * B[z](k,n);  z = (zb, zh);  element strides in elements.
struct BgemmP {
  const void* A; const void* B; void* C;
  int M, N, K, ZH;  // batch count = ZB*ZH, z -> (zb = z / ZH, zh = z % ZH)
  long long a_zb, a_zh, a_m, a_k;
  long long b_zb, b_zh, b_k, b_n;
  long long c_zb, c_zh, c_m, c_n;
  float alpha;
};

template <typename TA, typename TB, typename TC>
__global__ void __launch_bounds__(256)
bgemm_kernel(const BgemmP p) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int z = blockIdx.z, zb = z / p.ZH, zh = z % p.ZH;
  const TA* A = reinterpret_cast<const TA*>(p.A) + zb * p.a_zb + zh * p.a_zh;
  const TB* B = reinterpret_cast<const TB*>(p.B) + zb * p.b_zb + zh * p.b_zh;
  TC* C = reinterpret_cast<TC*>(p.C) + zb * p.c_zb + zh * p.c_zh;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // loader roles: choose the thread->element map so the unit-stride axis is walked by consecutive threads
  const bool a_k_fast = p.a_k == 1;
  const bool b_n_fast = p.b_n == 1;
  for (int k0 = 0; k0 < p.K; k0 += 16) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;  // 0..1023
      int m, k;
      if (a_k_fast) { k = idx & 15; m = idx >> 4; } else { m = idx & 63; k = idx >> 6; }
      const int gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < p.M && gk < p.K) ? ldf<TA>(A, gm * p.a_m + gk * p.a_k) : 0.f;
      int n, kb;
      if (b_n_fast) { n = idx & 63; kb = idx >> 6; } else { kb = idx & 15; n = idx >> 4; }
      const int gn = n0 + n, gkb = k0 + kb;
      Bs[kb][n] = (gn < p.N && gkb < p.K) ? ldf<TB>(B, gkb * p.b_k + gn * p.b_n) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
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
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty + 16 * i;
    if (gm >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx + 16 * j;
      if (gn < p.N) stf<TC>(C, gm * p.c_m + gn * p.c_n, p.alpha * acc[i][j]);
    }
  }
}

template <typename TA, typename TB, typename TC>
static int launch_bgemm(const BgemmP& p, int batches, cudaStream_t st, const char* what) {
  dim3 grid(cdiv(p.M, 64), cdiv(p.N, 64), batches);
  bgemm_kernel<TA, TB, TC><<<grid, 256, 0, st>>>(p);
  PB_LAUNCH_CHECK(what);
  return PB_OK;
}

// row softmax in place (fp32), one warp per row
__global__ void __launch_bounds__(256)
softmax_rows_kernel(float* __restrict__ s, long long rows, int n) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float* r = s + row * n;
  float mx = -INFINITY;
  for (int j = lane; j < n; j += 32) mx = fmaxf(mx, r[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int j = lane; j < n; j += 32) {
    const float e = expf(r[j] - mx);
    r[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  for (int j = lane; j < n; j += 32) r[j] *= inv;
}

// dS = P * (dP - rowsum(dP * P)), in place on dP
__global__ void __launch_bounds__(256)
softmax_bwd_rows_kernel(const float* __restrict__ probs, float* __restrict__ dp, long long rows, int n) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* pr = probs + row * n;
  float* d = dp + row * n;
  float dot = 0.f;
  for (int j = lane; j < n; j += 32) dot += d[j] * pr[j];
  dot = warp_sum(dot);
  for (int j = lane; j < n; j += 32) d[j] = pr[j] * (d[j] - dot);
}

template <typename T>
__global__ void __launch_bounds__(256)
gelu_bwd_kernel(const T* __restrict__ pre, const T* __restrict__ gy, T* __restrict__ gx, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = ldf<T>(pre, i);
    const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
    stf<T>(gx, i, ldf<T>(gy, i) * (cdf + x * pdf));
  }
}

// bf16 rows, 8 values (16 bytes) per thread; each block owns a row range and also emits the column sums of the gx it
// wrote (of the ROUNDED bf16 values: exactly what a separate pass over gx would add up)
__device__ __forceinline__ float gelu_grad(float x) {
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}

__global__ void __launch_bounds__(256)
gelu_bwd_rows_kernel(const __nv_bfloat16* __restrict__ pre, const __nv_bfloat16* __restrict__ gy,
                     __nv_bfloat16* __restrict__ gx, long long rows, int dim, float* __restrict__ partial) {
  __shared__ float red[256 * 8];
  const int tpr = dim >> 3, rpp = 256 / tpr;
  const int r = threadIdx.x / tpr, v = threadIdx.x - r * tpr;
  const long long per = (rows + gridDim.x - 1) / gridDim.x;
  const long long r0 = (long long)blockIdx.x * per, r1 = min(rows, r0 + per);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (r < rpp) {
    for (long long row = r0 + r; row < r1; row += rpp) {
      const long long off = row * dim + v * 8;
      const uint4 a = ld_stream16(pre + off), g = ld_stream16(gy + off);
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        o[e] = pack_bf16x2(bf16lo(gw[e]) * gelu_grad(bf16lo(aw[e])), bf16hi(gw[e]) * gelu_grad(bf16hi(aw[e])));
        acc[2 * e] += bf16lo(o[e]);
        acc[2 * e + 1] += bf16hi(o[e]);
      }
      *reinterpret_cast<uint4*>(gx + off) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.x * 8 + j] = acc[j];
  __syncthreads();
  for (int c = threadIdx.x; c < dim; c += 256) {
    float s = 0.f;
    for (int rr = 0; rr < rpp; ++rr) s += red[(rr * tpr + (c >> 3)) * 8 + (c & 7)];
    partial[(long long)blockIdx.x * dim + c] = s;
  }
}

// ------------------------------------------------------------------------------ min/max normalise
__device__ __forceinline__ uint32_t ord_key(float v) {
  if (v != v) return 0xFFFFFFFFu;
  v = v + 0.0f;
  const uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord_val(uint32_t k) {
  if (k == 0xFFFFFFFFu) return __uint_as_float(0x7FC00000u);
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

struct MinMaxScratch {  // 16 bytes
  uint32_t min_key, max_key;
  float min_v, max_v;
};

__global__ void minmax_init_kernel(MinMaxScratch* s) {
  s->min_key = 0xFFFFFFFFu;
  s->max_key = 0u;
}

__global__ void __launch_bounds__(256)
minmax_reduce_kernel(const float* __restrict__ x, MinMaxScratch* s, long long n) {
  __shared__ uint32_t smin[8], smax[8];
  uint32_t lo = 0xFFFFFFFFu, hi = 0u;
  bool nan = false;
  const long long n4 = n / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const uint4 raw = ld_stream16(x + i * 4);
    const uint32_t r[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float v = __uint_as_float(r[e]);
      nan |= (v != v);
      const uint32_t k = ord_key(v);
      lo = min(lo, k);
      hi = max(hi, k);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n4 * 4; i < n; ++i) {
      const uint32_t k = ord_key(x[i]);
      nan |= (x[i] != x[i]);
      lo = min(lo, k);
      hi = max(hi, k);
    }
  if (nan) lo = 0xFFFFFFFFu;  // torch: min() of a tensor holding NaN is NaN
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    nan |= __shfl_xor_sync(0xffffffffu, (int)nan, o) != 0;
  }
  if (nan) lo = 0xFFFFFFFFu;
  if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lo; smax[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    bool any_nan = false;
    for (int w = 0; w < 8; ++w) {
      any_nan |= (smin[w] == 0xFFFFFFFFu && smax[w] == 0xFFFFFFFFu);
      lo = min(lo, smin[w]);
      hi = max(hi, smax[w]);
    }
    if (hi == 0xFFFFFFFFu) {  // a NaN somewhere: force both ends to NaN
      atomicMax(&s->max_key, 0xFFFFFFFFu);
      atomicMax(&s->min_key, 0xFFFFFFFFu);
    } else {
      atomicMin(&s->min_key, lo);
      atomicMax(&s->max_key, hi);
    }
    (void)any_nan;
  }
}

__global__ void minmax_finish_kernel(MinMaxScratch* s) {
  // NaN anywhere -> max_key is all ones; make the minimum NaN too (torch semantics)
  if (s->max_key == 0xFFFFFFFFu) s->min_key = 0xFFFFFFFFu;
  s->min_v = ord_val(s->min_key);
  s->max_v = ord_val(s->max_key);
}

__global__ void __launch_bounds__(256)
minmax_apply_kernel(const float* __restrict__ x, float* __restrict__ y, const MinMaxScratch* s, long long n) {
  const float lo = s->min_v, range = s->max_v - s->min_v;
  const long long n4 = n / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    reinterpret_cast<float4*>(y)[i] =
        make_float4((v.x - lo) / range, (v.y - lo) / range, (v.z - lo) / range, (v.w - lo) / range);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n4 * 4; i < n; ++i) y[i] = (x[i] - lo) / range;
}

struct MinMaxBwdScratch {  // 32 bytes
  double s1, s2;                       // sum gy, sum gy*x
  unsigned long long argmin, argmax;   // first flat index holding the min / max
};

__global__ void minmax_bwd_init_kernel(MinMaxBwdScratch* b) {
  b->s1 = 0.0; b->s2 = 0.0;
  b->argmin = ~0ull; b->argmax = ~0ull;
}

__global__ void __launch_bounds__(256)
minmax_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ gy, const MinMaxScratch* s,
                         MinMaxBwdScratch* b, long long n) {
  __shared__ float red[32];
  const float lo = s->min_v, hi = s->max_v;
  float s1 = 0.f, s2 = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float xv = x[i], g = gy[i];
    s1 += g;
    s2 += g * xv;
    if (xv == lo) atomicMin(&b->argmin, (unsigned long long)i);
    if (xv == hi) atomicMin(&b->argmax, (unsigned long long)i);
  }
  s1 = block_sum(s1, red);
  s2 = block_sum(s2, red);
  if (threadIdx.x == 0) {
    atomicAdd(&b->s1, (double)s1);
    atomicAdd(&b->s2, (double)s2);
  }
}

__global__ void __launch_bounds__(256)
minmax_bwd_apply_kernel(const float* __restrict__ gy, const MinMaxScratch* s, const MinMaxBwdScratch* b,
                        float* __restrict__ gx, long long n) {
  const double lo = s->min_v, hi = s->max_v;
  const double range = hi - lo;
  const float inv = (float)(1.0 / range);
  // d y_j / d min = (x_j - max)/range^2 ; d y_j / d max = -(x_j - min)/range^2
  const float gmin = (float)((b->s2 - hi * b->s1) / (range * range));
  const float gmax = (float)(-(b->s2 - lo * b->s1) / (range * range));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float g = gy[i] * inv;
    if ((unsigned long long)i == b->argmin) g += gmin;
    if ((unsigned long long)i == b->argmax) g += gmax;
    gx[i] = g;
  }
}

static inline int grid_cap(long long items, int threads, int per_sm) {
  long long g = (items + threads - 1) / threads;
  const long long cap = (long long)sm_count() * per_sm;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

// min / max of x into `mm` (MinMaxScratch) and a cleared MinMaxBwdScratch at `bw`: the first pass of
// pb_minmax_mse_fwd_bwd (bandwidth.cu)
int minmax_reduce_launch(const float* x, void* mm, void* bw, long long n, cudaStream_t st) {
  MinMaxScratch* s = (MinMaxScratch*)mm;
  minmax_init_kernel<<<1, 1, 0, st>>>(s);
  minmax_bwd_init_kernel<<<1, 1, 0, st>>>((MinMaxBwdScratch*)bw);
  minmax_reduce_kernel<<<grid_cap(n / 4 + 1, 256, 8), 256, 0, st>>>(x, s, n);
  minmax_finish_kernel<<<1, 1, 0, st>>>(s);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "minmax_reduce_launch");
  note_launches(3);
  return PB_OK;
}

// ---- tensor-core attention (bf16): csrc/tc_bgemm.cu
struct BgOperand {
  const void* ptr;
  long long s_zb, s_zh, s_i, s_k;
  int extent_i;
};
int tc_bgemm(const BgOperand& A, const BgOperand& B, int M, int N, int K, int ZH, int ZB, void* C, long long c_zb,
             long long c_zh, long long c_m, bool c_bf16, float alpha, int mode, const void* P, cudaStream_t st);

// fused two-product kernels (csrc/tc_attn.cu); PB_ERR_UNSUPPORTED = shape outside them, use the single products
int attn_fwd_fused(const void* qkv, void* probs, void* out, int B, int S, int H, int D, float scale, cudaStream_t st);
int attn_bwd_fused(const void* qkv, const void* probs, const void* gout, void* gqkv, void* ds, int B, int S, int H,
                   int D, float scale, cudaStream_t st);

static bool attention_tc_ok(int S, int D) {
  const char* off = getenv("POSEB200_ATTN_SIMT");
  if (off != nullptr && off[0] == '1') return false;
  return S >= 16 && S <= 256 && (S & 7) == 0 && D >= 16 && D <= 256 && (D & 7) == 0;
}

// probabilities are kept in bf16 inside the caller's `probs` buffer ([B][H][S][S], first half of the fp32-sized
// allocation); pb_attention_bwd reads them back the same way
static int attention_fwd_tc(const pb_attention_fwd_args* a, cudaStream_t st) {
  const int S = a->S, D = a->D, H = a->H, HD = H * D;
  const __nv_bfloat16* qkv = (const __nv_bfloat16*)a->qkv;
  const long long qkv_b = (long long)S * 3 * HD, pr_b = (long long)H * S * S, pr_h = (long long)S * S;
  __nv_bfloat16* probs = (__nv_bfloat16*)a->probs;
  int rc = attn_fwd_fused(qkv, probs, a->out, a->B, S, H, D, a->scale, st);
  if (rc != PB_ERR_UNSUPPORTED) return rc;
  BgOperand q{qkv, qkv_b, D, 3 * HD, 1, S}, k{qkv + HD, qkv_b, D, 3 * HD, 1, S};
  rc = tc_bgemm(q, k, S, S, D, H, a->B, probs, pr_b, pr_h, S, true, a->scale, 1, nullptr, st);   // softmax(QK^T)
  if (rc != PB_OK) return rc;
  BgOperand pm{probs, pr_b, pr_h, S, 1, S};
  BgOperand v{qkv + 2 * HD, qkv_b, D, 1, 3 * HD, D};                                                // B[n=d][k=s'] = V[s'][d]
  return tc_bgemm(pm, v, S, D, S, H, a->B, a->out, (long long)S * HD, D, HD, true, 1.f, 0, nullptr, st);
}

static int attention_bwd_tc(const pb_attention_bwd_args* a, cudaStream_t st) {
  const int S = a->S, D = a->D, H = a->H, HD = H * D;
  const __nv_bfloat16* qkv = (const __nv_bfloat16*)a->qkv;
  const __nv_bfloat16* go = (const __nv_bfloat16*)a->gout;
  __nv_bfloat16* gqkv = (__nv_bfloat16*)a->gqkv;
  const __nv_bfloat16* probs = (const __nv_bfloat16*)a->probs;
  __nv_bfloat16* ds = (__nv_bfloat16*)a->dprobs_ws;
  const long long qkv_b = (long long)S * 3 * HD, pr_b = (long long)H * S * S, pr_h = (long long)S * S;
  const long long go_b = (long long)S * HD;
  int rc = attn_bwd_fused(qkv, probs, go, gqkv, ds, a->B, S, H, D, a->scale, st);
  if (rc != PB_ERR_UNSUPPORTED) return rc;
  // dV = P^T dO:  A[m=k][kk=q] = P[q][k] (m contiguous), B[n=d][kk=q] = dO[q][d] (n contiguous)
  BgOperand pt{probs, pr_b, pr_h, 1, S, S}, dot{go, go_b, D, 1, HD, D};
  rc = tc_bgemm(pt, dot, S, D, S, H, a->B, gqkv + 2 * HD, qkv_b, D, 3 * HD, true, 1.f, 0, nullptr, st);
  if (rc != PB_OK) return rc;
  // dS = softmax'(dO V^T) * scale
  BgOperand dO{go, go_b, D, HD, 1, S}, v{qkv + 2 * HD, qkv_b, D, 3 * HD, 1, S};
  rc = tc_bgemm(dO, v, S, S, D, H, a->B, ds, pr_b, pr_h, S, true, a->scale, 2, probs, st);
  if (rc != PB_OK) return rc;
  // dQ = dS K:  B[n=d][kk=k] = K[k][d] (n contiguous)
  BgOperand dsm{ds, pr_b, pr_h, S, 1, S}, kt{qkv + HD, qkv_b, D, 1, 3 * HD, D};
  rc = tc_bgemm(dsm, kt, S, D, S, H, a->B, gqkv, qkv_b, D, 3 * HD, true, 1.f, 0, nullptr, st);
  if (rc != PB_OK) return rc;
  // dK = dS^T Q:  A[m=k][kk=q] = dS[q][k], B[n=d][kk=q] = Q[q][d]
  BgOperand dst{ds, pr_b, pr_h, 1, S, S}, qt{qkv, qkv_b, D, 1, 3 * HD, D};
  return tc_bgemm(dst, qt, S, D, S, H, a->B, gqkv + HD, qkv_b, D, 3 * HD, true, 1.f, 0, nullptr, st);
}

template <typename T>
static int attention_fwd_t(const pb_attention_fwd_args* a, cudaStream_t st) {
  const int HD = a->H * a->D;
  BgemmP p;
  // scores = scale * Q K^T   -> probs buffer [B][H][S][S]
  p.A = a->qkv; p.B = (const T*)a->qkv + HD; p.C = a->probs;
  p.M = a->S; p.N = a->S; p.K = a->D; p.ZH = a->H;
  p.a_zb = (long long)a->S * 3 * HD; p.a_zh = a->D; p.a_m = 3 * HD; p.a_k = 1;
  p.b_zb = p.a_zb; p.b_zh = a->D; p.b_k = 1; p.b_n = 3 * HD;
  p.c_zb = (long long)a->H * a->S * a->S; p.c_zh = (long long)a->S * a->S; p.c_m = a->S; p.c_n = 1;
  p.alpha = a->scale;
  int rc = launch_bgemm<T, T, float>(p, a->B * a->H, st, "attention scores");
  if (rc != PB_OK) return rc;
  const long long rows = (long long)a->B * a->H * a->S;
  softmax_rows_kernel<<<cdiv(rows, 8), 256, 0, st>>>(a->probs, rows, a->S);
  PB_LAUNCH_CHECK("softmax_rows_kernel");
  // out = P V
  p.A = a->probs; p.B = (const T*)a->qkv + 2 * HD; p.C = a->out;
  p.M = a->S; p.N = a->D; p.K = a->S;
  p.a_zb = (long long)a->H * a->S * a->S; p.a_zh = (long long)a->S * a->S; p.a_m = a->S; p.a_k = 1;
  p.b_zb = (long long)a->S * 3 * HD; p.b_zh = a->D; p.b_k = 3 * HD; p.b_n = 1;
  p.c_zb = (long long)a->S * HD; p.c_zh = a->D; p.c_m = HD; p.c_n = 1;
  p.alpha = 1.f;
  return launch_bgemm<float, T, T>(p, a->B * a->H, st, "attention PV");
}

template <typename T>
static int attention_bwd_t(const pb_attention_bwd_args* a, cudaStream_t st) {
  const int HD = a->H * a->D;
  const long long qkv_b = (long long)a->S * 3 * HD;
  const long long pr_b = (long long)a->H * a->S * a->S, pr_h = (long long)a->S * a->S;
  BgemmP p;
  p.ZH = a->H;
  // dV = P^T dO            [S x D] = [S x S]^T [S x D]
  p.A = a->probs; p.B = a->gout; p.C = (T*)a->gqkv + 2 * HD;
  p.M = a->S; p.N = a->D; p.K = a->S;
  p.a_zb = pr_b; p.a_zh = pr_h; p.a_m = 1; p.a_k = a->S;
  p.b_zb = (long long)a->S * HD; p.b_zh = a->D; p.b_k = HD; p.b_n = 1;
  p.c_zb = qkv_b; p.c_zh = a->D; p.c_m = 3 * HD; p.c_n = 1;
  p.alpha = 1.f;
  int rc = launch_bgemm<float, T, T>(p, a->B * a->H, st, "attention dV");
  if (rc != PB_OK) return rc;
  // dP = dO V^T            [S x S] = [S x D] [S x D]^T
  p.A = a->gout; p.B = (const T*)a->qkv + 2 * HD; p.C = a->dprobs_ws;
  p.M = a->S; p.N = a->S; p.K = a->D;
  p.a_zb = (long long)a->S * HD; p.a_zh = a->D; p.a_m = HD; p.a_k = 1;
  p.b_zb = qkv_b; p.b_zh = a->D; p.b_k = 1; p.b_n = 3 * HD;
  p.c_zb = pr_b; p.c_zh = pr_h; p.c_m = a->S; p.c_n = 1;
  rc = launch_bgemm<T, T, float>(p, a->B * a->H, st, "attention dP");
  if (rc != PB_OK) return rc;
  const long long rows = (long long)a->B * a->H * a->S;
  softmax_bwd_rows_kernel<<<cdiv(rows, 8), 256, 0, st>>>(a->probs, a->dprobs_ws, rows, a->S);
  PB_LAUNCH_CHECK("softmax_bwd_rows_kernel");
  // dQ = scale * dS K      [S x D] = [S x S] [S x D]
  p.A = a->dprobs_ws; p.B = (const T*)a->qkv + HD; p.C = a->gqkv;
  p.M = a->S; p.N = a->D; p.K = a->S;
  p.a_zb = pr_b; p.a_zh = pr_h; p.a_m = a->S; p.a_k = 1;
  p.b_zb = qkv_b; p.b_zh = a->D; p.b_k = 3 * HD; p.b_n = 1;
  p.c_zb = qkv_b; p.c_zh = a->D; p.c_m = 3 * HD; p.c_n = 1;
  p.alpha = a->scale;
  rc = launch_bgemm<float, T, T>(p, a->B * a->H, st, "attention dQ");
  if (rc != PB_OK) return rc;
  // dK = scale * dS^T Q    [S x D] = [S x S]^T [S x D]
  p.A = a->dprobs_ws; p.B = a->qkv; p.C = (T*)a->gqkv + HD;
  p.a_m = 1; p.a_k = a->S;
  return launch_bgemm<float, T, T>(p, a->B * a->H, st, "attention dK");
}

}  // namespace pb

using namespace pb;

// ------------------------------------------------------------------------------ column-block move (pb_colblock)
template <typename T, int V>
__global__ void __launch_bounds__(256)
colblock_kernel(const T* __restrict__ src, T* __restrict__ dst, long long rows, long long ncols, long long srs,
                long long drs, long long sc0, long long dc0, long long mod, long long fstride, int nfold, int acc) {
  const long long per_row = ncols / V;
  const long long total = rows * per_row;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / per_row, j = (e - r * per_row) * V;
    const long long sr = mod > 0 ? r % mod : r;
    float v[V];
#pragma unroll
    for (int k = 0; k < V; ++k) v[k] = acc ? ldf<T>(dst, r * drs + dc0 + j + k) : 0.f;
    for (int f = 0; f < nfold; ++f) {
      const T* sp = src + (long long)f * fstride + sr * srs + sc0 + j;
#pragma unroll
      for (int k = 0; k < V; ++k) v[k] += ldf<T>(sp, k);
    }
#pragma unroll
    for (int k = 0; k < V; ++k) stf<T>(dst, r * drs + dc0 + j + k, v[k]);
  }
}

template <typename T>
static int launch_colblock(const pb_colblock_args* a, cudaStream_t st) {
  // 8 consecutive elements per thread when every row segment starts on a multiple of 8 elements (the compiler turns
  // the unrolled loads / stores into 16-byte accesses for bf16); scalar otherwise
  const bool v8 = ((a->ncols | a->src_row_stride | a->dst_row_stride | a->src_col0 | a->dst_col0 | a->fold_stride) & 7) == 0;
  const long long total = a->rows * (v8 ? a->ncols / 8 : a->ncols);
  const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  if (v8)
    colblock_kernel<T, 8><<<grid, 256, 0, st>>>((const T*)a->src, (T*)a->dst, a->rows, a->ncols, a->src_row_stride,
                                                a->dst_row_stride, a->src_col0, a->dst_col0, a->src_rows_mod,
                                                a->fold_stride, a->nfold, a->accumulate);
  else
    colblock_kernel<T, 1><<<grid, 256, 0, st>>>((const T*)a->src, (T*)a->dst, a->rows, a->ncols, a->src_row_stride,
                                                a->dst_row_stride, a->src_col0, a->dst_col0, a->src_rows_mod,
                                                a->fold_stride, a->nfold, a->accumulate);
  PB_LAUNCH_CHECK("colblock_kernel");
  return PB_OK;
}

extern "C" {

int pb_colblock(const pb_colblock_args* a, void* stream) {
  PB_REQUIRE(a != nullptr && a->src && a->dst, "pb_colblock: null args");
  PB_REQUIRE(a->rows >= 0 && a->ncols >= 0 && a->nfold >= 1 && a->src_rows_mod >= 0, "pb_colblock: bad shape");
  PB_REQUIRE_DEV(a->src, "src");
  PB_REQUIRE_DEV(a->dst, "dst");
  if (a->rows == 0 || a->ncols == 0) return PB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (a->dtype == PB_BF16) return launch_colblock<__nv_bfloat16>(a, st);
  if (a->dtype == PB_F16) return launch_colblock<__half>(a, st);
  return launch_colblock<float>(a, st);
}

int pb_patchify(const pb_patchify_args* a, void* stream) {
  PB_REQUIRE(a != nullptr && a->img && a->patches, "pb_patchify: null args");
  PB_REQUIRE(a->B > 0 && a->C > 0 && a->P > 0 && a->H % a->P == 0 && a->W % a->P == 0,
             "pb_patchify: image dimensions must be divisible by the patch size");
  PB_REQUIRE_DEV(a->img, "img");
  PB_REQUIRE_DEV(a->patches, "patches");
  const long long total = (long long)a->B * a->C * a->H * a->W;
  cudaStream_t st = (cudaStream_t)stream;
  if (a->act_dtype == PB_BF16)
    patchify_kernel<__nv_bfloat16><<<grid_cap(total, 256, 16), 256, 0, st>>>(a->img, (__nv_bfloat16*)a->patches, a->C,
                                                                            a->H, a->W, a->P, total);
  else
    patchify_kernel<float><<<grid_cap(total, 256, 16), 256, 0, st>>>(a->img, (float*)a->patches, a->C, a->H, a->W,
                                                                    a->P, total);
  PB_LAUNCH_CHECK("patchify_kernel");
  return PB_OK;
}

int pb_batched_transpose(const pb_transpose_args* a, void* stream) {
  PB_REQUIRE(a != nullptr && a->x && a->y && a->batch > 0 && a->rows > 0 && a->cols > 0, "pb_batched_transpose: bad args");
  PB_REQUIRE(a->batch <= 65535, "pb_batched_transpose: batch must be <= 65535");
  PB_REQUIRE_DEV(a->x, "x");
  PB_REQUIRE_DEV(a->y, "y");
  dim3 grid(cdiv(a->cols, 32), cdiv(a->rows, 32), a->batch);
  cudaStream_t st = (cudaStream_t)stream;
  if (a->act_dtype == PB_BF16)
    transpose_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)a->x, (__nv_bfloat16*)a->y, a->rows,
                                                         a->cols);
  else
    transpose_kernel<float><<<grid, 256, 0, st>>>((const float*)a->x, (float*)a->y, a->rows, a->cols);
  PB_LAUNCH_CHECK("transpose_kernel");
  return PB_OK;
}

int pb_layernorm_fwd(const pb_layernorm_fwd_args* a, void* stream) {
  PB_REQUIRE(a != nullptr && a->x && a->y && a->gamma && a->beta, "pb_layernorm_fwd: null args");
  PB_REQUIRE(a->rows > 0 && a->dim > 0 && a->dim <= 32 * LN_MAX_PER_LANE, "pb_layernorm_fwd: dim must be <= %d",
             32 * LN_MAX_PER_LANE);
  PB_REQUIRE(a->add == nullptr || a->add_rows > 0, "pb_layernorm_fwd: add_rows");
  PB_REQUIRE_DEV(a->x, "x");
  PB_REQUIRE_DEV(a->y, "y");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = cdiv(a->rows, 8);
  if (a->act_dtype == PB_BF16 && a->dim == 256 && ((((uintptr_t)a->x) | ((uintptr_t)a->y)) & 15) == 0) {
    layernorm_fwd256_kernel<<<grid, 256, 0, st>>>((const __nv_bfloat16*)a->x, a->gamma, a->beta, a->add,
                                                  (__nv_bfloat16*)a->y, a->mean, a->rstd, a->rows,
                                                  a->add_rows > 0 ? a->add_rows : 1, a->eps);
    PB_LAUNCH_CHECK("layernorm_fwd256_kernel");
    return PB_OK;
  }
  if (a->act_dtype == PB_BF16)
    layernorm_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)a->x, a->gamma, a->beta, a->add,
                                                             (__nv_bfloat16*)a->y, a->mean, a->rstd, a->rows, a->dim,
                                                             a->add_rows, a->eps);
  else
    layernorm_fwd_kernel<float><<<grid, 256, 0, st>>>((const float*)a->x, a->gamma, a->beta, a->add, (float*)a->y,
                                                     a->mean, a->rstd, a->rows, a->dim, a->add_rows, a->eps);
  PB_LAUNCH_CHECK("layernorm_fwd_kernel");
  return PB_OK;
}

int pb_layernorm_bwd(const pb_layernorm_bwd_args* a, void* stream) {
  PB_REQUIRE(a != nullptr && a->x && a->gy && a->gx && a->gamma && a->mean && a->rstd && a->dgamma_partial &&
                 a->dbeta_partial,
             "pb_layernorm_bwd: null args");
  PB_REQUIRE(a->rows > 0 && a->dim > 0 && a->dim <= 32 * LN_MAX_PER_LANE && a->nblk >= 1, "pb_layernorm_bwd: shape");
  PB_REQUIRE_DEV(a->x, "x");
  PB_REQUIRE_DEV(a->gy, "gy");
  PB_REQUIRE_DEV(a->gx, "gx");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)16 * a->dim * sizeof(float);
  if (a->act_dtype == PB_BF16 && a->dim == 256 &&
      ((((uintptr_t)a->x) | ((uintptr_t)a->gy) | ((uintptr_t)a->gx) | ((uintptr_t)a->gx_add)) & 15) == 0) {
    if (a->gx_colsum_partial != nullptr)
      layernorm_bwd256_kernel<true><<<a->nblk, 256, 0, st>>>(
          (const __nv_bfloat16*)a->x, (const __nv_bfloat16*)a->gy, a->gamma, a->mean, a->rstd,
          (const __nv_bfloat16*)a->gx_add, (__nv_bfloat16*)a->gx, a->dgamma_partial, a->dbeta_partial,
          a->gx_colsum_partial, a->rows);
    else
      layernorm_bwd256_kernel<false><<<a->nblk, 256, 0, st>>>(
          (const __nv_bfloat16*)a->x, (const __nv_bfloat16*)a->gy, a->gamma, a->mean, a->rstd,
          (const __nv_bfloat16*)a->gx_add, (__nv_bfloat16*)a->gx, a->dgamma_partial, a->dbeta_partial, nullptr, a->rows);
    PB_LAUNCH_CHECK("layernorm_bwd256_kernel");
    return PB_OK;
  }
  if (a->gx_colsum_partial != nullptr) {
    set_error("pb_layernorm_bwd: gx_colsum_partial is served by the bf16 dim = 256 kernel only (16-byte aligned tensors)");
    return PB_ERR_UNSUPPORTED;
  }
  if (a->act_dtype == PB_BF16) {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(layernorm_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    layernorm_bwd_kernel<__nv_bfloat16><<<a->nblk, 256, smem, st>>>(
        (const __nv_bfloat16*)a->x, (const __nv_bfloat16*)a->gy, a->gamma, a->mean, a->rstd,
        (const __nv_bfloat16*)a->gx_add, (__nv_bfloat16*)a->gx, a->dgamma_partial, a->dbeta_partial, a->rows, a->dim);
  } else {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(layernorm_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    layernorm_bwd_kernel<float><<<a->nblk, 256, smem, st>>>((const float*)a->x, (const float*)a->gy, a->gamma, a->mean,
                                                           a->rstd, (const float*)a->gx_add, (float*)a->gx,
                                                           a->dgamma_partial, a->dbeta_partial, a->rows, a->dim);
  }
  PB_LAUNCH_CHECK("layernorm_bwd_kernel");
  return PB_OK;
}

int pb_attention_fwd(const pb_attention_fwd_args* a, void* stream) {
  PB_REQUIRE(a != nullptr && a->qkv && a->out && a->probs, "pb_attention_fwd: qkv, out and probs are required");
  PB_REQUIRE(a->B > 0 && a->S > 0 && a->H > 0 && a->D > 0, "pb_attention_fwd: shape");
  PB_REQUIRE_DEV(a->qkv, "qkv");
  PB_REQUIRE_DEV(a->out, "out");
  PB_REQUIRE_DEV(a->probs, "probs");
  if (a->act_dtype == PB_BF16 && attention_tc_ok(a->S, a->D)) return attention_fwd_tc(a, (cudaStream_t)stream);
  return a->act_dtype == PB_BF16 ? attention_fwd_t<__nv_bfloat16>(a, (cudaStream_t)stream)
                                 : attention_fwd_t<float>(a, (cudaStream_t)stream);
}

int pb_attention_bwd(const pb_attention_bwd_args* a, void* stream) {
  PB_REQUIRE(a != nullptr && a->qkv && a->probs && a->gout && a->gqkv && a->dprobs_ws, "pb_attention_bwd: null args");
  PB_REQUIRE(a->B > 0 && a->S > 0 && a->H > 0 && a->D > 0, "pb_attention_bwd: shape");
  PB_REQUIRE_DEV(a->qkv, "qkv");
  PB_REQUIRE_DEV(a->gqkv, "gqkv");
  if (a->act_dtype == PB_BF16 && attention_tc_ok(a->S, a->D)) return attention_bwd_tc(a, (cudaStream_t)stream);
  return a->act_dtype == PB_BF16 ? attention_bwd_t<__nv_bfloat16>(a, (cudaStream_t)stream)
                                 : attention_bwd_t<float>(a, (cudaStream_t)stream);
}

int pb_gelu_bwd(const pb_gelu_bwd_args* a, void* stream) {
  PB_REQUIRE(a != nullptr && a->pre && a->gy && a->gx && a->n >= 0, "pb_gelu_bwd: bad args");
  PB_REQUIRE_DEV(a->pre, "pre");
  if (a->n == 0) return PB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (a->colsum_partial != nullptr) {
    PB_REQUIRE(a->act_dtype == PB_BF16 && a->dim >= 8 && (a->dim & 7) == 0 && a->dim <= 2048 && a->n % a->dim == 0 &&
                   a->nblk >= 1, "pb_gelu_bwd: the column-sum variant takes bf16 [rows][dim], dim % 8 == 0, dim <= 2048");
    PB_REQUIRE_DEV(a->colsum_partial, "colsum_partial");
    gelu_bwd_rows_kernel<<<a->nblk, 256, 0, st>>>((const __nv_bfloat16*)a->pre, (const __nv_bfloat16*)a->gy,
                                                   (__nv_bfloat16*)a->gx, a->n / a->dim, a->dim, a->colsum_partial);
    PB_LAUNCH_CHECK("gelu_bwd_rows_kernel");
    return PB_OK;
  }
  if (a->act_dtype == PB_BF16)
    gelu_bwd_kernel<__nv_bfloat16><<<grid_cap(a->n, 256, 16), 256, 0, st>>>(
        (const __nv_bfloat16*)a->pre, (const __nv_bfloat16*)a->gy, (__nv_bfloat16*)a->gx, a->n);
  else
    gelu_bwd_kernel<float><<<grid_cap(a->n, 256, 16), 256, 0, st>>>((const float*)a->pre, (const float*)a->gy,
                                                                   (float*)a->gx, a->n);
  PB_LAUNCH_CHECK("gelu_bwd_kernel");
  return PB_OK;
}

int pb_minmax_normalize_fwd(const pb_minmax_norm_fwd_args* a, void* stream) {
  PB_REQUIRE(a != nullptr && a->x && a->y && a->minmax && a->n > 0, "pb_minmax_normalize_fwd: bad args");
  PB_REQUIRE((((uintptr_t)a->x | (uintptr_t)a->y) & 15) == 0, "pb_minmax_normalize_fwd: 16-byte alignment");
  PB_REQUIRE_DEV(a->x, "x");
  PB_REQUIRE_DEV(a->y, "y");
  PB_REQUIRE_DEV(a->minmax, "minmax");
  cudaStream_t st = (cudaStream_t)stream;
  MinMaxScratch* s = (MinMaxScratch*)a->minmax;
  minmax_init_kernel<<<1, 1, 0, st>>>(s);
  minmax_reduce_kernel<<<grid_cap(a->n / 4 + 1, 256, 8), 256, 0, st>>>(a->x, s, a->n);
  minmax_finish_kernel<<<1, 1, 0, st>>>(s);
  minmax_apply_kernel<<<grid_cap(a->n / 4 + 1, 256, 8), 256, 0, st>>>(a->x, a->y, s, a->n);
  PB_LAUNCH_CHECK("minmax_normalize_fwd");
  note_launches(3);
  return PB_OK;
}

int pb_minmax_normalize_bwd(const pb_minmax_norm_bwd_args* a, void* stream) {
  PB_REQUIRE(a != nullptr && a->x && a->gy && a->gx && a->minmax && a->scratch && a->n > 0,
             "pb_minmax_normalize_bwd: bad args");
  PB_REQUIRE_DEV(a->x, "x");
  PB_REQUIRE_DEV(a->gy, "gy");
  PB_REQUIRE_DEV(a->gx, "gx");
  cudaStream_t st = (cudaStream_t)stream;
  const MinMaxScratch* s = (const MinMaxScratch*)a->minmax;
  MinMaxBwdScratch* b = (MinMaxBwdScratch*)a->scratch;
  minmax_bwd_init_kernel<<<1, 1, 0, st>>>(b);
  minmax_bwd_reduce_kernel<<<grid_cap(a->n, 256, 8), 256, 0, st>>>(a->x, a->gy, s, b, a->n);
  minmax_bwd_apply_kernel<<<grid_cap(a->n, 256, 8), 256, 0, st>>>(a->gy, s, b, a->gx, a->n);
  PB_LAUNCH_CHECK("minmax_normalize_bwd");
  note_launches(2);
  return PB_OK;
}

}  // extern "C"
