// HBM-bound kernels of the pose hot path: MSE loss (+gradient, +fused Gaussian target),
// gradient ingest, Gaussian targets, arg-max / soft-arg-max peaks, fused Adam, max-pool
// (+LeakyReLU) forward/backward, weight packing and weight-gradient reduction.
//
// Design rules (DESIGN.md "bandwidth kernels"): 16-byte coalesced accesses, streaming
// (L1::no_allocate) loads for read-once data, warp-shuffle reductions, one atomic per CTA.
#include <math.h>

#include <type_traits>

#include "common.cuh"

namespace pb {

// =====================================================================================
// MSE loss + gradient (pytorch/train_pytorch.py:110,134-137) / gradient ingest.
// One CTA handles TILE_PX consecutive pixels of one sample for all C channels, so that the
// NCHW -> NHWC transposition of the gradient goes through shared memory and both sides stay
// fully coalesced.
// =====================================================================================
constexpr int MSE_TILE_PX = 128;
constexpr int MSE_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(MSE_THREADS)
mse_kernel(const float* __restrict__ out, const float* __restrict__ target, const float* __restrict__ points,
           const float* __restrict__ gin, float inv_two_sigma2, float* loss_sum, double* loss_sum64,
           float* __restrict__ grad_nchw, T* __restrict__ grad_nhwc, int C, int H, int W, int Cpad,
           float grad_scale, float slope) {
  extern __shared__ float tile[];  // [MSE_TILE_PX][Cpad + 1] when grad_nhwc
  __shared__ float red[32];
  const int HW = H * W;
  const int tiles_per_img = HW / MSE_TILE_PX;
  const int b = blockIdx.x / tiles_per_img;
  const int px0 = (blockIdx.x % tiles_per_img) * MSE_TILE_PX;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarp = MSE_THREADS / 32;
  const int ldt = Cpad + 1;
  float acc = 0.f;
  for (int c = warp; c < Cpad; c += nwarp) {
    float g4[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < C) {
      const long long base = ((long long)(b * C + c)) * HW + px0 + lane * 4;
      float o4[4] = {0.f, 0.f, 0.f, 0.f};
      float d4[4];
      if (out != nullptr) {
        const float4 o = *reinterpret_cast<const float4*>(out + base);
        o4[0] = o.x; o4[1] = o.y; o4[2] = o.z; o4[3] = o.w;
      }
      if (gin != nullptr) {  // ingest mode: upstream gradient given
        const float4 g = *reinterpret_cast<const float4*>(gin + base);
        d4[0] = g.x; d4[1] = g.y; d4[2] = g.z; d4[3] = g.w;
      } else {
        float t4[4];
        if (target != nullptr) {
          const float4 t = *reinterpret_cast<const float4*>(target + base);
          t4[0] = t.x; t4[1] = t.y; t4[2] = t.z; t4[3] = t.w;
        } else {  // fused Gaussian target, tensorflow/simple_data_generator.py:119-125
          const float mx = points[(b * C + c) * 2 + 0], my = points[(b * C + c) * 2 + 1];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int p = px0 + lane * 4 + e;
            const float dx = (float)(p % W) - mx, dy = (float)(p / W) - my;
            t4[e] = expf(-(dx * dx + dy * dy) * inv_two_sigma2);
          }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          d4[e] = o4[e] - t4[e];
          acc += d4[e] * d4[e];
          d4[e] *= grad_scale;
        }
        if (grad_nchw != nullptr)
          *reinterpret_cast<float4*>(grad_nchw + base) = make_float4(d4[0], d4[1], d4[2], d4[3]);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e)
        g4[e] = (out != nullptr) ? d4[e] * (o4[e] > 0.f ? 1.f : slope) : d4[e];
    }
    if (grad_nhwc != nullptr) {
#pragma unroll
      for (int e = 0; e < 4; ++e) tile[(lane * 4 + e) * ldt + c] = g4[e];
    }
  }
  if (grad_nhwc != nullptr) {
    __syncthreads();
    T* dst = grad_nhwc + ((long long)b * HW + px0) * Cpad;
    if (sizeof(T) == 2 && (Cpad & 7) == 0) {
      const int c8n = Cpad >> 3;
      for (int i = threadIdx.x; i < MSE_TILE_PX * c8n; i += MSE_THREADS) {
        const float* src = tile + (i / c8n) * ldt + (i % c8n) * 8;
        uint4 t;
        t.x = pack_bf16x2(src[0], src[1]); t.y = pack_bf16x2(src[2], src[3]);
        t.z = pack_bf16x2(src[4], src[5]); t.w = pack_bf16x2(src[6], src[7]);
        reinterpret_cast<uint4*>(dst)[i] = t;
      }
    } else {
      const int n = MSE_TILE_PX * Cpad;
      for (int i = threadIdx.x; i < n; i += MSE_THREADS) stf<T>(dst, i, tile[(i / Cpad) * ldt + (i % Cpad)]);
    }
  }
  if (loss_sum != nullptr || loss_sum64 != nullptr) {
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) {
      if (loss_sum64 != nullptr) atomicAdd(loss_sum64, (double)acc);
      if (loss_sum != nullptr) atomicAdd(loss_sum, acc);
    }
  }
}

// Training-path specialisation (bf16 NHWC gradient only, no upstream gradient, Cpad = 8 * C8N <= 64):
// a warp owns channel PAIRS, every lane issues all of its (coalesced, 128 B per warp) loads before the first
// use -- 8*NP fp32 values in flight per thread -- and the transpose tile holds packed bf16 pairs with an odd
// word stride, so both the tile writes and the 16-byte row gathers are bank-conflict free.  The kernel is
// instruction-issue bound once the loads overlap (ncu: not_selected + math = 51 % of the stall samples), so
// integer divisions are kept out of the per-element path: Cpad is a template constant and the pixel coordinates
// of a lane's four pixels come from one division per thread.
template <int C8N, bool HAS_TARGET>
__global__ void __launch_bounds__(MSE_THREADS)
mse_nhwc_bf16_kernel(const float* __restrict__ out, const float* __restrict__ target,
                     const float* __restrict__ points, float inv_two_sigma2, float* loss_sum, double* loss_sum64,
                     __nv_bfloat16* __restrict__ grad_nhwc, int C, int H, int W, int tiles_per_img, float grad_scale,
                     float slope) {
  constexpr int Cpad = 8 * C8N;
  constexpr int NP = (C8N + 1) / 2;           // channel pairs per warp (8 warps)
  constexpr int npairs = Cpad >> 1;
  constexpr int ldw = npairs | 1;
  __shared__ uint32_t tile[MSE_TILE_PX * ldw];
  __shared__ float red[32];
  const int HW = H * W;
  const int b = blockIdx.x / tiles_per_img;
  const int px0 = (blockIdx.x - b * tiles_per_img) * MSE_TILE_PX;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float o[NP][2][4], t[NP][2][4];
  const float* obase = out + (long long)b * C * HW + px0 + lane;
  const float* tbase = HAS_TARGET ? target + (long long)b * C * HW + px0 + lane : nullptr;
#pragma unroll
  for (int k = 0; k < NP; ++k)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = 2 * (warp + 8 * k) + h;
      const bool live = c < C;
      const float* po = obase + (long long)c * HW;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        o[k][h][e] = live ? __ldcs(po + 32 * e) : 0.f;
        if (HAS_TARGET) t[k][h][e] = live ? __ldcs(tbase + (long long)c * HW + 32 * e) : 0.f;
      }
    }
  float acc = 0.f;
  const float neg_k2 = -inv_two_sigma2 * 1.4426950408889634f;
  float fx[4], fy[4];   // pixel coordinates of this lane's four pixels (same for every channel)
  if (!HAS_TARGET) {
    const int p0 = px0 + lane;
    int y = p0 / W, x = p0 - y * W;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      fx[e] = (float)x;
      fy[e] = (float)y;
      x += 32;
      while (x >= W) { x -= W; ++y; }
    }
  }
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const int cp = warp + 8 * k;
    if (cp < npairs) {
      float g[2][4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = 2 * cp + h;
        const bool live = c < C;
        float mx = 0.f, my = 0.f;
        if (!HAS_TARGET && live) {
          const float2 m = __ldg(reinterpret_cast<const float2*>(points) + (b * C + c));
          mx = m.x; my = m.y;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float tv;
          if (HAS_TARGET) {
            tv = t[k][h][e];
          } else {  // fused Gaussian target, tensorflow/simple_data_generator.py:119-125
            // exp(-r^2 / 2 sigma^2) = 2^(-r^2 * log2(e) / 2 sigma^2): one MUFU.EX2 (rel. error 2^-22) on an argument
            // whose constant carries one more rounding -- the target is within 4e-6 (absolute) of the float64 reference
            // Gaussian; expf's range reduction would double this kernel's arithmetic, and it is issue-bound
            const float dx = fx[e] - mx, dy = fy[e] - my;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(tv) : "f"((dx * dx + dy * dy) * neg_k2));
          }
          const float d = live ? o[k][h][e] - tv : 0.f;
          acc += d * d;
          g[h][e] = d * (o[k][h][e] > 0.f ? grad_scale : grad_scale * slope);
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) tile[(32 * e + lane) * ldw + cp] = pack_bf16x2(g[0][e], g[1][e]);
    }
  }
  __syncthreads();
  {
    uint4* dst = reinterpret_cast<uint4*>(grad_nhwc + ((long long)b * HW + px0) * Cpad);
#pragma unroll
    for (int i = threadIdx.x; i < MSE_TILE_PX * C8N; i += MSE_THREADS) {
      const uint32_t* src = tile + (i / C8N) * ldw + (i % C8N) * 4;
      dst[i] = make_uint4(src[0], src[1], src[2], src[3]);
    }
  }
  if (loss_sum != nullptr || loss_sum64 != nullptr) {
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) {
      if (loss_sum64 != nullptr) atomicAdd(loss_sum64, (double)acc);
      if (loss_sum != nullptr) atomicAdd(loss_sum, acc);
    }
  }
}

template <int C8N>
static void launch_mse_nhwc_bf16(const float* out, const float* target, const float* points, float inv,
                                 float* loss_sum, double* loss64, void* grad_nhwc, int grid, int C, int H, int W,
                                 float grad_scale, float slope, cudaStream_t st) {
  const int tiles_per_img = H * W / MSE_TILE_PX;
  if (target != nullptr)
    mse_nhwc_bf16_kernel<C8N, true><<<grid, MSE_THREADS, 0, st>>>(out, target, points, inv, loss_sum, loss64,
                                                                  (__nv_bfloat16*)grad_nhwc, C, H, W, tiles_per_img,
                                                                  grad_scale, slope);
  else
    mse_nhwc_bf16_kernel<C8N, false><<<grid, MSE_THREADS, 0, st>>>(out, target, points, inv, loss_sum, loss64,
                                                                   (__nv_bfloat16*)grad_nhwc, C, H, W, tiles_per_img,
                                                                   grad_scale, slope);
}

// =====================================================================================
// Fused tail of the ViT model's training step (pb_minmax_mse_fwd_bwd): the same 128-pixel x all-channel tile walk
// as mse_nhwc_bf16_kernel over the PRE-normalisation heatmaps x.
//   PASS 1: y = (x - lo) / range, d = y - t;  loss += d^2;  g = d * grad_scale;  s1 += g;  s2 += g * x;
//           first flat index of x == lo / x == hi                                    (no gradient is written)
//   PASS 2: dx = g / range (+ the min / max terms at those two indices), times LeakyReLU'(x) -> bf16 NHWC
// =====================================================================================
struct TailScratch {   // = {MinMaxScratch, MinMaxBwdScratch} of vit.cu
  uint32_t min_key, max_key;
  float min_v, max_v;
  double s1, s2;
  unsigned long long argmin, argmax;
};

template <int C8N, bool HAS_TARGET, int PASS>
__global__ void __launch_bounds__(MSE_THREADS)
minmax_mse_kernel(const float* __restrict__ xin, const float* __restrict__ target, const float* __restrict__ points,
                  float inv_two_sigma2, TailScratch* sc, float* loss_sum, __nv_bfloat16* __restrict__ grad_nhwc, int C,
                  int H, int W, int tiles_per_img, float grad_scale, float slope) {
  constexpr int Cpad = 8 * C8N;
  constexpr int NP = (C8N + 1) / 2;
  constexpr int npairs = Cpad >> 1;
  constexpr int ldw = npairs | 1;
  __shared__ uint32_t tile[PASS == 2 ? MSE_TILE_PX * ldw : 1];
  __shared__ float red[32];
  const int HW = H * W;
  const int b = blockIdx.x / tiles_per_img;
  const int px0 = (blockIdx.x - b * tiles_per_img) * MSE_TILE_PX;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float o[NP][2][4], t[NP][2][4];
  const float* obase = xin + (long long)b * C * HW + px0 + lane;
  const float* tbase = HAS_TARGET ? target + (long long)b * C * HW + px0 + lane : nullptr;
#pragma unroll
  for (int k = 0; k < NP; ++k)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = 2 * (warp + 8 * k) + h;
      const bool live = c < C;
      const float* po = obase + (long long)c * HW;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        o[k][h][e] = live ? __ldcs(po + 32 * e) : 0.f;
        if (HAS_TARGET) t[k][h][e] = live ? __ldcs(tbase + (long long)c * HW + 32 * e) : 0.f;
      }
    }
  const float lo = sc->min_v, hi = sc->max_v, range = hi - lo;
  float inv = 0.f, gmin = 0.f, gmax = 0.f;
  unsigned long long amin = ~0ull, amax = ~0ull;
  if (PASS == 2) {   // as minmax_bwd_apply_kernel: d y_j / d min = (x_j - max) / range^2, d y_j / d max = -(x_j - min) / range^2
    const double dlo = lo, dhi = hi, r = dhi - dlo;
    inv = (float)(1.0 / r);
    gmin = (float)((sc->s2 - dhi * sc->s1) / (r * r));
    gmax = (float)(-(sc->s2 - dlo * sc->s1) / (r * r));
    amin = sc->argmin;
    amax = sc->argmax;
  }
  float acc = 0.f, s1 = 0.f, s2 = 0.f;
  const float neg_k2 = -inv_two_sigma2 * 1.4426950408889634f;
  float fx[4], fy[4];
  if (!HAS_TARGET) {
    const int p0 = px0 + lane;
    int y = p0 / W, x = p0 - y * W;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      fx[e] = (float)x;
      fy[e] = (float)y;
      x += 32;
      while (x >= W) { x -= W; ++y; }
    }
  }
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const int cp = warp + 8 * k;
    if (cp < npairs) {
      float g[2][4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = 2 * cp + h;
        const bool live = c < C;
        float mx = 0.f, my = 0.f;
        if (!HAS_TARGET && live) {
          const float2 m = __ldg(reinterpret_cast<const float2*>(points) + (b * C + c));
          mx = m.x; my = m.y;
        }
        const unsigned long long idx0 = ((unsigned long long)b * C + c) * HW + px0 + lane;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float tv;
          if (HAS_TARGET) {
            tv = t[k][h][e];
          } else {   // fused Gaussian target, as in mse_nhwc_bf16_kernel
            const float dx = fx[e] - mx, dy = fy[e] - my;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(tv) : "f"((dx * dx + dy * dy) * neg_k2));
          }
          const float xv = o[k][h][e];
          const float yv = (xv - lo) / range;                  // minmax_apply_kernel's arithmetic
          const float gy = live ? (yv - tv) * grad_scale : 0.f;
          if (PASS == 1) {
            const float d = live ? yv - tv : 0.f;
            acc += d * d;
            s1 += gy;
            s2 += gy * xv;
            if (live && xv == lo) atomicMin(&sc->argmin, idx0 + 32 * e);
            if (live && xv == hi) atomicMin(&sc->argmax, idx0 + 32 * e);
          } else {
            float gx = gy * inv;
            if (idx0 + 32 * e == amin) gx += gmin;
            if (idx0 + 32 * e == amax) gx += gmax;
            g[h][e] = live ? gx * (xv > 0.f ? 1.f : slope) : 0.f;
          }
        }
      }
      if (PASS == 2) {
#pragma unroll
        for (int e = 0; e < 4; ++e) tile[(32 * e + lane) * ldw + cp] = pack_bf16x2(g[0][e], g[1][e]);
      }
    }
  }
  if (PASS == 2) {
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(grad_nhwc + ((long long)b * HW + px0) * Cpad);
#pragma unroll
    for (int i = threadIdx.x; i < MSE_TILE_PX * C8N; i += MSE_THREADS) {
      const uint32_t* src = tile + (i / C8N) * ldw + (i % C8N) * 4;
      dst[i] = make_uint4(src[0], src[1], src[2], src[3]);
    }
  } else {
    acc = block_sum(acc, red);
    s1 = block_sum(s1, red);
    s2 = block_sum(s2, red);
    if (threadIdx.x == 0) {
      atomicAdd(loss_sum, acc);
      atomicAdd(&sc->s1, (double)s1);
      atomicAdd(&sc->s2, (double)s2);
    }
  }
}

template <int C8N>
static void launch_minmax_mse(const pb_minmax_mse_args* a, int grid, float inv, cudaStream_t st) {
  const int tiles_per_img = a->H * a->W / MSE_TILE_PX;
  TailScratch* sc = (TailScratch*)a->scratch;
  __nv_bfloat16* g = (__nv_bfloat16*)a->grad_nhwc;
#define PB_TAIL(HT, PASS) minmax_mse_kernel<C8N, HT, PASS><<<grid, MSE_THREADS, 0, st>>>( \
      a->x, a->target, a->points, inv, sc, a->loss_sum, g, a->C, a->H, a->W, tiles_per_img, a->grad_scale, a->slope)
  if (a->target != nullptr) { PB_TAIL(true, 1); PB_TAIL(true, 2); }
  else { PB_TAIL(false, 1); PB_TAIL(false, 2); }
#undef PB_TAIL
}

int minmax_reduce_launch(const float* x, void* mm, void* bw, long long n, cudaStream_t st);   // vit.cu

template <typename T>
static int launch_mse(const float* out, const float* target, const float* points, const float* gin, float sigma,
                      float* loss_sum, double* loss64, float* grad_nchw, void* grad_nhwc, int B, int C, int H,
                      int W, int Cpad, float grad_scale, float slope, cudaStream_t st) {
  const int HW = H * W;
  const int grid = B * (HW / MSE_TILE_PX);
  const size_t smem = grad_nhwc ? (size_t)MSE_TILE_PX * (Cpad + 1) * sizeof(float) : 0;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(mse_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "mse smem attr");
  }
  const float inv = sigma > 0.f ? 1.f / (2.f * sigma * sigma) : 0.f;
  if (sizeof(T) == 2 && out != nullptr && gin == nullptr && grad_nchw == nullptr && grad_nhwc != nullptr &&
      (Cpad & 7) == 0 && Cpad <= 64 && getenv("POSEB200_MSE_GENERIC") == nullptr) {
#define PB_MSE_CASE(N)                                                                                          \
  case N:                                                                                                       \
    launch_mse_nhwc_bf16<N>(out, target, points, inv, loss_sum, loss64, grad_nhwc, grid, C, H, W, grad_scale, \
                            slope, st);                                                                         \
    break;
    switch (Cpad >> 3) {
      PB_MSE_CASE(1) PB_MSE_CASE(2) PB_MSE_CASE(3) PB_MSE_CASE(4) PB_MSE_CASE(5) PB_MSE_CASE(6) PB_MSE_CASE(7)
      default:
        launch_mse_nhwc_bf16<8>(out, target, points, inv, loss_sum, loss64, grad_nhwc, grid, C, H, W, grad_scale,
                                slope, st);
        break;
    }
#undef PB_MSE_CASE
    PB_LAUNCH_CHECK("mse_nhwc_bf16_kernel");
    return PB_OK;
  }
  mse_kernel<T><<<grid, MSE_THREADS, smem, st>>>(out, target, points, gin, inv, loss_sum, loss64, grad_nchw,
                                                 (T*)grad_nhwc, C, H, W, grad_nhwc ? Cpad : C, grad_scale, slope);
  PB_LAUNCH_CHECK("mse_kernel");
  return PB_OK;
}

// =====================================================================================
// Gaussian target heatmaps (tensorflow/simple_data_generator.py:119-136): write-only.
// =====================================================================================
__global__ void __launch_bounds__(256)
gaussian_kernel(const float* __restrict__ points, float* __restrict__ out, int HW, int W, float inv_two_sigma2,
                long long total4) {
  // grid (vector chunks of one map, maps): no 64-bit division; one 32-bit division per 16-byte store
  const int map = blockIdx.y;
  const int nvec = HW >> 2;
  const float2 m = __ldg(reinterpret_cast<const float2*>(points) + map);
  const float mx = m.x, my = m.y;
  float* dst = out + (long long)map * HW;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += gridDim.x * blockDim.x) {
    const int p0 = i * 4;
    float v[4];
    if ((W & 3) == 0) {   // the four pixels share a row
      const int y = p0 / W, x0 = p0 - y * W;
      const float dy = (float)y - my;
      const float dy2 = dy * dy;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float dx = (float)(x0 + e) - mx;
        v[e] = expf(-(dx * dx + dy2) * inv_two_sigma2);
      }
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int p = p0 + e;
        const float dx = (float)(p % W) - mx, dy = (float)(p / W) - my;
        v[e] = expf(-(dx * dx + dy * dy) * inv_two_sigma2);
      }
    }
    *reinterpret_cast<float4*>(dst + p0) = make_float4(v[0], v[1], v[2], v[3]);
  }
  (void)total4;
}

// Round 2: the Gaussian is separable, exp(-(dx^2 + dy^2) k) = exp(-dx^2 k) * exp(-dy^2 k).  A CTA builds the W column
// factors and its rows' factors once (in DOUBLE, rounded to fp32: W + rows exponentials per CTA instead of one expf
// per pixel -- the per-pixel form was issue-bound at 77 % SM throughput / 34 % of DRAM,
// profiles/r1e_gauss_ncu_summary.txt) and every pixel is one LDS + one FMUL: a pure write stream.  Against the
// reference's float64 rendering rounded to fp32 the product of two correctly rounded factors is within 3 fp32
// roundings (1.8e-7 relative), tighter than the per-pixel fp32 expf whose ARGUMENT rounding alone costs up to
// 87 * 2^-23 ~ 1e-5 relative on the far tail.
constexpr int GAUSS_TAB = 2048;
__global__ void __launch_bounds__(256)
gaussian_sep_kernel(const float* __restrict__ points, float* __restrict__ out, int H, int W, double inv_two_sigma2,
                    int rows_per_cta) {
  __shared__ __align__(16) float ex[GAUSS_TAB];
  __shared__ float ey[GAUSS_TAB];
  const int map = blockIdx.y;
  const float2 m = __ldg(reinterpret_cast<const float2*>(points) + map);
  const int y0 = blockIdx.x * rows_per_cta;
  const int y1 = min(H, y0 + rows_per_cta);
  for (int x = threadIdx.x; x < W; x += 256) {
    const double d = (double)x - (double)m.x;
    ex[x] = (float)exp(-(d * d) * inv_two_sigma2);
  }
  for (int y = y0 + threadIdx.x; y < y1; y += 256) {
    const double d = (double)y - (double)m.y;
    ey[y - y0] = (float)exp(-(d * d) * inv_two_sigma2);
  }
  __syncthreads();
  const int wv = W >> 2;                       // the launcher checks W % 4 == 0
  const int nvec = (y1 - y0) * wv;
  float* dst = out + (long long)map * H * W + (long long)y0 * W;
  int r = (int)threadIdx.x / wv, xv = (int)threadIdx.x - r * wv;
  const int dr = 256 / wv, dx = 256 - dr * wv;
  for (int i = threadIdx.x; i < nvec; i += 256) {
    const float4 e = *reinterpret_cast<const float4*>(ex + 4 * xv);
    const float fy = ey[r];
    *reinterpret_cast<float4*>(dst + (long long)r * W + 4 * xv) = make_float4(e.x * fy, e.y * fy, e.z * fy, e.w * fy);
    xv += dx; r += dr;
    if (xv >= wv) { xv -= wv; ++r; }
  }
}

// =====================================================================================
// Affine / flip augmentation (Datagenerators.py:153-186): nearest-neighbour resampling with
// torch's own grid arithmetic.  affine_grid (align_corners=False) builds
//   base = (x + (0.5 - W/2), y + (0.5 - H/2), 1),  grid = base @ (theta^T / (W/2, H/2))
// with the 3-term dot product evaluated as fma(y, b, x*a) + c (verified element-exact against
// torchvision on 2.2 M pixels); grid_sample un-normalises with ((g + 1) * size - 1) / 2 and rounds
// half-to-even.  One thread = one output pixel, all channels (the source index is shared).  The batch gather
// (src_index) and ToTensor's uint8 -> /255 ride along.
// =====================================================================================
template <typename TIn, int PX>
__global__ void __launch_bounds__(256)
affine_nearest_kernel(const TIn* __restrict__ in, float* __restrict__ out, const float* __restrict__ theta,
                      const int* __restrict__ flips, const int* __restrict__ src_index, int C, int H, int W) {
  // ToTensor's x / 255 for the 256 possible bytes, correctly rounded once per CTA (a divide per value would make
  // the uint8 path issue-bound)
  __shared__ float u8_tab[sizeof(TIn) == 1 ? 256 : 1];
  __shared__ float rr[6];
  const int b = blockIdx.y;
  const int HW = H * W;
  const float hw = 0.5f * (float)W, hh = 0.5f * (float)H;
  // the sample's six normalised matrix entries: one correctly rounded division each, once per CTA (every thread
  // used to repeat the six dependent loads + divisions before its first gather)
  if (threadIdx.x < 6) rr[threadIdx.x] = __fdiv_rn(__ldg(theta + 6 * b + threadIdx.x), threadIdx.x < 3 ? hw : hh);
  if constexpr (sizeof(TIn) == 1) u8_tab[threadIdx.x] = __fdiv_rn((float)threadIdx.x, 255.f);
  __syncthreads();
  const float r00 = rr[0], r10 = rr[1], r20 = rr[2], r01 = rr[3], r11 = rr[4], r21 = rr[5];
  const int fl = flips != nullptr ? __ldg(flips + b) : 0;
  const long long sb = src_index != nullptr ? (long long)__ldg(src_index + b) : (long long)b;   // dataset row
  const float x_off = 0.5f - hw, y_off = 0.5f - hh;   // exact in fp32 for even and odd sizes alike
  // one CTA = a (TX*PX) x TY pixel tile of the output: its rotated source footprint stays compact (sector reuse in
  // L1); a thread owns PX consecutive pixels of a row (one 16-byte store per channel when PX = 4)
  constexpr int TX = PX == 4 ? 16 : 32, TY = 256 / TX;
  const int tiles_x = (W + TX * PX - 1) / (TX * PX);
  const int x0 = ((blockIdx.x % tiles_x) * TX + (threadIdx.x % TX)) * PX;
  const int y = (blockIdx.x / tiles_x) * TY + threadIdx.x / TX;
  if (x0 >= W || y >= H) return;     // PX = 4 is only launched for W % 4 == 0: a thread's pixels are all in or out
  const int ya = (fl & 2) ? H - 1 - y : y;
  const float by = __fadd_rn((float)ya, y_off);
  int src[PX];
  bool ok[PX];
#pragma unroll
  for (int j = 0; j < PX; ++j) {
    const int x = x0 + j;
    const int xa = (fl & 1) ? W - 1 - x : x;     // the flips act on the affine result
    const float bx = __fadd_rn((float)xa, x_off);
    const float gx = __fadd_rn(__fmaf_rn(by, r10, __fmul_rn(bx, r00)), r20);
    const float gy = __fadd_rn(__fmaf_rn(by, r11, __fmul_rn(bx, r01)), r21);
    const float ix = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)W), 1.f), 2.f);
    const float iy = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)H), 1.f), 2.f);
    const float fxr = rintf(ix), fyr = rintf(iy);
    ok[j] = fxr >= 0.f && fxr < (float)W && fyr >= 0.f && fyr < (float)H;   // false for NaN
    src[j] = ok[j] ? (int)fyr * W + (int)fxr : 0;
  }
  const TIn* ib = in + sb * C * HW;
  float* ob = out + (long long)b * C * HW + y * W + x0;
  auto fetch = [&](int c, int j) -> float {
    if (!ok[j]) return 0.f;
    if constexpr (sizeof(TIn) == 1) return u8_tab[__ldg(ib + (long long)c * HW + src[j])];
    else return __ldg(ib + (long long)c * HW + src[j]);
  };
  constexpr int CU = 4;   // channels per step: 16 / 4 independent loads in flight per thread
  int c = 0;
  for (; c + CU <= C; c += CU) {
    float v[CU][PX];
#pragma unroll
    for (int k = 0; k < CU; ++k)
#pragma unroll
      for (int j = 0; j < PX; ++j) v[k][j] = fetch(c + k, j);
#pragma unroll
    for (int k = 0; k < CU; ++k) {
      if constexpr (PX == 4) *reinterpret_cast<float4*>(ob + (long long)(c + k) * HW) = make_float4(v[k][0], v[k][1], v[k][2], v[k][3]);
      else ob[(long long)(c + k) * HW] = v[k][0];
    }
  }
  for (; c < C; ++c) {
    float v[PX];
#pragma unroll
    for (int j = 0; j < PX; ++j) v[j] = fetch(c, j);
    if constexpr (PX == 4) *reinterpret_cast<float4*>(ob + (long long)c * HW) = make_float4(v[0], v[1], v[2], v[3]);
    else ob[(long long)c * HW] = v[0];
  }
}

// =====================================================================================
// Peaks.  Arg-max uses a 64-bit key  (order-preserving float bits << 32) | (~flat_index)
// so "maximum value, lowest index on ties, NaN greatest" is a plain unsigned max and the
// cross-CTA combine is one atomicMax per CTA and map.  The peaks buffer itself ([N][C][2]
// floats = 8 bytes per map) holds the keys until the finalize kernel decodes them.
// =====================================================================================
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long k) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, k, o);
    k = other > k ? other : k;
  }
  return k;
}

// running best in increasing index order: strict '>' keeps the lowest index, a NaN is taken
// once and then never replaced.
__device__ __forceinline__ void upd(float v, uint32_t idx, float& bv, uint32_t& bi) {
  if (v > bv || (v != v && bv == bv)) { bv = v; bi = idx; }
}

__global__ void zero_u64_kernel(unsigned long long* p, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0ull;
}

// planar maps (stride_x == 1): grid (maps, splits).  Each CTA scans a contiguous range of rows.
template <typename T>
__global__ void __launch_bounds__(256)
argmax_planar_kernel(const T* __restrict__ hm, unsigned long long* keys, int C, int H, int W, long long stride_n,
                     long long stride_c, long long stride_y, int rows_per_split) {
  __shared__ unsigned long long red[8];
  const int map = blockIdx.x;
  const int n = map / C, c = map % C;
  const T* base = hm + n * stride_n + c * stride_c;
  const int y0 = blockIdx.y * rows_per_split;
  const int y1 = min(H, y0 + rows_per_split);
  float bv = -INFINITY;
  uint32_t bi = (uint32_t)(y0 * W);
  bool first = true;
  constexpr int V = 16 / sizeof(T);
  const bool vec = (W % V == 0) && (stride_y % V == 0) && ((((uintptr_t)base) & 15) == 0);
  if (vec) {
    const int wv = W / V;
    const int total = (y1 - y0) * wv;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const int y = y0 + i / wv, xv = i % wv;
      const uint4 raw = ld_stream16(base + (long long)y * stride_y + xv * V);
      const uint32_t idx0 = (uint32_t)(y * W + xv * V);
      if (first) { bi = idx0; first = false; }
      if constexpr (sizeof(T) == 4) {
        upd(__uint_as_float(raw.x), idx0 + 0, bv, bi);
        upd(__uint_as_float(raw.y), idx0 + 1, bv, bi);
        upd(__uint_as_float(raw.z), idx0 + 2, bv, bi);
        upd(__uint_as_float(raw.w), idx0 + 3, bv, bi);
      } else {
        const uint32_t r[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          upd(bf16lo(r[e]), idx0 + 2 * e, bv, bi);
          upd(bf16hi(r[e]), idx0 + 2 * e + 1, bv, bi);
        }
      }
    }
  } else {
    const int total = (y1 - y0) * W;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const int y = y0 + i / W, x = i % W;
      const uint32_t idx = (uint32_t)(y * W + x);
      if (first) { bi = idx; first = false; }
      upd(ldf<T>(base, (long long)y * stride_y + x), idx, bv, bi);
    }
  }
  unsigned long long k = first ? 0ull : make_key(bv, bi);
  k = warp_max_u64(k);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = k;
  __syncthreads();
  if (threadIdx.x < 32) {
    k = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0ull;
    k = warp_max_u64(k);
    if (threadIdx.x == 0) atomicMax(keys + map, k);
  }
}

// Round 2: the same scan in two phases.  The element-wise running best above costs ~12 instructions per element
// (96 per 16-byte load for bf16: the kernel was issue-bound at 80 % SM throughput and 56 % of DRAM,
// profiles/r1e_argmax_bf16_ncu_summary.txt).  Phase 1 keeps the running best per 16-BYTE VECTOR -- its NaN-propagating
// maximum from three packed HMNMX2.NAN (bf16) / FMNMX.NAN (fp32) -- under the same rule (strict '>' in increasing index
// order, a NaN taken once); phase 2 re-reads the one winning vector (an L2 hit) and runs the element-wise rule inside
// it.  The winner is the same element: the first NaN lies in the first vector holding one, else the first element
// equal to the maximum lies in the first vector whose maximum equals it.
__device__ __forceinline__ float max_nan(float a, float b) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
template <typename T>
__device__ __forceinline__ float vec16_max_nan(const uint4& raw) {
  if constexpr (sizeof(T) == 4) {
    return max_nan(max_nan(__uint_as_float(raw.x), __uint_as_float(raw.y)),
                   max_nan(__uint_as_float(raw.z), __uint_as_float(raw.w)));
  } else {
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
    const __nv_bfloat162 c = *reinterpret_cast<const __nv_bfloat162*>(&raw.z);
    const __nv_bfloat162 d = *reinterpret_cast<const __nv_bfloat162*>(&raw.w);
    const __nv_bfloat162 m = __hmax2_nan(__hmax2_nan(a, b), __hmax2_nan(c, d));
    const uint32_t mu = *reinterpret_cast<const uint32_t*>(&m);
    return max_nan(bf16lo(mu), bf16hi(mu));
  }
}

// planar maps whose rows are 16-byte aligned vectors (the launcher checks): grid (maps, splits) as above.
// kFlat: the map is contiguous (stride_y == W), so a vector's address is its flat index -- no division in the loop.
// The CTA's threads are folded at the VECTOR level (key = vector maximum, lowest vector index on ties, NaN greatest)
// and one thread runs phase 2 on the CTA's winning vector: a phase 2 per thread re-read 256 scattered sectors per
// map, 24 % extra DRAM traffic (profiles/r2z_argmax_bf16_ncu_summary.txt).  `direct` (one CTA per map, gridDim.y == 1):
// that thread writes the peak and the maximum itself -- no key buffer, no zero / finalize launches.
template <typename T, bool kFlat>
__global__ void __launch_bounds__(256)
argmax_planar_vec_kernel(const T* __restrict__ hm, unsigned long long* keys, float* __restrict__ values, int C, int H,
                         int W, long long stride_n, long long stride_c, long long stride_y, int rows_per_split,
                         int direct) {
  __shared__ unsigned long long red[8];
  const int map = blockIdx.x;
  const int n = map / C, c = map % C;
  const T* base = hm + n * stride_n + c * stride_c;
  const int y0 = blockIdx.y * rows_per_split;
  const int y1 = min(H, y0 + rows_per_split);
  constexpr int V = 16 / sizeof(T);
  const int wv = W / V;
  const int total = (y1 - y0) * wv;
  auto vec_ptr = [&](uint32_t idx0) -> const T* {       // idx0 = flat element index y * W + x of a vector's first element
    if constexpr (kFlat) return base + idx0;
    else return base + (long long)(idx0 / (uint32_t)W) * stride_y + idx0 % (uint32_t)W;
  };
  auto vec_index = [&](int i) -> uint32_t {
    if constexpr (kFlat) return (uint32_t)(y0 * W + i * V);
    else return (uint32_t)((y0 + i / wv) * W + (i % wv) * V);
  };
  float bv = -INFINITY;
  uint32_t bi = 0;
  bool first = true;
  int i = threadIdx.x;
  if (i < total) { bi = vec_index(i); first = false; }
  for (; i + 256 < total; i += 512) {                    // two loads in flight per thread
    const uint32_t ia = vec_index(i), ib = vec_index(i + 256);
    const uint4 ra = ld_stream16(vec_ptr(ia));
    const uint4 rb = ld_stream16(vec_ptr(ib));
    upd(vec16_max_nan<T>(ra), ia, bv, bi);
    upd(vec16_max_nan<T>(rb), ib, bv, bi);
  }
  if (i < total) {
    const uint32_t ia = vec_index(i);
    upd(vec16_max_nan<T>(ld_stream16(vec_ptr(ia))), ia, bv, bi);
  }
  unsigned long long k = first ? 0ull : make_key(bv, bi);
  k = warp_max_u64(k);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = k;
  __syncthreads();
  if (threadIdx.x >= 32) return;
  k = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0ull;
  k = warp_max_u64(k);
  if (threadIdx.x != 0 || k == 0ull) return;
  // phase 2: the element-wise rule inside the CTA's winning vector
  const uint32_t idx0 = 0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull);
  const uint4 raw = *reinterpret_cast<const uint4*>(vec_ptr(idx0));
  float ev = -INFINITY;
  uint32_t ei = idx0;
  if constexpr (sizeof(T) == 4) {
    upd(__uint_as_float(raw.x), idx0 + 0, ev, ei);
    upd(__uint_as_float(raw.y), idx0 + 1, ev, ei);
    upd(__uint_as_float(raw.z), idx0 + 2, ev, ei);
    upd(__uint_as_float(raw.w), idx0 + 3, ev, ei);
  } else {
    const uint32_t r[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      upd(bf16lo(r[e]), idx0 + 2 * e, ev, ei);
      upd(bf16hi(r[e]), idx0 + 2 * e + 1, ev, ei);
    }
  }
  if (direct) {   // what zero_u64_kernel + atomicMax + argmax_finalize_kernel produce for this map
    reinterpret_cast<float2*>(keys)[map] = make_float2((float)(ei % (uint32_t)W), (float)(ei / (uint32_t)W));
    if (values != nullptr) values[map] = key_to_float(order_key(ev));
  } else {
    atomicMax(keys + map, make_key(ev, ei));
  }
}

// channel-interleaved maps (stride_c == 1, the reference's (N,H,W,C) argument layout):
// thread (g, c) walks pixels g, g+G, ... so a warp reads consecutive addresses.
template <typename T>
__global__ void __launch_bounds__(1024)
argmax_interleaved_kernel(const T* __restrict__ hm, unsigned long long* keys, int C, int HW, int W,
                          long long stride_n, long long stride_y, long long stride_x, int px_per_split, int G) {
  const int n = blockIdx.x;
  const int c = threadIdx.x % C, g = threadIdx.x / C;
  if (g >= G) return;
  const int p0 = blockIdx.y * px_per_split;
  const int p1 = min(HW, p0 + px_per_split);
  const T* base = hm + n * stride_n + c;
  float bv = -INFINITY;
  uint32_t bi = 0;
  bool first = true;
  for (int p = p0 + g; p < p1; p += G) {
    const int y = p / W, x = p % W;
    if (first) { bi = (uint32_t)p; first = false; }
    upd(ldf<T>(base, y * stride_y + x * stride_x), (uint32_t)p, bv, bi);
  }
  if (!first) atomicMax(keys + n * C + c, make_key(bv, bi));
}

__global__ void argmax_finalize_kernel(float* peaks, float* values, int maps, int W) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= maps) return;
  const unsigned long long k = reinterpret_cast<const unsigned long long*>(peaks)[i];
  const uint32_t idx = 0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull);
  const float v = key_to_float((uint32_t)(k >> 32));
  reinterpret_cast<float2*>(peaks)[i] = make_float2((float)(idx % (uint32_t)W), (float)(idx / (uint32_t)W));
  if (values != nullptr) values[i] = v;
}

// torch.linspace(0, 1, steps) element i in fp32 (symmetric evaluation used by ATen)
__device__ __forceinline__ float linspace01(int i, int steps) {
  const float step = 1.0f / (float)(steps - 1);
  return (i < steps / 2) ? step * (float)i : 1.0f - step * (float)(steps - 1 - i);
}

// soft arg-max (pytorch/utils.py:47-83): one CTA per map (planar) -- three fp32 sums.
// kTable (round 2, planar maps of H, W <= SOFT_TABLE_W): the column / row weights linspace01(x, W), linspace01(y, H)
// come from shared-memory tables (one 16-byte LDS per four elements instead of ~8 instructions per element) and (row, vector) advance
// incrementally instead of by a division per load: 88 -> ~30 instructions per 16-byte load
// (profiles/r1e_softargmax_ncu_summary.txt: the kernel sat at 69 % SM throughput / 55 % of DRAM).  Same products, same
// order of additions as the direct form -> bit-identical sums.
constexpr int SOFT_TABLE_W = 2048;
template <typename T, bool kTable>
__global__ void __launch_bounds__(256)
softargmax_kernel(const T* __restrict__ hm, float* __restrict__ peaks, int C, int H, int W, long long stride_n,
                  long long stride_c, long long stride_y, long long stride_x) {
  __shared__ float red[32];
  const int map = blockIdx.x;
  const int n = map / C, c = map % C;
  const T* base = hm + n * stride_n + c * stride_c;
  float s = 0.f, sx = 0.f, sy = 0.f;
  const int total = H * W;
  constexpr int V = 16 / sizeof(T);
  if constexpr (kTable) {
    // the launcher has checked: stride_x == 1, W % V == 0, W <= SOFT_TABLE_W, rows / maps 16-byte aligned
    __shared__ __align__(16) float wx[SOFT_TABLE_W];
    __shared__ float wy[SOFT_TABLE_W];
    for (int x = threadIdx.x; x < W; x += blockDim.x) wx[x] = linspace01(x, W);
    for (int r = threadIdx.x; r < H; r += blockDim.x) wy[r] = linspace01(r, H);
    __syncthreads();
    const int wv = W / V;
    const int nvec = H * wv;
    int y = (int)threadIdx.x / wv, xv = (int)threadIdx.x - y * wv;
    const int dy = 256 / wv, dx = 256 - dy * wv;
#pragma unroll 2
    for (int i = threadIdx.x; i < nvec; i += 256) {
      const int x0 = xv * V;
      const uint4 raw = ld_stream16(base + (long long)y * stride_y + x0);
      float v[V], w[V];
      *reinterpret_cast<float4*>(w) = *reinterpret_cast<const float4*>(wx + x0);
      if constexpr (sizeof(T) == 4) {
        v[0] = __uint_as_float(raw.x); v[1] = __uint_as_float(raw.y);
        v[2] = __uint_as_float(raw.z); v[3] = __uint_as_float(raw.w);
      } else {
        *reinterpret_cast<float4*>(w + 4) = *reinterpret_cast<const float4*>(wx + x0 + 4);
        const uint32_t r[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) { v[2 * e] = bf16lo(r[e]); v[2 * e + 1] = bf16hi(r[e]); }
      }
      float rs = 0.f;
#pragma unroll
      for (int e = 0; e < V; ++e) {
        rs += v[e];
        sx += w[e] * v[e];
      }
      s += rs;
      sy += wy[y] * rs;
      xv += dx; y += dy;
      if (xv >= wv) { xv -= wv; ++y; }
    }
  } else if (stride_x == 1 && (W % V) == 0 && (stride_y % V) == 0 && ((((uintptr_t)base) & 15) == 0)) {
    // planar maps: 16-byte loads, one (row, column) decode per vector, the row weight applied to the vector sum
    const int wv = W / V;
    const int nvec = H * wv;
#pragma unroll 4
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
      const int y = i / wv, x0 = (i - y * wv) * V;
      const uint4 raw = ld_stream16(base + (long long)y * stride_y + x0);
      float v[V];
      if constexpr (sizeof(T) == 4) {
        v[0] = __uint_as_float(raw.x); v[1] = __uint_as_float(raw.y);
        v[2] = __uint_as_float(raw.z); v[3] = __uint_as_float(raw.w);
      } else {
        const uint32_t r[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) { v[2 * e] = bf16lo(r[e]); v[2 * e + 1] = bf16hi(r[e]); }
      }
      float rs = 0.f;
#pragma unroll
      for (int e = 0; e < V; ++e) {
        rs += v[e];
        sx += linspace01(x0 + e, W) * v[e];
      }
      s += rs;
      sy += linspace01(y, H) * rs;
    }
  } else {
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const int y = i / W, x = i % W;
      const float v = ldf<T>(base, y * stride_y + x * stride_x);
      s += v;
      sx += linspace01(x, W) * v;
      sy += linspace01(y, H) * v;
    }
  }
  // one exchange for the three sums (was three block reductions = six barriers per 147 KB map)
  __shared__ float red3[3][8];
  s = warp_sum(s); sx = warp_sum(sx); sy = warp_sum(sy);
  if ((threadIdx.x & 31) == 0) {
    red3[0][threadIdx.x >> 5] = s; red3[1][threadIdx.x >> 5] = sx; red3[2][threadIdx.x >> 5] = sy;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    s = sx = sy = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { s += red3[0][w]; sx += red3[1][w]; sy += red3[2][w]; }
    float cx = sx / s * (float)(W - 1), cy = sy / s * (float)(H - 1);
    // torch.clamp propagates NaN; fminf/fmaxf would not
    if (cx == cx) cx = fminf(fmaxf(cx, 0.f), (float)(W - 1));
    if (cy == cy) cy = fminf(fmaxf(cy, 0.f), (float)(H - 1));
    peaks[2 * map] = cx;
    peaks[2 * map + 1] = cy;
  }
}

// =====================================================================================
// Fused Adam (torch.optim.Adam single-tensor math, no amsgrad) over the flat buffer.
// =====================================================================================
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            long long n, float lr, float b1, float b2, float eps, float wd, float gscale, float bc1, float bc2_sqrt,
            const int* __restrict__ found_inf, const int* __restrict__ step_dev, const float* __restrict__ lr_dev) {
  if (found_inf != nullptr && *found_inf != 0) return;
  if (step_dev != nullptr) {
    // graph-safe form: the step number lives on the device; bias corrections in double like the host path
    __shared__ float s_bc[2];
    if (threadIdx.x == 0) {
      const double t = (double)(*step_dev + 1);
      s_bc[0] = (float)(1.0 - pow((double)b1, t));
      s_bc[1] = (float)sqrt(1.0 - pow((double)b2, t));
    }
    __syncthreads();
    bc1 = s_bc[0];
    bc2_sqrt = s_bc[1];
  }
  if (lr_dev != nullptr) lr = *lr_dev;
  const float step_size = lr / bc1;
  const long long n4 = n / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pa = &pp.x; const float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float gr = ga[e] * gscale;
      if (wd != 0.f) gr += wd * pa[e];
      ma[e] = b1 * ma[e] + (1.f - b1) * gr;
      va[e] = b2 * va[e] + (1.f - b2) * gr * gr;
      const float denom = sqrtf(va[e]) / bc2_sqrt + eps;
      pa[e] -= step_size * (ma[e] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  // tail
  const long long t0 = n4 * 4;
  const long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gi < n - t0) {
    const long long i = t0 + gi;
    float gr = g[i] * gscale;
    if (wd != 0.f) gr += wd * p[i];
    const float mi = b1 * m[i] + (1.f - b1) * gr;
    const float vi = b2 * v[i] + (1.f - b2) * gr * gr;
    m[i] = mi; v[i] = vi;
    p[i] -= step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
  }
}

// =====================================================================================
// 2x2/2 max-pool + LeakyReLU (pytorch/CNNs.py:77,82), NHWC.
// =====================================================================================
template <typename T>
__global__ void __launch_bounds__(256)
pool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int H, int W, int C, float slope, long long total) {
  const int OH = H / 2, OW = W / 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int ox = (int)(r % OW); r /= OW;
    const int oy = (int)(r % OH);
    const long long n = r / OH;
    const long long b = ((n * H + 2 * oy) * W + 2 * ox) * C + c;
    // scan order (0,0),(0,1),(1,0),(1,1); '>' or NaN takes, like ATen's max_pool2d
    float m = ldf<T>(x, b);
    float v = ldf<T>(x, b + C); if (v > m || v != v) m = v;
    v = ldf<T>(x, b + (long long)W * C); if (v > m || v != v) m = v;
    v = ldf<T>(x, b + (long long)W * C + C); if (v > m || v != v) m = v;
    stf<T>(y, i, lrelu(m, slope));
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
pool_bwd_kernel(const T* __restrict__ x, const T* __restrict__ gy, const uint32_t* __restrict__ mask,
                T* __restrict__ gx, T* __restrict__ gxm, int H, int W, int C, float slope, long long total) {
  const int OH = H / 2, OW = W / 2;
  const int words = (C + 31) / 32;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int ox = (int)(r % OW); r /= OW;
    const int oy = (int)(r % OH);
    const long long n = r / OH;
    const long long pix0 = (n * H + 2 * oy) * W + 2 * ox;
    const long long off[4] = {0, 1, (long long)W, (long long)W + 1};
    float m = ldf<T>(x, pix0 * C + c);
    int am = 0;
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      const float v = ldf<T>(x, (pix0 + off[k]) * C + c);
      if (v > m || v != v) { m = v; am = k; }
    }
    const float g = ldf<T>(gy, i) * (m > 0.f ? 1.f : slope);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long pix = pix0 + off[k];
      const float gk = (k == am) ? g : 0.f;
      stf<T>(gx, pix * C + c, gk);
      if (gxm != nullptr) {
        float s = 1.f;
        if (mask != nullptr) s = ((mask[pix * words + (c >> 5)] >> (c & 31)) & 1u) ? 1.f : slope;
        stf<T>(gxm, pix * C + c, gk * s);
      }
    }
  }
}

// ---- vectorised bf16 variants (C % 8 == 0): one thread = 8 channels (16 bytes) of one pooled pixel, every global
//      access a coalesced 16-byte load / store
template <bool F16 = false>
__device__ __forceinline__ void unpack8(const uint4& t, float* v) {
  const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    v[2 * k] = lo16<F16>(w[k]);
    v[2 * k + 1] = hi16<F16>(w[k]);
  }
}
template <bool F16 = false>
__device__ __forceinline__ uint4 pack8(const float* v) {
  uint4 t;
  t.x = pack16x2<F16>(v[0], v[1]); t.y = pack16x2<F16>(v[2], v[3]);
  t.z = pack16x2<F16>(v[4], v[5]); t.w = pack16x2<F16>(v[6], v[7]);
  return t;
}

template <bool F16>   // x and y in IEEE half ("fp16" precision) instead of bf16
__global__ void __launch_bounds__(256)
pool_fwd_vec_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ y2,
                    int H, int W, int C, float slope, long long total8) {
  const int OH = H / 2, OW = W / 2, C8 = C / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8;
       i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    long long r = i / C8;
    const int ox = (int)(r % OW); r /= OW;
    const int oy = (int)(r % OH);
    const long long n = r / OH;
    const __nv_bfloat16* b = x + (((n * H + 2 * oy) * W + 2 * ox) * C + c8 * 8);
    const uint4 q[4] = {ld_stream16(b), ld_stream16(b + C), ld_stream16(b + (long long)W * C),
                        ld_stream16(b + (long long)W * C + C)};
    float m[8], v[8];
    unpack8<F16>(q[0], m);
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      unpack8<F16>(q[k], v);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (v[j] > m[j] || v[j] != v[j]) m[j] = v[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = lrelu(m[j], slope);
    *reinterpret_cast<uint4*>(y + i * 8) = pack8<F16>(m);
    if (F16 && y2 != nullptr) *reinterpret_cast<uint4*>(y2 + i * 8) = pack8<false>(m);   // bf16 twin for the weight gradient
  }
}

template <bool XF16>   // forward activations x in IEEE half; gradients stay bf16
__global__ void __launch_bounds__(256)
pool_bwd_vec_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ gy,
                    const uint32_t* __restrict__ mask, __nv_bfloat16* __restrict__ gx, __nv_bfloat16* __restrict__ gxm,
                    int H, int W, int C, float slope, long long total8) {
  const int OH = H / 2, OW = W / 2, C8 = C / 8;
  const int words = (C + 31) / 32;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8;
       i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    long long r = i / C8;
    const int ox = (int)(r % OW); r /= OW;
    const int oy = (int)(r % OH);
    const long long n = r / OH;
    const long long pix0 = (n * H + 2 * oy) * W + 2 * ox;
    const long long off[4] = {0, 1, (long long)W, (long long)W + 1};
    uint4 q[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) q[k] = ld_stream16(x + (pix0 + off[k]) * C + c8 * 8);
    const uint4 gq = ld_stream16(gy + i * 8);
    uint32_t mw[4] = {0, 0, 0, 0};
    if (gxm != nullptr && mask != nullptr) {
#pragma unroll
      for (int k = 0; k < 4; ++k) mw[k] = __ldg(mask + (pix0 + off[k]) * words + (c8 >> 2)) >> ((c8 & 3) * 8);
    }
    float m[8], v[8], g[8];
    int am[8];
    unpack8<XF16>(q[0], m);
#pragma unroll
    for (int j = 0; j < 8; ++j) am[j] = 0;
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      unpack8<XF16>(q[k], v);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (v[j] > m[j] || v[j] != v[j]) { m[j] = v[j]; am[j] = k; }
    }
    unpack8(gq, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] *= (m[j] > 0.f ? 1.f : slope);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float o[8], om[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o[j] = (am[j] == k) ? g[j] : 0.f;
        const float sc = (mask == nullptr || ((mw[k] >> j) & 1u)) ? 1.f : slope;
        om[j] = o[j] * sc;
      }
      const long long e = (pix0 + off[k]) * C + c8 * 8;
      *reinterpret_cast<uint4*>(gx + e) = pack8(o);
      if (gxm != nullptr) *reinterpret_cast<uint4*>(gxm + e) = pack8(om);
    }
  }
}

// =====================================================================================
// weight packing / weight-gradient reduction / column sums / add
// =====================================================================================
struct KposArr { int v[PB_MAX_TAPS]; };

template <typename T>
__global__ void pack_weights_kernel(const float* __restrict__ src, T* __restrict__ dst, int ntaps, int I, int Ipad,
                                    int J, int Jpad, long long si, long long sj, KposArr kpos) {
  const long long total = (long long)ntaps * Ipad * Jpad;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(e % Jpad);
    const int i = (int)((e / Jpad) % Ipad);
    const int t = (int)(e / ((long long)Jpad * Ipad));
    const float v = (i < I && j < J) ? src[i * si + j * sj + kpos.v[t]] : 0.f;
    stf<T>(dst, e, v);
  }
}

// every layer's packed operand in one launch: grid (tile slots, items); descriptors live in device memory.
// dst[t][i][j] = src[i*si + j*sj + kpos[t]] in 32 x 32 (i, j) tiles: stores run along j (contiguous in dst); when the
// source is contiguous along i instead (si < sj: the transposing role of an nn.Linear / 1x1 weight) the tile is read
// along i and turned through shared memory, so neither side issues one 32-byte sector per element.
__global__ void __launch_bounds__(256)
pack_weights_multi_kernel(const pb_pack_weights_args* __restrict__ items) {
  __shared__ float tile[32][33];
  const pb_pack_weights_args& a = items[blockIdx.y];
  const int ntaps = a.ntaps, I = a.I, Ipad = a.Ipad, J = a.J, Jpad = a.Jpad;
  const long long si = a.stride_i, sj = a.stride_j;
  const float* __restrict__ src = a.src;
  const int dt = a.dst_dtype;
  const bool turn = si < sj;
  const int tiles_j = (Jpad + 31) >> 5, tiles_i = (Ipad + 31) >> 5;
  const int total_tiles = ntaps * tiles_i * tiles_j;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int tl = blockIdx.x; tl < total_tiles; tl += gridDim.x) {
    const int tj = tl % tiles_j, ti = (tl / tiles_j) % tiles_i, t = tl / (tiles_j * tiles_i);
    const int i0 = ti << 5, j0 = tj << 5, kp = a.kpos[t];
    float v[4];
    if (turn) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int j = j0 + ty + 8 * r, i = i0 + tx;
        tile[ty + 8 * r][tx] = (i < I && j < J) ? __ldg(src + i * si + j * sj + kp) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < 4; ++r) v[r] = tile[tx][ty + 8 * r];
      __syncthreads();
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int i = i0 + ty + 8 * r, j = j0 + tx;
        v[r] = (i < I && j < J) ? __ldg(src + i * si + j * sj + kp) : 0.f;
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = i0 + ty + 8 * r, j = j0 + tx;
      if (i < Ipad && j < Jpad) {
        const long long e = ((long long)t * Ipad + i) * Jpad + j;
        if (dt == PB_BF16) reinterpret_cast<__nv_bfloat16*>(a.dst)[e] = __float2bfloat16(v[r]);
        else if (dt == PB_F16) reinterpret_cast<__half*>(a.dst)[e] = __float2half_rn(v[r]);
        else reinterpret_cast<float*>(a.dst)[e] = v[r];
      }
    }
  }
}

// first-layer im2col: one thread per (pixel, 8 consecutive k) -> one 16-byte (bf16) store
template <typename T>
__global__ void __launch_bounds__(256)
im2col_first_kernel(const float* __restrict__ in, T* __restrict__ out, int C, int H, int W, int ks, int dil, int Kpad,
                    long long total) {
  const int groups = Kpad / 8;
  const int kk = ks * ks, ctr = (ks - 1) / 2;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int gidx = (int)(e % groups);
    long long pix = e / groups;
    const int x = (int)(pix % W);
    const int y = (int)((pix / W) % H);
    const long long n = pix / ((long long)W * H);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = gidx * 8 + j;
      float val = 0.f;
      if (k < C * kk) {
        const int ci = k / kk, r = (k % kk) / ks, s = k % ks;
        const int iy = y + dil * (r - ctr), ix = x + dil * (s - ctr);
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) val = __ldg(in + ((n * C + ci) * H + iy) * (long long)W + ix);
      }
      v[j] = val;
    }
    T* dst = out + pix * Kpad + gidx * 8;
    if constexpr (sizeof(T) == 2) {
      constexpr bool F16 = sizeof(T) == 2 && !std::is_same<T, __nv_bfloat16>::value;
      uint4 t;
      t.x = pack16x2<F16>(v[0], v[1]); t.y = pack16x2<F16>(v[2], v[3]);
      t.z = pack16x2<F16>(v[4], v[5]); t.w = pack16x2<F16>(v[6], v[7]);
      *reinterpret_cast<uint4*>(dst) = t;
    } else {
      reinterpret_cast<float4*>(dst)[0] = make_float4(v[0], v[1], v[2], v[3]);
      reinterpret_cast<float4*>(dst)[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
}

// first-layer im2col, 3x3 fast path (C*9 <= 64 = Kpad, bf16 out): one lane per pixel so every
// (ci, r, s) load is a coalesced 128-byte row segment (the 9-fold re-reads hit L1); the 128-byte
// output row of each pixel is staged in shared memory (16-byte chunks XOR-swizzled by pixel) and
// written back as fully coalesced 512-byte warp stores.
template <int C, bool F16>
__global__ void __launch_bounds__(256)
im2col3_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int H, int W, int dil,
                    long long npix) {
  __shared__ uint4 stage[8][32 * 8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long nwarps = (long long)gridDim.x * 8;
  for (long long base = ((long long)blockIdx.x * 8 + warp) * 32; base < npix; base += nwarps * 32) {
    const long long pix = base + lane;
    float v[64];
#pragma unroll
    for (int k = 0; k < 64; ++k) v[k] = 0.f;
    if (pix < npix) {
      const int x = (int)(pix % W);
      const int y = (int)((pix / W) % H);
      const long long n = pix / ((long long)W * H);
#pragma unroll
      for (int ci = 0; ci < C; ++ci) {
        const float* plane = in + (n * C + ci) * (long long)H * W;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int iy = y + dil * (r - 1);
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            const int ix = x + dil * (s - 1);
            if (iy >= 0 && iy < H && ix >= 0 && ix < W) v[ci * 9 + r * 3 + s] = __ldg(plane + (long long)iy * W + ix);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint4 t;
      t.x = pack16x2<F16>(v[8 * j + 0], v[8 * j + 1]); t.y = pack16x2<F16>(v[8 * j + 2], v[8 * j + 3]);
      t.z = pack16x2<F16>(v[8 * j + 4], v[8 * j + 5]); t.w = pack16x2<F16>(v[8 * j + 6], v[8 * j + 7]);
      stage[warp][lane * 8 + (j ^ (lane & 7))] = t;
    }
    __syncwarp();
    uint4* dst = reinterpret_cast<uint4*>(out + base * 64);
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int idx = it * 32 + lane;       // 16-byte chunk index inside the warp's 4 KB output span
      const int row = idx >> 3, j = idx & 7;
      if (base + row < npix) dst[idx] = stage[warp][row * 8 + (j ^ (row & 7))];
    }
    __syncwarp();
  }
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw,
                                    float* __restrict__ dbias, int ksplit, int ntaps, int Ca, int Cg, long long sa,
                                    long long sg, KposArr kpos, float beta, float alpha, int Ca_valid) {
  const long long nW = (long long)ntaps * Ca * Cg;
  const long long L = nW + Cg;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < L;
       e += (long long)gridDim.x * blockDim.x) {
    if (e >= nW && dbias == nullptr) continue;
    if (e < nW && (int)((e / Cg) % Ca) >= Ca_valid) continue;  // channel padding rows
    float s = 0.f;
    for (int k = 0; k < ksplit; ++k) s += partial[k * L + e];
    s *= alpha;
    if (e < nW) {
      const int co = (int)(e % Cg);
      const int ci = (int)((e / Cg) % Ca);
      const int t = (int)(e / ((long long)Cg * Ca));
      const long long d = ci * sa + co * sg + kpos.v[t];
      dw[d] = (beta != 0.f ? beta * dw[d] : 0.f) + s;
    } else {
      const long long d = e - nW;
      dbias[d] = (beta != 0.f ? beta * dbias[d] : 0.f) + s;
    }
  }
}

// nn.Linear / 1x1 weights (one tap, parameter contiguous along ci: dw[co*sg + ci]): the partial tiles are [ci][co]
// (co contiguous), so a thread-per-element fold writes one 4-byte element per 32-byte sector.  Fold 32 x 32 tiles
// instead: read along co, turn through shared memory, write along ci.
__global__ void __launch_bounds__(256)
wgrad_reduce_turn_kernel(const float* __restrict__ partial, float* __restrict__ dw, float* __restrict__ dbias, int ksplit,
                         int Ca, int Cg, long long sg, float beta, float alpha, int Ca_valid) {
  __shared__ float tile[32][33];
  const long long nW = (long long)Ca * Cg, L = nW + Cg;
  const int tiles_co = (Cg + 31) >> 5, tiles_ci = (Ca_valid + 31) >> 5;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int tl = blockIdx.x; tl < tiles_ci * tiles_co; tl += gridDim.x) {
    const int ci0 = (tl / tiles_co) << 5, co0 = (tl % tiles_co) << 5;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int ci = ci0 + ty + 8 * r, co = co0 + tx;
      float s = 0.f;
      if (ci < Ca_valid && co < Cg) {
        const float* src = partial + (long long)ci * Cg + co;
#pragma unroll 4
        for (int k = 0; k < ksplit; ++k) s += __ldg(src + k * L);
      }
      tile[ty + 8 * r][tx] = s * alpha;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int co = co0 + ty + 8 * r, ci = ci0 + tx;
      if (ci < Ca_valid && co < Cg) {
        const long long d = (long long)co * sg + ci;
        dw[d] = (beta != 0.f ? beta * dw[d] : 0.f) + tile[tx][ty + 8 * r];
      }
    }
    __syncthreads();
  }
  if (dbias != nullptr)
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < Cg; c += gridDim.x * blockDim.x) {
      float s = 0.f;
      for (int k = 0; k < ksplit; ++k) s += partial[k * L + nW + c];
      dbias[c] = (beta != 0.f ? beta * dbias[c] : 0.f) + s * alpha;
    }
}

// out[j] = beta*out[j] + alpha * sum_b partial[b][j]: a block owns 32 columns, its 8 warps take rows b = w, w+8, ...
// (four independent loads in flight each) and are folded in warp order through shared memory -- a fixed summation
// order, and no single thread walks all nblk rows as one latency chain.
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ partial, float* __restrict__ out, const T* __restrict__ partial2,
              float* __restrict__ out2, int nblk, int dim, float alpha, float beta) {
  __shared__ float red[8][32];
  if (blockIdx.y == 1) { partial = partial2; out = out2; }
  const int tx = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (j < dim) {
    int b = w;
    for (; b + 24 < nblk; b += 32) {
      s0 += ldf<T>(partial, (long long)b * dim + j);
      s1 += ldf<T>(partial, (long long)(b + 8) * dim + j);
      s2 += ldf<T>(partial, (long long)(b + 16) * dim + j);
      s3 += ldf<T>(partial, (long long)(b + 24) * dim + j);
    }
    for (; b < nblk; b += 8) s0 += ldf<T>(partial, (long long)b * dim + j);
  }
  red[w][tx] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (w == 0 && j < dim) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][tx];
    out[j] = (beta != 0.f ? beta * out[j] : 0.f) + alpha * s;
  }
}

template <typename T>
__global__ void add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, long long n,
                           const uint32_t* __restrict__ mask, int C, float slope) {
  const int words = (C + 31) / 32;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = ldf<T>(a, i) + (b ? ldf<T>(b, i) : 0.f);
    if (mask != nullptr) {
      const int c = (int)(i % C);
      const uint32_t bit = (mask[(i / C) * words + (c >> 5)] >> (c & 31)) & 1u;
      v *= bit ? 1.f : slope;
    }
    stf<T>(out, i, v);
  }
}

static inline int grid_for(long long work_items, int threads, int per_sm = 8) {
  long long g = (work_items + threads - 1) / threads;
  const long long cap = (long long)sm_count() * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace pb

using namespace pb;

extern "C" {

int pb_mse_loss_fwd_bwd(const pb_mse_args* a, void* stream) {
  PB_REQUIRE(a != nullptr, "pb_mse_loss_fwd_bwd: null args");
  PB_REQUIRE(a->out != nullptr && (a->target != nullptr || a->points != nullptr),
             "pb_mse_loss_fwd_bwd: out and (target or points) required");
  PB_REQUIRE(a->B > 0 && a->C > 0 && a->H > 0 && a->W > 0, "pb_mse_loss_fwd_bwd: empty shape");
  PB_REQUIRE((a->H * a->W) % MSE_TILE_PX == 0, "pb_mse_loss_fwd_bwd: H*W must be a multiple of %d", MSE_TILE_PX);
  PB_REQUIRE(a->grad_nhwc == nullptr || a->Cpad >= a->C, "pb_mse_loss_fwd_bwd: Cpad < C");
  PB_REQUIRE_DEV(a->out, "out");
  PB_REQUIRE_DEV(a->target, "target");
  PB_REQUIRE_DEV(a->points, "points");
  PB_REQUIRE_DEV(a->loss_sum, "loss_sum");
  PB_REQUIRE_DEV(a->grad_nchw, "grad_nchw");
  PB_REQUIRE_DEV(a->grad_nhwc, "grad_nhwc");
  cudaStream_t st = (cudaStream_t)stream;
  if (a->act_dtype == PB_BF16)
    return launch_mse<__nv_bfloat16>(a->out, a->target, a->points, nullptr, a->sigma, a->loss_sum, a->loss_sum_f64,
                                     a->grad_nchw, a->grad_nhwc, a->B, a->C, a->H, a->W, a->Cpad, a->grad_scale,
                                     a->slope, st);
  return launch_mse<float>(a->out, a->target, a->points, nullptr, a->sigma, a->loss_sum, a->loss_sum_f64,
                           a->grad_nchw, a->grad_nhwc, a->B, a->C, a->H, a->W, a->Cpad, a->grad_scale, a->slope, st);
}

int pb_minmax_mse_fwd_bwd(const pb_minmax_mse_args* a, void* stream) {
  PB_REQUIRE(a != nullptr && a->x && a->loss_sum && a->grad_nhwc && a->scratch, "pb_minmax_mse_fwd_bwd: null args");
  PB_REQUIRE(a->target != nullptr || a->points != nullptr, "pb_minmax_mse_fwd_bwd: target or points required");
  PB_REQUIRE(a->B > 0 && a->C > 0 && a->H > 0 && a->W > 0, "pb_minmax_mse_fwd_bwd: empty shape");
  PB_REQUIRE((a->H * a->W) % MSE_TILE_PX == 0, "pb_minmax_mse_fwd_bwd: H*W must be a multiple of %d", MSE_TILE_PX);
  PB_REQUIRE(a->Cpad >= a->C && (a->Cpad & 7) == 0 && a->Cpad <= 64,
             "pb_minmax_mse_fwd_bwd: Cpad must be a multiple of 8, >= C and <= 64");
  PB_REQUIRE((((uintptr_t)a->scratch) & 15) == 0 && (((uintptr_t)a->x) & 15) == 0, "pb_minmax_mse_fwd_bwd: alignment");
  PB_REQUIRE_DEV(a->x, "x");
  PB_REQUIRE_DEV(a->target, "target");
  PB_REQUIRE_DEV(a->points, "points");
  PB_REQUIRE_DEV(a->loss_sum, "loss_sum");
  PB_REQUIRE_DEV(a->grad_nhwc, "grad_nhwc");
  PB_REQUIRE_DEV(a->scratch, "scratch");
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)a->B * a->C * a->H * a->W;
  int rc = minmax_reduce_launch(a->x, a->scratch, (char*)a->scratch + 16, n, st);
  if (rc != PB_OK) return rc;
  const int grid = a->B * (a->H * a->W / MSE_TILE_PX);
  const float inv = a->sigma > 0.f ? 1.f / (2.f * a->sigma * a->sigma) : 0.f;
  switch (a->Cpad >> 3) {
    case 1: launch_minmax_mse<1>(a, grid, inv, st); break;
    case 2: launch_minmax_mse<2>(a, grid, inv, st); break;
    case 3: launch_minmax_mse<3>(a, grid, inv, st); break;
    case 4: launch_minmax_mse<4>(a, grid, inv, st); break;
    case 5: launch_minmax_mse<5>(a, grid, inv, st); break;
    case 6: launch_minmax_mse<6>(a, grid, inv, st); break;
    case 7: launch_minmax_mse<7>(a, grid, inv, st); break;
    default: launch_minmax_mse<8>(a, grid, inv, st); break;
  }
  PB_LAUNCH_CHECK("minmax_mse_kernel");
  note_launches(1);
  return PB_OK;
}

int pb_grad_ingest(const pb_grad_ingest_args* a, void* stream) {
  PB_REQUIRE(a != nullptr && a->grad_nchw != nullptr && a->grad_nhwc != nullptr, "pb_grad_ingest: null args");
  PB_REQUIRE((a->H * a->W) % MSE_TILE_PX == 0, "pb_grad_ingest: H*W must be a multiple of %d", MSE_TILE_PX);
  PB_REQUIRE(a->Cpad >= a->C, "pb_grad_ingest: Cpad < C");
  PB_REQUIRE_DEV(a->grad_nchw, "grad_nchw");
  PB_REQUIRE_DEV(a->out_nchw, "out_nchw");
  PB_REQUIRE_DEV(a->grad_nhwc, "grad_nhwc");
  cudaStream_t st = (cudaStream_t)stream;
  if (a->act_dtype == PB_BF16)
    return launch_mse<__nv_bfloat16>(a->out_nchw, nullptr, nullptr, a->grad_nchw, 0.f, nullptr, nullptr, nullptr,
                                     a->grad_nhwc, a->B, a->C, a->H, a->W, a->Cpad, 1.f, a->slope, st);
  return launch_mse<float>(a->out_nchw, nullptr, nullptr, a->grad_nchw, 0.f, nullptr, nullptr, nullptr, a->grad_nhwc,
                           a->B, a->C, a->H, a->W, a->Cpad, 1.f, a->slope, st);
}

int pb_gaussian_heatmaps(const pb_gaussian_args* a, void* stream) {
  PB_REQUIRE(a != nullptr && a->points != nullptr && a->out != nullptr, "pb_gaussian_heatmaps: null args");
  PB_REQUIRE(a->BC >= 0 && a->H > 0 && a->W > 0 && a->sigma > 0.f, "pb_gaussian_heatmaps: bad shape/sigma");
  PB_REQUIRE((a->H * a->W) % 4 == 0, "pb_gaussian_heatmaps: H*W must be a multiple of 4");
  PB_REQUIRE_DEV(a->points, "points");
  PB_REQUIRE_DEV(a->out, "out");
  if (a->BC == 0) return PB_OK;
  if ((a->W & 3) == 0 && a->W <= GAUSS_TAB && a->H <= GAUSS_TAB && getenv("POSEB200_GAUSS_V1") == nullptr &&
      (reinterpret_cast<uintptr_t>(a->out) & 15) == 0) {
    // separable form: a CTA renders `rows` rows of one map; enough CTAs for ~8 per SM when there are few maps
    int splits = 1;
    const int target_ctas = sm_count() * 8;
    if (a->BC < target_ctas) splits = min(a->H, cdiv(target_ctas, a->BC));
    const int rows = cdiv(a->H, splits);
    splits = cdiv(a->H, rows);
    const double inv = 1.0 / (2.0 * (double)a->sigma * (double)a->sigma);
    for (int m0 = 0; m0 < a->BC; m0 += 65535) {          // gridDim.y limit
      const int nm = a->BC - m0 < 65535 ? a->BC - m0 : 65535;
      gaussian_sep_kernel<<<dim3((unsigned)splits, (unsigned)nm), 256, 0, (cudaStream_t)stream>>>(
          a->points + 2 * (size_t)m0, a->out + (size_t)m0 * a->H * a->W, a->H, a->W, inv, rows);
    }
    PB_LAUNCH_CHECK("gaussian_sep_kernel");
    return PB_OK;
  }
  const long long total4 = (long long)a->BC * a->H * a->W / 4;
  const int nvec = a->H * a->W / 4;
  int chunks = (nvec + 256 * 4 - 1) / (256 * 4);      // four 16-byte stores per thread
  if (chunks < 1) chunks = 1;
  for (int m0 = 0; m0 < a->BC; m0 += 65535) {          // gridDim.y limit
    const int nm = a->BC - m0 < 65535 ? a->BC - m0 : 65535;
    gaussian_kernel<<<dim3((unsigned)chunks, (unsigned)nm), 256, 0, (cudaStream_t)stream>>>(
        a->points + 2 * (size_t)m0, a->out + (size_t)m0 * a->H * a->W, a->H * a->W, a->W,
        1.f / (2.f * a->sigma * a->sigma), total4);
  }
  PB_LAUNCH_CHECK("gaussian_kernel");
  return PB_OK;
}

}  // extern "C"

static int peaks_common_check(const pb_peaks_args* a, const char* fn) {
  if (a == nullptr) {
    set_error("%s: null args", fn);
    return PB_ERR_INVALID;
  }
  if (a->N == 0) return PB_OK;  // empty batch: nothing to read or write
  if (a->heatmaps == nullptr || a->peaks == nullptr) {
    set_error("%s: null heatmaps/peaks", fn);
    return PB_ERR_INVALID;
  }
  if (a->N < 0 || a->C <= 0 || a->H <= 0 || a->W <= 0 || (long long)a->H * a->W >= (1ll << 31)) {
    set_error("%s: bad shape", fn);
    return PB_ERR_INVALID;
  }
  if (!is_device_ptr(a->heatmaps) || !is_device_ptr(a->peaks)) {
    set_error("%s: heatmaps/peaks must be device pointers (no CPU fallback)", fn);
    return PB_ERR_NOT_DEVICE;
  }
  return PB_OK;
}

template <typename T>
static int launch_argmax(const pb_peaks_args* a, cudaStream_t st) {
  const int maps = a->N * a->C;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(a->peaks);
  const T* hm = (const T*)a->heatmaps;
  constexpr int V = 16 / (int)sizeof(T);
  const bool v1 = getenv("POSEB200_ARGMAX_V1") != nullptr;   // A/B switch: the element-wise scan
  const bool vec = !v1 && a->stride_x == 1 && a->W % V == 0 && a->stride_y % V == 0 && a->stride_c % V == 0 &&
                   a->stride_n % V == 0 && (reinterpret_cast<uintptr_t>(hm) & 15) == 0;
  int splits = 1, rows = a->H;
  if (a->stride_x == 1) {
    const int target_ctas = sm_count() * 4;
    if (maps < target_ctas) splits = min(a->H, cdiv(target_ctas, maps));
    rows = cdiv(a->H, splits);
    splits = cdiv(a->H, rows);
  }
  if (vec && splits == 1) {
    // one CTA per map: peaks and maxima written by the scan itself, one launch
    if (a->stride_y == a->W)
      argmax_planar_vec_kernel<T, true><<<dim3(maps, 1), 256, 0, st>>>(hm, keys, a->values, a->C, a->H, a->W,
                                                                       a->stride_n, a->stride_c, a->stride_y, rows, 1);
    else
      argmax_planar_vec_kernel<T, false><<<dim3(maps, 1), 256, 0, st>>>(hm, keys, a->values, a->C, a->H, a->W,
                                                                        a->stride_n, a->stride_c, a->stride_y, rows, 1);
    PB_LAUNCH_CHECK("argmax_planar_vec_kernel");
    return PB_OK;
  }
  zero_u64_kernel<<<cdiv(maps, 256), 256, 0, st>>>(keys, maps);
  PB_LAUNCH_CHECK("zero_u64_kernel");
  if (a->stride_x == 1) {
    if (vec && a->stride_y == a->W)
      argmax_planar_vec_kernel<T, true><<<dim3(maps, splits), 256, 0, st>>>(hm, keys, nullptr, a->C, a->H, a->W,
                                                                            a->stride_n, a->stride_c, a->stride_y, rows, 0);
    else if (vec)
      argmax_planar_vec_kernel<T, false><<<dim3(maps, splits), 256, 0, st>>>(hm, keys, nullptr, a->C, a->H, a->W,
                                                                             a->stride_n, a->stride_c, a->stride_y, rows, 0);
    else
      argmax_planar_kernel<T><<<dim3(maps, splits), 256, 0, st>>>(hm, keys, a->C, a->H, a->W, a->stride_n,
                                                                  a->stride_c, a->stride_y, rows);
    PB_LAUNCH_CHECK("argmax_planar_kernel");
  } else {
    PB_REQUIRE(a->stride_c == 1 && a->C <= 1024, "pb_peaks_argmax: layout must have stride_x==1 or stride_c==1");
    const int G = max(1, 1024 / a->C >= 1 ? min(1024 / a->C, 32) : 1);
    const int threads = ((G * a->C + 31) / 32) * 32;
    const int HW = a->H * a->W;
    int isplits = max(1, min(HW / (G * 8) + 1, cdiv(sm_count() * 2, max(1, a->N))));
    const int per = cdiv(HW, isplits);
    isplits = cdiv(HW, per);
    argmax_interleaved_kernel<T><<<dim3(a->N, isplits), threads, 0, st>>>(hm, keys, a->C, HW, a->W, a->stride_n,
                                                                         a->stride_y, a->stride_x, per, G);
    PB_LAUNCH_CHECK("argmax_interleaved_kernel");
  }
  argmax_finalize_kernel<<<cdiv(maps, 256), 256, 0, st>>>(a->peaks, a->values, maps, a->W);
  PB_LAUNCH_CHECK("argmax_finalize_kernel");
  return PB_OK;
}

namespace pb {
int launch_zero_u64(unsigned long long* p, int n, cudaStream_t st) {
  zero_u64_kernel<<<cdiv(n, 256), 256, 0, st>>>(p, n);
  PB_LAUNCH_CHECK("zero_u64_kernel");
  return PB_OK;
}
int launch_argmax_finalize(float* peaks, float* values, int maps, int W, cudaStream_t st) {
  argmax_finalize_kernel<<<cdiv(maps, 256), 256, 0, st>>>(peaks, values, maps, W);
  PB_LAUNCH_CHECK("argmax_finalize_kernel");
  return PB_OK;
}
}  // namespace pb

extern "C" {

int pb_peaks_argmax(const pb_peaks_args* a, void* stream) {
  int rc = peaks_common_check(a, "pb_peaks_argmax");
  if (rc != PB_OK) return rc;
  if (a->N == 0) return PB_OK;
  return a->dtype == PB_BF16 ? launch_argmax<__nv_bfloat16>(a, (cudaStream_t)stream)
                             : launch_argmax<float>(a, (cudaStream_t)stream);
}

int pb_peaks_softargmax(const pb_peaks_args* a, void* stream) {
  int rc = peaks_common_check(a, "pb_peaks_softargmax");
  if (rc != PB_OK) return rc;
  if (a->N == 0) return PB_OK;
  PB_REQUIRE(a->H > 1 && a->W > 1, "pb_peaks_softargmax: H, W must be > 1");
  const int maps = a->N * a->C;
  cudaStream_t st = (cudaStream_t)stream;
  const int V = a->dtype == PB_BF16 ? 8 : 4;
  const bool table = getenv("POSEB200_SOFTARGMAX_V1") == nullptr &&   // A/B switch: weights computed per element
                     a->stride_x == 1 && a->W % V == 0 && a->W <= SOFT_TABLE_W && a->H <= SOFT_TABLE_W && a->stride_y % V == 0 &&
                     a->stride_c % V == 0 && a->stride_n % V == 0 &&
                     (reinterpret_cast<uintptr_t>(a->heatmaps) & 15) == 0;
#define PB_SOFT(T, TABLE) softargmax_kernel<T, TABLE><<<maps, 256, 0, st>>>( \
      (const T*)a->heatmaps, a->peaks, a->C, a->H, a->W, a->stride_n, a->stride_c, a->stride_y, a->stride_x)
  if (a->dtype == PB_BF16) { if (table) PB_SOFT(__nv_bfloat16, true); else PB_SOFT(__nv_bfloat16, false); }
  else { if (table) PB_SOFT(float, true); else PB_SOFT(float, false); }
#undef PB_SOFT
  PB_LAUNCH_CHECK("softargmax_kernel");
  return PB_OK;
}

__global__ void counter_inc_kernel(int* c) { *c += 1; }

int pb_adam_step(const pb_adam_args* a, void* stream) {
  PB_REQUIRE(a != nullptr && a->param && a->grad && a->exp_avg && a->exp_avg_sq, "pb_adam_step: null args");
  PB_REQUIRE(a->n >= 0 && (a->step >= 1 || a->step_dev != nullptr), "pb_adam_step: n >= 0 and step >= 1 (or step_dev) required");
  PB_REQUIRE_DEV(a->step_dev, "step_dev");
  PB_REQUIRE_DEV(a->lr_dev, "lr_dev");
  PB_REQUIRE_DEV(a->param, "param");
  PB_REQUIRE_DEV(a->grad, "grad");
  PB_REQUIRE((((uintptr_t)a->param | (uintptr_t)a->grad | (uintptr_t)a->exp_avg | (uintptr_t)a->exp_avg_sq) & 15) == 0,
             "pb_adam_step: buffers must be 16-byte aligned");
  // bias corrections in double like torch (python floats), then rounded once
  const int step = a->step >= 1 ? a->step : 1;
  const double bc1d = 1.0 - pow((double)a->beta1, (double)step);
  const double bc2d = 1.0 - pow((double)a->beta2, (double)step);
  if (a->n > 0) {
    adam_kernel<<<grid_for(a->n / 4 + 1, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        a->param, a->grad, a->exp_avg, a->exp_avg_sq, a->n, a->lr, a->beta1, a->beta2, a->eps, a->weight_decay,
        a->grad_scale, (float)bc1d, (float)sqrt(bc2d), a->found_inf, a->step_dev, a->lr_dev);
    PB_LAUNCH_CHECK("adam_kernel");
  }
  if (a->step_dev != nullptr && a->inc_step != 0) {
    counter_inc_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(a->step_dev);
    PB_LAUNCH_CHECK("counter_inc_kernel");
  }
  return PB_OK;
}

int pb_maxpool_lrelu_fwd(const pb_pool_fwd_args* a, void* stream) {
  PB_REQUIRE(a != nullptr && a->x && a->y, "pb_maxpool_lrelu_fwd: null args");
  PB_REQUIRE(a->H % 2 == 0 && a->W % 2 == 0 && a->N > 0 && a->C > 0, "pb_maxpool_lrelu_fwd: bad shape");
  PB_REQUIRE_DEV(a->x, "x");
  PB_REQUIRE_DEV(a->y, "y");
  const long long total = (long long)a->N * (a->H / 2) * (a->W / 2) * a->C;
  cudaStream_t st = (cudaStream_t)stream;
  if (a->act_dtype == PB_F16) {
    PB_REQUIRE((a->C & 7) == 0, "pb_maxpool_lrelu_fwd: fp16 activations need C %% 8 == 0");
    PB_REQUIRE_DEV(a->y2, "y2");
    pool_fwd_vec_kernel<true><<<grid_for(total / 8, 256, 16), 256, 0, st>>>(
        (const __nv_bfloat16*)a->x, (__nv_bfloat16*)a->y, (__nv_bfloat16*)a->y2, a->H, a->W, a->C, a->slope, total / 8);
  } else if (a->act_dtype == PB_BF16 && (a->C & 7) == 0)
    pool_fwd_vec_kernel<false><<<grid_for(total / 8, 256, 16), 256, 0, st>>>(
        (const __nv_bfloat16*)a->x, (__nv_bfloat16*)a->y, nullptr, a->H, a->W, a->C, a->slope, total / 8);
  else if (a->act_dtype == PB_BF16)
    pool_fwd_kernel<__nv_bfloat16><<<grid_for(total, 256, 16), 256, 0, st>>>(
        (const __nv_bfloat16*)a->x, (__nv_bfloat16*)a->y, a->H, a->W, a->C, a->slope, total);
  else
    pool_fwd_kernel<float><<<grid_for(total, 256, 16), 256, 0, st>>>((const float*)a->x, (float*)a->y, a->H, a->W,
                                                                    a->C, a->slope, total);
  PB_LAUNCH_CHECK("pool_fwd_kernel");
  return PB_OK;
}

int pb_maxpool_lrelu_bwd(const pb_pool_bwd_args* a, void* stream) {
  PB_REQUIRE(a != nullptr && a->x && a->gy && a->gx, "pb_maxpool_lrelu_bwd: null args");
  PB_REQUIRE(a->H % 2 == 0 && a->W % 2 == 0 && a->N > 0 && a->C > 0, "pb_maxpool_lrelu_bwd: bad shape");
  PB_REQUIRE_DEV(a->x, "x");
  PB_REQUIRE_DEV(a->gy, "gy");
  PB_REQUIRE_DEV(a->gx, "gx");
  const long long total = (long long)a->N * (a->H / 2) * (a->W / 2) * a->C;
  cudaStream_t st = (cudaStream_t)stream;
  if (a->x_dtype == PB_F16) {
    PB_REQUIRE(a->act_dtype == PB_BF16 && (a->C & 7) == 0,
               "pb_maxpool_lrelu_bwd: fp16 activations come with bf16 gradients and C %% 8 == 0");
    pool_bwd_vec_kernel<true><<<grid_for(total / 8, 256, 16), 256, 0, st>>>(
        (const __nv_bfloat16*)a->x, (const __nv_bfloat16*)a->gy, a->mask, (__nv_bfloat16*)a->gx,
        (__nv_bfloat16*)a->gx_masked, a->H, a->W, a->C, a->slope, total / 8);
  } else if (a->act_dtype == PB_BF16 && (a->C & 7) == 0)
    pool_bwd_vec_kernel<false><<<grid_for(total / 8, 256, 16), 256, 0, st>>>(
        (const __nv_bfloat16*)a->x, (const __nv_bfloat16*)a->gy, a->mask, (__nv_bfloat16*)a->gx,
        (__nv_bfloat16*)a->gx_masked, a->H, a->W, a->C, a->slope, total / 8);
  else if (a->act_dtype == PB_BF16)
    pool_bwd_kernel<__nv_bfloat16><<<grid_for(total, 256, 16), 256, 0, st>>>(
        (const __nv_bfloat16*)a->x, (const __nv_bfloat16*)a->gy, a->mask, (__nv_bfloat16*)a->gx,
        (__nv_bfloat16*)a->gx_masked, a->H, a->W, a->C, a->slope, total);
  else
    pool_bwd_kernel<float><<<grid_for(total, 256, 16), 256, 0, st>>>((const float*)a->x, (const float*)a->gy,
                                                                    a->mask, (float*)a->gx, (float*)a->gx_masked,
                                                                    a->H, a->W, a->C, a->slope, total);
  PB_LAUNCH_CHECK("pool_bwd_kernel");
  return PB_OK;
}

int pb_pack_weights(const pb_pack_weights_args* a, void* stream) {
  PB_REQUIRE(a != nullptr && a->src && a->dst, "pb_pack_weights: null args");
  PB_REQUIRE(a->ntaps >= 1 && a->ntaps <= PB_MAX_TAPS && a->I > 0 && a->Ipad >= a->I && a->J > 0 && a->Jpad >= a->J,
             "pb_pack_weights: bad shape");
  PB_REQUIRE_DEV(a->src, "src");
  PB_REQUIRE_DEV(a->dst, "dst");
  KposArr kp;
  for (int t = 0; t < PB_MAX_TAPS; ++t) kp.v[t] = a->kpos[t];
  const long long total = (long long)a->ntaps * a->Ipad * a->Jpad;
  cudaStream_t st = (cudaStream_t)stream;
  if (a->dst_dtype == PB_BF16)
    pack_weights_kernel<__nv_bfloat16><<<grid_for(total, 256, 8), 256, 0, st>>>(
        a->src, (__nv_bfloat16*)a->dst, a->ntaps, a->I, a->Ipad, a->J, a->Jpad, a->stride_i, a->stride_j, kp);
  else if (a->dst_dtype == PB_F16)
    pack_weights_kernel<__half><<<grid_for(total, 256, 8), 256, 0, st>>>(
        a->src, (__half*)a->dst, a->ntaps, a->I, a->Ipad, a->J, a->Jpad, a->stride_i, a->stride_j, kp);
  else
    pack_weights_kernel<float><<<grid_for(total, 256, 8), 256, 0, st>>>(
        a->src, (float*)a->dst, a->ntaps, a->I, a->Ipad, a->J, a->Jpad, a->stride_i, a->stride_j, kp);
  PB_LAUNCH_CHECK("pack_weights_kernel");
  return PB_OK;
}

int pb_pack_weights_multi(const pb_pack_weights_multi_args* a, void* stream) {
  PB_REQUIRE(a != nullptr && a->items != nullptr, "pb_pack_weights_multi: null args");
  PB_REQUIRE(a->count >= 0 && a->count <= 65535 && a->max_elems > 0, "pb_pack_weights_multi: bad count / max_elems");
  PB_REQUIRE_DEV(a->items, "items");
  if (a->count == 0) return PB_OK;
  long long chunks = (a->max_elems + 1023) / 1024;     // 32 x 32 tiles of the largest item
  if (chunks > 512) chunks = 512;
  if (chunks < 1) chunks = 1;
  pack_weights_multi_kernel<<<dim3((unsigned)chunks, (unsigned)a->count), 256, 0, (cudaStream_t)stream>>>(
      (const pb_pack_weights_args*)a->items);
  PB_LAUNCH_CHECK("pack_weights_multi_kernel");
  return PB_OK;
}

int pb_affine_nearest(const pb_affine_nearest_args* a, void* stream) {
  PB_REQUIRE(a != nullptr && a->in && a->out && a->theta, "pb_affine_nearest: null args");
  PB_REQUIRE(a->B >= 0 && a->C > 0 && a->H > 0 && a->W > 0 && a->B <= 65535, "pb_affine_nearest: bad shape");
  PB_REQUIRE(a->in != (const void*)a->out, "pb_affine_nearest: in-place resampling is not supported");
  PB_REQUIRE(a->in_u8 == 0 || a->in_u8 == 1, "pb_affine_nearest: in_u8 must be 0 or 1");
  PB_REQUIRE_DEV(a->in, "in");
  PB_REQUIRE_DEV(a->out, "out");
  PB_REQUIRE_DEV(a->theta, "theta");
  PB_REQUIRE_DEV(a->flips, "flips");
  PB_REQUIRE_DEV(a->src_index, "src_index");
  if (a->B == 0) return PB_OK;
  const cudaStream_t st = (cudaStream_t)stream;
  // four pixels per thread pay for byte sources (one sector = 32 pixels); for fp32 sources one pixel per thread keeps
  // a warp's gather on fewer sectors (measured: 130 vs 151 us on 64 x 36 x 192^2)
  const bool vec = a->in_u8 && (a->W & 3) == 0 && (reinterpret_cast<uintptr_t>(a->out) & 15) == 0 &&
                   getenv("POSEB200_AFFINE_PX1") == nullptr;   // A/B switch: one pixel per thread for byte sources too
  const int tw = vec ? 64 : 32, thh = vec ? 16 : 8;
  const dim3 grid((unsigned)(((a->W + tw - 1) / tw) * ((a->H + thh - 1) / thh)), (unsigned)a->B);
#define PB_AFFINE(T, PX) affine_nearest_kernel<T, PX><<<grid, 256, 0, st>>>( \
      (const T*)a->in, a->out, a->theta, a->flips, a->src_index, a->C, a->H, a->W)
  if (a->in_u8) { if (vec) PB_AFFINE(uint8_t, 4); else PB_AFFINE(uint8_t, 1); }
  else PB_AFFINE(float, 1);
#undef PB_AFFINE
  PB_LAUNCH_CHECK("affine_nearest_kernel");
  return PB_OK;
}

int pb_wgrad_reduce(const pb_wgrad_reduce_args* a, void* stream) {
  PB_REQUIRE(a != nullptr && a->partial && a->dw, "pb_wgrad_reduce: null args");
  PB_REQUIRE(a->ksplit >= 1 && a->ntaps >= 1 && a->ntaps <= PB_MAX_TAPS && a->Ca > 0 && a->Cg > 0,
             "pb_wgrad_reduce: bad shape");
  PB_REQUIRE_DEV(a->partial, "partial");
  PB_REQUIRE_DEV(a->dw, "dw");
  KposArr kp;
  for (int t = 0; t < PB_MAX_TAPS; ++t) kp.v[t] = a->kpos[t];
  const long long L = (long long)a->ntaps * a->Ca * a->Cg + a->Cg;
  if (a->ntaps == 1 && a->stride_a == 1 && a->kpos[0] == 0 && (long long)a->Ca * a->Cg >= (1 << 16)) {
    const int cav = a->Ca_valid > 0 ? a->Ca_valid : a->Ca;
    const long long tiles = (long long)((cav + 31) / 32) * ((a->Cg + 31) / 32);
    wgrad_reduce_turn_kernel<<<(unsigned)(tiles < 8 * 148 ? tiles : 8 * 148), 256, 0, (cudaStream_t)stream>>>(
        a->partial, a->dw, a->dbias, a->ksplit, a->Ca, a->Cg, a->stride_g, a->beta, a->alpha, cav);
    PB_LAUNCH_CHECK("wgrad_reduce_turn_kernel");
    return PB_OK;
  }
  wgrad_reduce_kernel<<<grid_for(L, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      a->partial, a->dw, a->dbias, a->ksplit, a->ntaps, a->Ca, a->Cg, a->stride_a, a->stride_g, kp, a->beta,
      a->alpha, a->Ca_valid > 0 ? a->Ca_valid : a->Ca);
  PB_LAUNCH_CHECK("wgrad_reduce_kernel");
  return PB_OK;
}

int pb_im2col_first(const pb_im2col_args* a, void* stream) {
  PB_REQUIRE(a != nullptr && a->in && a->out, "pb_im2col_first: null args");
  PB_REQUIRE(a->N > 0 && a->C > 0 && a->H > 0 && a->W > 0 && a->ksize >= 1 && (a->ksize & 1) && a->dilation >= 1,
             "pb_im2col_first: bad shape");
  PB_REQUIRE(a->Kpad % 8 == 0 && a->Kpad >= a->C * a->ksize * a->ksize, "pb_im2col_first: Kpad must be a multiple of 8 "
             "and >= C*k*k");
  PB_REQUIRE_DEV(a->in, "in");
  PB_REQUIRE_DEV(a->out, "out");
  const long long total = (long long)a->N * a->H * a->W * (a->Kpad / 8);
  cudaStream_t st = (cudaStream_t)stream;
  if ((a->act_dtype == PB_BF16 || a->act_dtype == PB_F16) && a->ksize == 3 && a->Kpad == 64 &&
      (a->C == 4 || a->C == 3 || a->C == 1)) {
    const long long npix = (long long)a->N * a->H * a->W;
    const int grid = grid_for(npix, 256, 8);
    __nv_bfloat16* o = (__nv_bfloat16*)a->out;
    if (a->act_dtype == PB_F16) {
      if (a->C == 4) im2col3_bf16_kernel<4, true><<<grid, 256, 0, st>>>(a->in, o, a->H, a->W, a->dilation, npix);
      else if (a->C == 3) im2col3_bf16_kernel<3, true><<<grid, 256, 0, st>>>(a->in, o, a->H, a->W, a->dilation, npix);
      else im2col3_bf16_kernel<1, true><<<grid, 256, 0, st>>>(a->in, o, a->H, a->W, a->dilation, npix);
    } else if (a->C == 4) im2col3_bf16_kernel<4, false><<<grid, 256, 0, st>>>(a->in, o, a->H, a->W, a->dilation, npix);
    else if (a->C == 3) im2col3_bf16_kernel<3, false><<<grid, 256, 0, st>>>(a->in, o, a->H, a->W, a->dilation, npix);
    else im2col3_bf16_kernel<1, false><<<grid, 256, 0, st>>>(a->in, o, a->H, a->W, a->dilation, npix);
  } else if (a->act_dtype == PB_F16)
    im2col_first_kernel<__half><<<grid_for(total, 256, 16), 256, 0, st>>>(
        a->in, (__half*)a->out, a->C, a->H, a->W, a->ksize, a->dilation, a->Kpad, total);
  else if (a->act_dtype == PB_BF16)
    im2col_first_kernel<__nv_bfloat16><<<grid_for(total, 256, 16), 256, 0, st>>>(
        a->in, (__nv_bfloat16*)a->out, a->C, a->H, a->W, a->ksize, a->dilation, a->Kpad, total);
  else
    im2col_first_kernel<float><<<grid_for(total, 256, 16), 256, 0, st>>>(a->in, (float*)a->out, a->C, a->H, a->W,
                                                                        a->ksize, a->dilation, a->Kpad, total);
  PB_LAUNCH_CHECK("im2col_first_kernel");
  return PB_OK;
}

int pb_colsum(const pb_colsum_args* a, void* stream) {
  PB_REQUIRE(a != nullptr && a->partial && a->out && a->nblk >= 1 && a->dim >= 1, "pb_colsum: bad args");
  PB_REQUIRE_DEV(a->partial, "partial");
  PB_REQUIRE_DEV(a->out, "out");
  PB_REQUIRE((a->partial2 == nullptr) == (a->out2 == nullptr), "pb_colsum: partial2 and out2 come together");
  PB_REQUIRE_DEV(a->partial2, "partial2");
  PB_REQUIRE_DEV(a->out2, "out2");
  const dim3 grid((unsigned)cdiv(a->dim, 32), a->partial2 != nullptr ? 2u : 1u);
  if (a->in_dtype == PB_BF16)
    colsum_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)a->partial, a->out, (const __nv_bfloat16*)a->partial2, a->out2, a->nblk, a->dim, a->alpha,
        a->beta);
  else
    colsum_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)a->partial, a->out,
                                                                (const float*)a->partial2, a->out2, a->nblk, a->dim,
                                                                a->alpha, a->beta);
  PB_LAUNCH_CHECK("colsum_kernel");
  return PB_OK;
}

int pb_add(const pb_add_args* a, void* stream) {
  PB_REQUIRE(a != nullptr && a->a && a->out && a->n >= 0 && (a->mask == nullptr || a->C > 0), "pb_add: bad args");
  PB_REQUIRE_DEV(a->a, "a");
  PB_REQUIRE_DEV(a->out, "out");
  if (a->n == 0) return PB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (a->act_dtype == PB_BF16)
    add_kernel<__nv_bfloat16><<<grid_for(a->n, 256, 16), 256, 0, st>>>(
        (const __nv_bfloat16*)a->a, (const __nv_bfloat16*)a->b, (__nv_bfloat16*)a->out, a->n, a->mask, a->C,
        a->slope);
  else
    add_kernel<float><<<grid_for(a->n, 256, 16), 256, 0, st>>>((const float*)a->a, (const float*)a->b,
                                                              (float*)a->out, a->n, a->mask, a->C, a->slope);
  PB_LAUNCH_CHECK("add_kernel");
  return PB_OK;
}

}  // extern "C"
