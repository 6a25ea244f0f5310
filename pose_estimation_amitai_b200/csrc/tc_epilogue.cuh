// Fused epilogue shared by the tcgen05 contraction kernels: bias / skip add / pre-activation store /
// LeakyReLU (+sign mask) / LeakyReLU' / GELU / residual, on CH consecutive channels of one pixel.
#pragma once

#include "tc_common.cuh"

namespace pb {

struct EpiP {
  int Cout;
  const float* bias;
  const __nv_bfloat16* add0;
  const __nv_bfloat16* add1;
  __nv_bfloat16* pre_out;
  void* out;
  __nv_bfloat16* out2;   // bf16 twin of an fp16 `out` (NHWC), or nullptr
  uint32_t* mask_out;
  const uint32_t* mask_in;
  int act;
  float slope;
};

// 16-bit element <-> float in the storage format of the call (F16: IEEE half, else bf16); EpiP's pointers are typed
// bf16 for their 2-byte stride only
template <bool F16>
__device__ __forceinline__ float ld16(const __nv_bfloat16* p) {
  const uint32_t u = *reinterpret_cast<const unsigned short*>(p);
  return F16 ? f16lo(u) : __uint_as_float(u << 16);
}
template <bool F16>
__device__ __forceinline__ void st16(__nv_bfloat16* p, float v) {
  if (F16) *reinterpret_cast<__half*>(p) = __float2half_rn(v);
  else *p = __float2bfloat16_rn(v);
}

// epilogue on CH consecutive channels (c .. c+CH-1) of one output pixel
template <int CH, bool F16 = false>
__device__ __forceinline__ void epilogue_chunk(const EpiP& p, uint32_t (&r)[CH], long long pix, int c,
                                               bool pixel_ok) {
  if (!pixel_ok || c >= p.Cout) return;
  const int words = (p.Cout + 31) >> 5;
  const bool full = (c + CH <= p.Cout) && ((p.Cout & 7) == 0);
  float v[CH];
#pragma unroll
  for (int j = 0; j < CH; ++j) v[j] = __uint_as_float(r[j]);
  if (p.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < CH; ++j)
      if (c + j < p.Cout) v[j] += __ldg(p.bias + c + j);
  }
  const long long base = pix * p.Cout + c;
  if (p.add0 != nullptr) {
    if (full) {
#pragma unroll
      for (int q = 0; q < CH / 8; ++q) {
        const uint4 t = *reinterpret_cast<const uint4*>(p.add0 + base + q * 8);
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          v[q * 8 + 2 * e] += lo16<F16>(w[e]);
          v[q * 8 + 2 * e + 1] += hi16<F16>(w[e]);
        }
      }
    } else {
      for (int j = 0; j < CH; ++j)
        if (c + j < p.Cout) v[j] += ld16<F16>(p.add0 + base + j);
    }
  }
  if (p.pre_out != nullptr) {
    if (full) {
#pragma unroll
      for (int q = 0; q < CH / 8; ++q) {
        uint4 t;
        t.x = pack16x2<F16>(v[q * 8 + 0], v[q * 8 + 1]);
        t.y = pack16x2<F16>(v[q * 8 + 2], v[q * 8 + 3]);
        t.z = pack16x2<F16>(v[q * 8 + 4], v[q * 8 + 5]);
        t.w = pack16x2<F16>(v[q * 8 + 6], v[q * 8 + 7]);
        *reinterpret_cast<uint4*>(p.pre_out + base + q * 8) = t;
      }
    } else {
      for (int j = 0; j < CH; ++j)
        if (c + j < p.Cout) st16<F16>(p.pre_out + base + j, v[j]);
    }
  }
  if (p.act == PB_ACT_LRELU) {
    uint32_t bits = 0;
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      bits |= (v[j] > 0.f ? 1u : 0u) << j;
      v[j] = v[j] > 0.f ? v[j] : p.slope * v[j];
    }
    if (p.mask_out != nullptr) {
      // CH == 32: one whole word; CH == 16: the low or high half of a word owned by this thread
      if (CH == 32) p.mask_out[pix * words + (c >> 5)] = bits;
      else reinterpret_cast<uint16_t*>(p.mask_out + pix * words + (c >> 5))[(c >> 4) & 1] = (uint16_t)bits;
    }
  } else if (p.act == PB_ACT_MASKMUL) {
    uint32_t bits = p.mask_in[pix * words + (c >> 5)];
    if (CH == 16) bits >>= (c & 16);
#pragma unroll
    for (int j = 0; j < CH; ++j) v[j] *= ((bits >> j) & 1u) ? 1.f : p.slope;
  } else if (p.act == PB_ACT_GELU) {
#pragma unroll
    for (int j = 0; j < CH; ++j) v[j] = 0.5f * v[j] * (1.f + erff(v[j] * 0.70710678118654752440f));
  }
  if (p.add1 != nullptr) {
    if (full) {
#pragma unroll
      for (int q = 0; q < CH / 8; ++q) {
        const uint4 t = *reinterpret_cast<const uint4*>(p.add1 + base + q * 8);
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          v[q * 8 + 2 * e] += lo16<F16>(w[e]);
          v[q * 8 + 2 * e + 1] += hi16<F16>(w[e]);
        }
      }
    } else {
      for (int j = 0; j < CH; ++j)
        if (c + j < p.Cout) v[j] += ld16<F16>(p.add1 + base + j);
    }
  }
#pragma unroll
  for (int j = 0; j < CH; ++j) r[j] = __float_as_uint(v[j]);
}

template <int CH, bool F16 = false>
__device__ __forceinline__ void store_nhwc(const EpiP& p, const uint32_t (&r)[CH], long long pix, int c,
                                           bool pixel_ok) {
  if (!pixel_ok || c >= p.Cout) return;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
  const long long base = pix * p.Cout + c;
  if ((c + CH <= p.Cout) && ((p.Cout & 7) == 0)) {
#pragma unroll
    for (int q = 0; q < CH / 8; ++q) {
      uint4 t;
      t.x = pack16x2<F16>(__uint_as_float(r[q * 8 + 0]), __uint_as_float(r[q * 8 + 1]));
      t.y = pack16x2<F16>(__uint_as_float(r[q * 8 + 2]), __uint_as_float(r[q * 8 + 3]));
      t.z = pack16x2<F16>(__uint_as_float(r[q * 8 + 4]), __uint_as_float(r[q * 8 + 5]));
      t.w = pack16x2<F16>(__uint_as_float(r[q * 8 + 6]), __uint_as_float(r[q * 8 + 7]));
      *reinterpret_cast<uint4*>(out + base + q * 8) = t;
    }
  } else {
    for (int j = 0; j < CH; ++j)
      if (c + j < p.Cout) st16<F16>(out + base + j, __uint_as_float(r[j]));
  }
  if (F16 && p.out2 != nullptr) {
    for (int j = 0; j < CH; ++j)
      if (c + j < p.Cout) p.out2[base + j] = __float2bfloat16_rn(__uint_as_float(r[j]));
  }
}


// ---- fast path: 32 channels of one pixel, Cout % 8 == 0, chunk entirely inside Cout ---------------
// Operands that live in global memory (skip / residual tensors, LeakyReLU' mask word) are fetched one
// chunk AHEAD into registers (EpiPre) so their latency overlaps the previous chunk's math; the bias
// comes from a shared-memory copy made once per CTA.
struct EpiPre {
  uint4 a0[4], a1[4];
  uint32_t m;
  long long pix;
  uint32_t col;     // TMEM column of the chunk
  int c0, width;    // first channel, 32 or 16
  bool ok, fast;
};

__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

__device__ __forceinline__ void epi_prefetch(const EpiP& p, EpiPre& e) {
  e.fast = e.ok && e.width == 32 && (e.c0 + 32 <= p.Cout) && ((p.Cout & 15) == 0);   // 32-byte aligned rows (256-bit stores)
  if (!e.fast) return;
  const long long base = e.pix * p.Cout + e.c0;
  if (p.add0 != nullptr) {
#pragma unroll
    for (int q = 0; q < 4; ++q) e.a0[q] = ldg16(p.add0 + base + q * 8);
  }
  if (p.add1 != nullptr) {
#pragma unroll
    for (int q = 0; q < 4; ++q) e.a1[q] = ldg16(p.add1 + base + q * 8);
  }
  if (p.act == PB_ACT_MASKMUL) e.m = __ldg(p.mask_in + e.pix * ((p.Cout + 31) >> 5) + (e.c0 >> 5));
}

template <bool F16>
__device__ __forceinline__ void add16x8(float* v, const uint4& t) {
  const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    v[2 * k] += lo16<F16>(w[k]);
    v[2 * k + 1] += hi16<F16>(w[k]);
  }
}
__device__ __forceinline__ void add_bf16x8(float* v, const uint4& t) { add16x8<false>(v, t); }

template <bool F16>
__device__ __forceinline__ uint4 pack16x8(const float* v) {
  uint4 t;
  t.x = pack16x2<F16>(v[0], v[1]); t.y = pack16x2<F16>(v[2], v[3]);
  t.z = pack16x2<F16>(v[4], v[5]); t.w = pack16x2<F16>(v[6], v[7]);
  return t;
}
__device__ __forceinline__ uint4 pack_bf16x8(const float* v) { return pack16x8<false>(v); }

// sbias: shared-memory bias, zero where the layer has none
// 32 bytes per instruction: a full sector per lane (16-byte stores at a >= 128-byte lane stride fill every sector in
// two partial writes)
__device__ __forceinline__ void st_global_256(void* dst, const uint4& lo, const uint4& hi) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(lo.x), "r"(lo.y), "r"(lo.z),
               "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w)
               : "memory");
}

// v[32] already holds accumulator + bias
template <bool F16 = false>
__device__ __forceinline__ void epi32_tail(const EpiP& p, float (&v)[32], const EpiPre& e) {
  const long long base = e.pix * p.Cout + e.c0;
  if (p.add0 != nullptr) {
#pragma unroll
    for (int q = 0; q < 4; ++q) add16x8<F16>(v + 8 * q, e.a0[q]);
  }
  if (p.pre_out != nullptr) {
#pragma unroll
    for (int q = 0; q < 2; ++q) st_global_256(p.pre_out + base + q * 16, pack16x8<F16>(v + 16 * q), pack16x8<F16>(v + 16 * q + 8));
  }
  if (p.act == PB_ACT_LRELU) {
    uint32_t bits = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      bits |= (v[j] > 0.f ? 1u : 0u) << j;
      v[j] = v[j] > 0.f ? v[j] : p.slope * v[j];
    }
    if (p.mask_out != nullptr) p.mask_out[e.pix * ((p.Cout + 31) >> 5) + (e.c0 >> 5)] = bits;
  } else if (p.act == PB_ACT_MASKMUL) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= ((e.m >> j) & 1u) ? 1.f : p.slope;
  } else if (p.act == PB_ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.5f * v[j] * (1.f + erff(v[j] * 0.70710678118654752440f));
  }
  if (p.add1 != nullptr) {
#pragma unroll
    for (int q = 0; q < 4; ++q) add16x8<F16>(v + 8 * q, e.a1[q]);
  }
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
#pragma unroll
  for (int q = 0; q < 2; ++q) st_global_256(out + base + q * 16, pack16x8<F16>(v + 16 * q), pack16x8<F16>(v + 16 * q + 8));
  if (F16 && p.out2 != nullptr) {
#pragma unroll
    for (int q = 0; q < 2; ++q)
      st_global_256(p.out2 + base + q * 16, pack16x8<false>(v + 16 * q), pack16x8<false>(v + 16 * q + 8));
  }
}

template <bool F16 = false>
__device__ __forceinline__ void epi32_fast(const EpiP& p, const float* sbias, const uint32_t (&r)[32],
                                           const EpiPre& e) {
  float v[32];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 b = *reinterpret_cast<const float4*>(sbias + e.c0 + 4 * q);
    v[4 * q + 0] = __uint_as_float(r[4 * q + 0]) + b.x;
    v[4 * q + 1] = __uint_as_float(r[4 * q + 1]) + b.y;
    v[4 * q + 2] = __uint_as_float(r[4 * q + 2]) + b.z;
    v[4 * q + 3] = __uint_as_float(r[4 * q + 3]) + b.w;
  }
  epi32_tail<F16>(p, v, e);
}

// same with the bias read from global memory (any alignment: parameters are views into a flat buffer); every lane
// reads the same addresses, so each load is one broadcast wavefront
template <bool F16 = false>
__device__ __forceinline__ void epi32_fast_gbias(const EpiP& p, const uint32_t (&r)[32], const EpiPre& e) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  if (p.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += __ldg(p.bias + e.c0 + j);
  }
  epi32_tail<F16>(p, v, e);
}

}  // namespace pb
