// Kernels of FourCamerasDisentanglement (pytorch/CNNs.py:240-352): FTL / InvFTL re-projections and train-mode
// BatchNorm (+ ReLU) on NHWC rows.  All HBM-bound elementwise / reduction passes over small (48 x 48) feature maps.
#include "common.cuh"

namespace pb {

// ------------------------------------------------------------------------------------------------ FTL / InvFTL
template <typename T, int KIN, int KOUT>
__global__ void __launch_bounds__(256)
ftl_kernel(const T* __restrict__ in, T* __restrict__ out, const float* __restrict__ mats, long long groups,
           long long ibs, long long obs, int in_mod, int accumulate) {
  const int b = blockIdx.y;
  const T* ip = in + (long long)(in_mod > 0 ? b % in_mod : b) * ibs;
  T* op = out + (long long)b * obs;
  float m[KOUT][KIN];
#pragma unroll
  for (int i = 0; i < KOUT; ++i)
#pragma unroll
    for (int j = 0; j < KIN; ++j) m[i][j] = __ldg(mats + ((long long)b * KOUT + i) * KIN + j);
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (long long)gridDim.x * blockDim.x) {
    float v[KIN];
#pragma unroll
    for (int j = 0; j < KIN; ++j) v[j] = ldf<T>(ip, g * KIN + j);
#pragma unroll
    for (int i = 0; i < KOUT; ++i) {
      float s = accumulate ? ldf<T>(op, g * KOUT + i) : 0.f;
#pragma unroll
      for (int j = 0; j < KIN; ++j) s = fmaf(m[i][j], v[j], s);
      stf<T>(op, g * KOUT + i, s);
    }
  }
}

template <typename T>
static int launch_ftl(const pb_ftl_args* a, cudaStream_t st) {
  const long long want = ((long long)a->groups + 255) / 256;
  const dim3 grid((unsigned)(want < 148 * 8 ? want : 148 * 8), (unsigned)a->B);
  if (a->kin == 4 && a->kout == 3)
    ftl_kernel<T, 4, 3><<<grid, 256, 0, st>>>((const T*)a->in, (T*)a->out, a->mats, a->groups, a->in_batch_stride,
                                              a->out_batch_stride, a->in_batch_mod, a->accumulate);
  else
    ftl_kernel<T, 3, 4><<<grid, 256, 0, st>>>((const T*)a->in, (T*)a->out, a->mats, a->groups, a->in_batch_stride,
                                              a->out_batch_stride, a->in_batch_mod, a->accumulate);
  PB_LAUNCH_CHECK("ftl_kernel");
  return PB_OK;
}

// ------------------------------------------------------------------------------------------------ BatchNorm
// pass 1: per (group, row block, channel) partial sums of two quantities; a thread owns channels tid, tid+256, ...
// (consecutive threads = consecutive channels of a row: coalesced)
//   forward : (x, x^2)            backward: (gy', gy' * xhat) with gy' = relu-masked gy
template <typename T, bool BWD>
__global__ void __launch_bounds__(256)
bn_partial_kernel(const T* __restrict__ x, const T* __restrict__ y, const T* __restrict__ gy,
                  const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ partial,
                  int rows_per_group, int C, int Cs, int nblk, int relu) {
  const int g = blockIdx.y, blk = blockIdx.x;
  const int rows_per_blk = (rows_per_group + nblk - 1) / nblk;
  const int r0 = blk * rows_per_blk, r1 = min(rows_per_group, r0 + rows_per_blk);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s0 = 0.f, s1 = 0.f;
    float mu = 0.f, rs = 0.f;
    if (BWD) { mu = mean[g * C + c]; rs = rstd[g * C + c]; }
    for (int r = r0; r < r1; ++r) {
      const long long e = ((long long)g * rows_per_group + r) * Cs + c;
      if (BWD) {
        float gv = ldf<T>(gy, e);
        if (relu && !(ldf<T>(y, e) > 0.f)) gv = 0.f;
        s0 += gv;
        s1 += gv * (ldf<T>(x, e) - mu) * rs;
      } else {
        const float v = ldf<T>(x, e);
        s0 += v;
        s1 += v * v;
      }
    }
    float* p = partial + (((long long)g * nblk + blk) * 2) * C;
    p[c] = s0;
    p[C + c] = s1;
  }
}

// pass 2 (forward): fold the partials in double; batch statistics; running statistics group after group
__global__ void __launch_bounds__(256)
bn_finalize_fwd_kernel(const float* __restrict__ partial, float* __restrict__ save_mean, float* __restrict__ save_rstd,
                       float* __restrict__ running_mean, float* __restrict__ running_var, int groups,
                       int rows_per_group, int C, int nblk, float eps, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float rm = running_mean != nullptr ? running_mean[c] : 0.f;
  float rv = running_var != nullptr ? running_var[c] : 1.f;
  const double n = (double)rows_per_group;
  for (int g = 0; g < groups; ++g) {
    double s0 = 0.0, s1 = 0.0;
    for (int b = 0; b < nblk; ++b) {
      const float* p = partial + (((long long)g * nblk + b) * 2) * C;
      s0 += (double)p[c];
      s1 += (double)p[C + c];
    }
    const double mean = s0 / n;
    double var = s1 / n - mean * mean;
    if (var < 0.0) var = 0.0;
    save_mean[g * C + c] = (float)mean;
    save_rstd[g * C + c] = (float)(1.0 / sqrt(var + (double)eps));
    rm = (1.f - momentum) * rm + momentum * (float)mean;
    rv = (1.f - momentum) * rv + momentum * (float)(rows_per_group > 1 ? var * n / (n - 1.0) : var);
  }
  if (running_mean != nullptr) running_mean[c] = rm;
  if (running_var != nullptr) running_var[c] = rv;
}

// eval mode: the "batch" statistics are the running ones
__global__ void __launch_bounds__(256)
bn_eval_stats_kernel(const float* __restrict__ running_mean, const float* __restrict__ running_var,
                     float* __restrict__ save_mean, float* __restrict__ save_rstd, int groups, int C, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= groups * C) return;
  const int c = i % C;
  save_mean[i] = running_mean[c];
  save_rstd[i] = rsqrtf(running_var[c] + eps);
}

// pass 3 (forward): y = relu?((x - mean) * rstd * gamma + beta); padding channels -> 0
template <typename T>
__global__ void __launch_bounds__(256)
bn_apply_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ gamma,
                const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ rstd,
                long long total, int rows_per_group, int C, int Cs, int relu) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % Cs);
    const long long row = e / Cs;
    float o = 0.f;
    if (c < C) {
      const int g = (int)(row / rows_per_group);
      o = (ldf<T>(x, e) - mean[g * C + c]) * rstd[g * C + c] * gamma[c] + beta[c];
      if (relu) o = fmaxf(o, 0.f);
    }
    stf<T>(y, e, o);
  }
}

// pass 2 (backward): fold; parameter gradients summed over the groups; per-group row means kept in `partial`'s head
__global__ void __launch_bounds__(256)
bn_finalize_bwd_kernel(float* __restrict__ partial, float* __restrict__ dgamma, float* __restrict__ dbeta, int groups,
                       int rows_per_group, int C, int nblk, float beta_acc) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double tg = 0.0, tb = 0.0;
  for (int g = 0; g < groups; ++g) {
    double s0 = 0.0, s1 = 0.0;
    for (int b = 0; b < nblk; ++b) {
      const float* p = partial + (((long long)g * nblk + b) * 2) * C;
      s0 += (double)p[c];
      s1 += (double)p[C + c];
    }
    tb += s0;
    tg += s1;
    // the group's means overwrite its block-0 slot (read by pass 3 after this kernel)
    float* p0 = partial + (((long long)g * nblk) * 2) * C;
    p0[c] = (float)(s0 / rows_per_group);
    p0[C + c] = (float)(s1 / rows_per_group);
  }
  dbeta[c] = (beta_acc != 0.f ? beta_acc * dbeta[c] : 0.f) + (float)tb;
  dgamma[c] = (beta_acc != 0.f ? beta_acc * dgamma[c] : 0.f) + (float)tg;
}

// pass 3 (backward): gx = gamma * rstd * (gy' - mean(gy') - xhat * mean(gy' * xhat))
template <typename T>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const T* __restrict__ x, const T* __restrict__ y, const T* __restrict__ gy, T* __restrict__ gx,
                    const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                    const float* __restrict__ partial, long long total, int rows_per_group, int C, int Cs, int nblk,
                    int relu) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % Cs);
    const long long row = e / Cs;
    float o = 0.f;
    if (c < C) {
      const int g = (int)(row / rows_per_group);
      const float* p0 = partial + (((long long)g * nblk) * 2) * C;
      float gv = ldf<T>(gy, e);
      if (relu && !(ldf<T>(y, e) > 0.f)) gv = 0.f;
      const float rs = rstd[g * C + c];
      const float xh = (ldf<T>(x, e) - mean[g * C + c]) * rs;
      o = gamma[c] * rs * (gv - p0[c] - xh * p0[C + c]);
    }
    stf<T>(gx, e, o);
  }
}

static int grid_1d(long long n) {
  const long long want = (n + 255) / 256;
  return (int)(want < 148 * 16 ? want : 148 * 16);
}

template <typename T>
static int bn_fwd(const pb_batchnorm_fwd_args* a, cudaStream_t st) {
  const long long total = (long long)a->groups * a->rows_per_group * a->Cs;
  if (a->training) {
    bn_partial_kernel<T, false><<<dim3(a->nblk, a->groups), 256, 0, st>>>((const T*)a->x, nullptr, nullptr, nullptr, nullptr,
                                                                         a->partial, a->rows_per_group, a->C, a->Cs,
                                                                         a->nblk, 0);
    PB_LAUNCH_CHECK("bn_partial_kernel");
    bn_finalize_fwd_kernel<<<cdiv(a->C, 256), 256, 0, st>>>(a->partial, a->save_mean, a->save_rstd, a->running_mean,
                                                           a->running_var, a->groups, a->rows_per_group, a->C, a->nblk,
                                                           a->eps, a->momentum);
    PB_LAUNCH_CHECK("bn_finalize_fwd_kernel");
  } else {
    bn_eval_stats_kernel<<<cdiv(a->groups * a->C, 256), 256, 0, st>>>(a->running_mean, a->running_var, a->save_mean,
                                                                    a->save_rstd, a->groups, a->C, a->eps);
    PB_LAUNCH_CHECK("bn_eval_stats_kernel");
  }
  bn_apply_kernel<T><<<grid_1d(total), 256, 0, st>>>((const T*)a->x, (T*)a->y, a->gamma, a->beta, a->save_mean,
                                                    a->save_rstd, total, a->rows_per_group, a->C, a->Cs, a->relu);
  PB_LAUNCH_CHECK("bn_apply_kernel");
  return PB_OK;
}

template <typename T>
static int bn_bwd(const pb_batchnorm_bwd_args* a, cudaStream_t st) {
  const long long total = (long long)a->groups * a->rows_per_group * a->Cs;
  bn_partial_kernel<T, true><<<dim3(a->nblk, a->groups), 256, 0, st>>>((const T*)a->x, (const T*)a->y, (const T*)a->gy,
                                                                      a->save_mean, a->save_rstd, a->partial,
                                                                      a->rows_per_group, a->C, a->Cs, a->nblk, a->relu);
  PB_LAUNCH_CHECK("bn_partial_kernel");
  bn_finalize_bwd_kernel<<<cdiv(a->C, 256), 256, 0, st>>>(a->partial, a->dgamma, a->dbeta, a->groups, a->rows_per_group,
                                                         a->C, a->nblk, a->beta_acc);
  PB_LAUNCH_CHECK("bn_finalize_bwd_kernel");
  bn_bwd_apply_kernel<T><<<grid_1d(total), 256, 0, st>>>((const T*)a->x, (const T*)a->y, (const T*)a->gy, (T*)a->gx,
                                                        a->gamma, a->save_mean, a->save_rstd, a->partial, total,
                                                        a->rows_per_group, a->C, a->Cs, a->nblk, a->relu);
  PB_LAUNCH_CHECK("bn_bwd_apply_kernel");
  return PB_OK;
}

}  // namespace pb

using namespace pb;

extern "C" {

int pb_ftl(const pb_ftl_args* a, void* stream) {
  PB_REQUIRE(a != nullptr && a->in && a->out && a->mats, "pb_ftl: null args");
  PB_REQUIRE(a->B > 0 && a->groups > 0 && ((a->kin == 4 && a->kout == 3) || (a->kin == 3 && a->kout == 4)),
             "pb_ftl: (kin, kout) must be (4, 3) [FTL] or (3, 4) [InvFTL]");
  PB_REQUIRE_DEV(a->in, "in");
  PB_REQUIRE_DEV(a->out, "out");
  PB_REQUIRE_DEV(a->mats, "mats");
  cudaStream_t st = (cudaStream_t)stream;
  if (a->act_dtype == PB_BF16) return launch_ftl<__nv_bfloat16>(a, st);
  if (a->act_dtype == PB_F16) return launch_ftl<__half>(a, st);
  return launch_ftl<float>(a, st);
}

int pb_batchnorm_fwd(const pb_batchnorm_fwd_args* a, void* stream) {
  PB_REQUIRE(a != nullptr && a->x && a->y && a->gamma && a->beta && a->save_mean && a->save_rstd, "pb_batchnorm_fwd: null args");
  PB_REQUIRE(a->groups > 0 && a->rows_per_group > 0 && a->C > 0 && a->Cs >= a->C && a->nblk >= 1, "pb_batchnorm_fwd: shape");
  PB_REQUIRE(a->training ? a->partial != nullptr : (a->running_mean != nullptr && a->running_var != nullptr),
             "pb_batchnorm_fwd: training needs `partial`, eval needs the running statistics");
  PB_REQUIRE_DEV(a->x, "x");
  PB_REQUIRE_DEV(a->y, "y");
  cudaStream_t st = (cudaStream_t)stream;
  if (a->act_dtype == PB_BF16) return bn_fwd<__nv_bfloat16>(a, st);
  if (a->act_dtype == PB_F16) return bn_fwd<__half>(a, st);
  return bn_fwd<float>(a, st);
}

int pb_batchnorm_bwd(const pb_batchnorm_bwd_args* a, void* stream) {
  PB_REQUIRE(a != nullptr && a->x && a->gy && a->gx && a->gamma && a->save_mean && a->save_rstd && a->dgamma && a->dbeta &&
                 a->partial,
             "pb_batchnorm_bwd: null args");
  PB_REQUIRE(!a->relu || a->y != nullptr, "pb_batchnorm_bwd: relu needs the forward output y");
  PB_REQUIRE(a->groups > 0 && a->rows_per_group > 0 && a->C > 0 && a->Cs >= a->C && a->nblk >= 1, "pb_batchnorm_bwd: shape");
  PB_REQUIRE_DEV(a->x, "x");
  PB_REQUIRE_DEV(a->gy, "gy");
  PB_REQUIRE_DEV(a->gx, "gx");
  cudaStream_t st = (cudaStream_t)stream;
  if (a->act_dtype == PB_BF16) return bn_bwd<__nv_bfloat16>(a, st);
  if (a->act_dtype == PB_F16) return bn_bwd<__half>(a, st);
  return bn_bwd<float>(a, st);
}

}  // extern "C"
