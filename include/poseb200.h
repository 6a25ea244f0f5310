/*
 * poseb200.h -- C ABI of libposeb200.so: the sm_100a implementation of the pose-estimation
 * heatmap-regression hot path (forward / backward of the conv and ViT heatmap networks,
 * Gaussian targets, MSE loss + gradient, per-joint peak extraction, fused Adam).
 *
 * The reference (lior-kotlar/pose-estimation-amitai) has NO native layer: every entry point
 * below replaces an implicit PyTorch operator dispatch (cuDNN / cuBLAS / ATen) made from the
 * reference call site cited next to it.  The Python binding is ctypes
 * (pose_estimation_amitai_b200/_lib.py); INTEGRATION.md shows the stub a reference maintainer
 * would add.
 *
 * Conventions
 *   - every function:  int pb_<op>(const pb_<op>_args* args, void* cuda_stream)
 *     returns 0 on success, a negative pb_status otherwise; the message is available through
 *     pb_last_error_string() (thread local).  No exceptions cross the ABI.
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless the field name
 *     starts with host_.  The caller owns every buffer including workspaces; the library never
 *     allocates on the hot path.  Launches are stream-ordered, re-entrant, with no hidden sync.
 *   - activations are NHWC ("pixels x channels"); pb_dtype says fp32 or bf16.  Network input
 *     crops are NCHW fp32 and output heatmaps are NCHW fp32, exactly the reference's tensors
 *     (pytorch/CNNs.py:183-186).
 *   - there is no CPU fallback: host pointers where device pointers are expected are an error
 *     (PB_ERR_NOT_DEVICE), and an absent GPU is PB_ERR_CUDA.
 */
#ifndef POSEB200_H
#define POSEB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB_ABI_VERSION 2
#define PB_MAX_TAPS 16

typedef enum {
  PB_OK = 0,
  PB_ERR_INVALID = -1,     /* bad argument / unsupported shape */
  PB_ERR_CUDA = -2,        /* CUDA runtime or driver error */
  PB_ERR_NOT_DEVICE = -3,  /* a host pointer was passed where device memory is required */
  PB_ERR_UNSUPPORTED = -4  /* shape outside what the tcgen05 path tiles (caller may use the simt op) */
} pb_status;

/* PB_F16 (IEEE half): the "fp16" precision -- FORWARD operands (activations, packed forward weights) carry 11
 * significand bits instead of bf16's 8, the dtype the reference itself trains in under torch.cuda.amp autocast
 * (pytorch/train_pytorch.py:115,133).  Gradients stay bf16 (fp32 exponent range, so no loss scaling is needed).
 * tcgen05.mma kind::f16 wants A and B in ONE format (fp16 x bf16 is an illegal instruction on sm_100a -- measured
 * with pb_gemm_selftest), so in training the forward epilogues also store a bf16 twin of every activation
 * (pb_conv_args.out2) and the weight gradients contract bf16 x bf16 as in the bf16 precision. */
typedef enum { PB_F32 = 0, PB_BF16 = 1, PB_F16 = 2 } pb_dtype;

/* epilogue activation of a contraction */
typedef enum {
  PB_ACT_NONE = 0,
  PB_ACT_LRELU = 1,   /* v = v > 0 ? v : slope * v; optionally records the sign bit (mask_out) */
  PB_ACT_MASKMUL = 2, /* v *= mask_in bit ? 1 : slope  (LeakyReLU backward) */
  PB_ACT_GELU = 3     /* exact erf GELU (pytorch_vit_encoder.py:21) */
} pb_act;

const char* pb_last_error_string(void);
int pb_abi_version(void);
/* number of CUDA kernels this library has launched in this process (bench.py: gpu_launches) */
int pb_launch_count(unsigned long long* out);
/* fills name[<=len] with the device name, sm = 10*major+minor, n_sm = SM count */
int pb_device_info(char* name, int len, int* sm, int* n_sm);

/* ------------------------------------------------------------------------------------------
 * Gather-convolution: one descriptor covers every contraction of the conv networks.
 *
 *   out[n,oy,ox,co] = sum_t sum_ci in[n, iy, ix, ci] * w[t][ci][co]
 *   iy = (oy*out_mul + dy[t]) / in_div   (tap skipped unless divisible and 0 <= iy < IH), same in x
 *
 *   nn.Conv2d(k3, dilation d, padding d)      pytorch/CNNs.py:45-49   out_mul 1 in_div 1 dy=d(r-1)
 *   nn.ConvTranspose2d(k3,s1,p1)              pytorch/CNNs.py:113-122 out_mul 1 in_div 1 dy=1-r
 *   nn.ConvTranspose2d(k3,s2,p1,op1)          pytorch/CNNs.py:108-110,125-128; VITs.py:23-34
 *                                                                      out_mul 1 in_div 2 dy=1-r
 *   their input gradients (autograd of the above; pytorch/train_pytorch.py:137) are the same
 *   form with negated offsets / swapped mul,div and the channel roles exchanged.
 *   nn.Linear (pytorch_vit_encoder.py:20-23,52,55,122) is the 1-tap case on a 1 x rows image.
 *
 * Epilogue (in this order, every piece optional):
 *   v = acc + bias[co] + add0[idx];  pre_out[idx] = v;  v = act(v);  v += add1[idx];  out[idx] = v
 * which is bias+LeakyReLU+residual in the forward (CNNs.py:74-86,152-155) and
 * skip-gradient add + LeakyReLU-backward in the input-gradient pass.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int32_t ntaps;
  int32_t out_mul, in_div;
  int8_t dy[PB_MAX_TAPS], dx[PB_MAX_TAPS];
} pb_taps;

typedef struct {
  const void* in;        /* [N, IH, IW, Cin]  act_dtype */
  const void* w;         /* simt: fp32 [ntaps][Cin][Cout]; tc: act_dtype (bf16 / fp16) [ntaps][CoutPad][Cin] */
  const float* bias;     /* [Cout] or NULL */
  const void* add0;      /* [N,OH,OW,Cout] act_dtype or NULL */
  const void* add1;      /* [N,OH,OW,Cout] act_dtype or NULL */
  void* pre_out;         /* [N,OH,OW,Cout] act_dtype or NULL */
  void* out;             /* [N,OH,OW,Cout] act_dtype, or NCHW fp32 when out_nchw_f32 */
  void* out2;            /* optional bf16 twin of `out` (act_dtype PB_F16, NHWC only): the copy the weight gradient reads */
  uint32_t* mask_out;    /* [N*OH*OW][ceil(Cout/32)] sign bits of the pre-activation, or NULL */
  const uint32_t* mask_in;
  int32_t N, IH, IW, Cin, OH, OW, Cout;
  int32_t act;           /* pb_act */
  float slope;
  int32_t act_dtype;     /* pb_dtype of in/add0/add1/pre_out/out */
  int32_t out_nchw_f32;  /* 1: out is [N,Cout,OH,OW] fp32 (network heatmap output) */
  int32_t in_nchw_f32;   /* 1: in is [N,Cin,IH,IW] fp32 (network crop input; simt only) */
  pb_taps taps;
  /* fused 2x2/2 max-pool + LeakyReLU behind the epilogue (pytorch/CNNs.py:77,82: x = leakyrelu(maxpool(x3))), tensor-core
   * path, stride-1 layers with Cout % 64 == 0 and even OH, OW: pool_out[n, oy/2, ox/2, co] = lrelu(max of the 2x2 block
   * of `out`), computed from the 16-bit values `out` holds.  pool_only != 0: `out` itself is not written (inference;
   * pass any device pointer). */
  void* pool_out;        /* [N, OH/2, OW/2, Cout] act_dtype or NULL */
  int32_t pool_only;
} pb_conv_args;

/* CUDA-core fp32-accumulate implementation ("fp32 mode" and odd shapes such as Cin=4) */
int pb_conv_simt(const pb_conv_args* a, void* stream);
/* tcgen05 / TMEM / TMA implementation (bf16 operands, fp32 accumulate) */
int pb_conv_tc(const pb_conv_args* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused network head: the LAST layer of the conv heatmap networks -- nn.ConvTranspose2d(k3, s2, p1,
 * op1) + LeakyReLU, pytorch/CNNs.py:125-128,155 -- whose epilogue consumes the heatmaps on chip instead of storing
 * them.  `conv` describes the layer exactly as for pb_conv_tc (bias + LeakyReLU only; conv.out / out_nchw_f32 are
 * ignored, conv.out may be NULL).
 *
 * pb_convT_argmax_fused   inference: replaces the heatmap D2H copy + host transpose + arg-max of
 *     Trainer.find_points / get_points_from_confmaps (pytorch/train_pytorch.py:199-213,327-331) and the heatmap
 *     tensor itself: only peaks[N][Cout][2] = (x = col, y = row) (and optionally the maxima) leave the SMs.
 *     Bit-identical to pb_peaks_argmax on the materialised heatmaps (lowest flat index on ties, NaN greatest).
 * pb_convT_mse_fused      training: replaces torch.nn.MSELoss + the start of autograd's backward
 *     (pytorch/train_pytorch.py:110,134-137): loss_sum[0] += sum (o - t)^2 and
 *     grad_nhwc[n,oy,ox,c] = (o - t) * grad_scale * (o > 0 ? 1 : slope)  (bf16, channel-padded to Cpad, zero padding)
 *     with the target read from `target` (NCHW fp32) or rendered from `points` (Gaussian, sigma) -- the operand the
 *     head's own weight- and input-gradient contractions consume; the fp32 heatmaps are never written or re-read.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  pb_conv_args conv;
  float* peaks;             /* argmax: [N][Cout][2] fp32; doubles as the 8-byte-per-map key scratch */
  float* values;            /* argmax: [N][Cout] maxima, or NULL */
  const float* target;      /* mse: [N,Cout,OH,OW] fp32, or NULL with points != NULL */
  const float* points;      /* mse: [N,Cout,2] (x,y) */
  float sigma;
  float* loss_sum;          /* mse: 1 float, caller zeroes it */
  void* grad_nhwc;          /* mse: [N,OH,OW,Cpad] bf16 */
  int32_t Cpad;             /* Cout rounded up to 16 */
  float grad_scale;         /* 2*scale/(numel*accumulation_steps) */
  float* dbias;             /* mse, optional: [dbias_rows][Cout] fp32, caller zeroes it; CTA b adds the sum over ITS pixels of
                               the (unrounded) gradient to row b (fixed summation order: deterministic), so the head's
                               bias gradient is the column sum (pb_colsum) and no separate pass over grad_nhwc is
                               needed.  dbias_rows >= the device's SM count.  Only the folded-parity kernel
                               (csrc/tc_head.cu) provides it: PB_ERR_UNSUPPORTED otherwise */
  int32_t dbias_rows;
} pb_head_fused_args;
int pb_convT_argmax_fused(const pb_head_fused_args* a, void* stream);
int pb_convT_mse_fused(const pb_head_fused_args* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * Weight gradient of a gather-convolution (autograd of the layers above):
 *   dw[t][ci][co] = sum_{n,py,px} a[n, py*mul_a+dya[t], px*mul_a+dxa[t], ci]
 *                               * g[n, py*mul_g+dyg[t], px*mul_g+dxg[t], co]
 *   dbias[co]     = sum g
 * The contraction writes fp32 partial sums (split over pixels) into `partial`
 * ([ksplit][ntaps*Ca*Cg + Cg] floats); pb_wgrad_reduce folds them into the parameter-shaped
 * gradient tensors.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  const void* a;         /* layer input  [N, AH, AW, Ca] act_dtype */
  const void* g;         /* grad wrt pre-activation [N, GH, GW, Cg] act_dtype */
  float* partial;        /* [ksplit][ntaps*Ca*Cg + Cg] */
  int32_t N, PH, PW;     /* base pixel grid the sum runs over */
  int32_t AH, AW, Ca, GH, GW, Cg;
  int32_t mul_a, mul_g;
  int32_t ntaps;
  int8_t dya[PB_MAX_TAPS], dxa[PB_MAX_TAPS], dyg[PB_MAX_TAPS], dxg[PB_MAX_TAPS];
  int32_t ksplit;
  int32_t act_dtype;
  int32_t a_nchw_f32;    /* 1: a is the NCHW fp32 network input (first layer) */
  int32_t want_bias;
  int32_t g_cstride;     /* channel extent of g in memory (>= Cg, channel-padded tensors); 0 means Cg */
} pb_wgrad_args;

int pb_wgrad_simt(const pb_wgrad_args* a, void* stream);
int pb_wgrad_tc(const pb_wgrad_args* a, void* stream);

typedef struct {
  const float* partial;  /* [ksplit][ntaps*Ca*Cg + Cg] */
  float* dw;             /* parameter-shaped: element (t,ci,co) lives at ci*stride_a + co*stride_g + kpos[t] */
  float* dbias;          /* [Cg] or NULL */
  int32_t ksplit, ntaps, Ca, Cg;
  int64_t stride_a, stride_g;
  int32_t kpos[PB_MAX_TAPS];
  float beta;            /* dw = beta*dw + sum(partials): 0 overwrite, 1 accumulate (accumulation_steps) */
  float alpha;           /* scale applied to the summed partials (1/loss-scale, 1/world) */
  int32_t Ca_valid;      /* rows ci >= Ca_valid are channel padding and are skipped; 0 means Ca */
} pb_wgrad_reduce_args;

int pb_wgrad_reduce(const pb_wgrad_reduce_args* a, void* stream);

/* First-layer im2col: the network input is NCHW fp32 with Cin = 4 (pytorch/preprocessor.py:33-39),
 * far below the 64-channel K chunk of the tensor-core kernels.  out[n,y,x,k] with k = ci*K*K + r*K + s
 * holds in[n, ci, y + d*(r-c), x + d*(s-c)] (zero outside the image; k >= Cin*K*K is zero padding), so
 * conv1 (pytorch/CNNs.py:24) becomes a 1-tap contraction whose weight matrix is conv1.weight viewed as
 * [Cout][Cin*K*K] -- and so does its weight gradient. */
typedef struct {
  const float* in;       /* [N, C, H, W] fp32 */
  void* out;             /* [N, H, W, Kpad] act_dtype */
  int32_t N, C, H, W, ksize, dilation, Kpad;
  int32_t act_dtype;
} pb_im2col_args;
int pb_im2col_first(const pb_im2col_args* a, void* stream);

/* First layer forward WITHOUT the im2col tensor (csrc/tc_conv1.cu): out = LeakyReLU(conv(in) + bias) read straight
 * from the NCHW fp32 crops -- pytorch/CNNs.py:24,74 (x1 = leakyrelu(conv1(x))).  The kernel's producer warps build
 * each pixel's K-major operand row (k = ci*9 + r*3 + s, zero padded to 64) in shared memory.  w is the same packed
 * operand the 1-tap form uses: [Cout][64] act_dtype, conv1.weight viewed as [Cout][Cin*9].  Inference and the bf16
 * training forward use it (the weight gradient still contracts over the im2col tensor). */
typedef struct {
  const float* in;       /* [N, C, H, W] fp32 */
  const void* w;         /* [Cout][64] act_dtype */
  const float* bias;     /* [Cout] or NULL */
  void* out;             /* [N, H, W, Cout] act_dtype */
  uint32_t* mask_out;    /* [N*H*W][Cout/32] sign bits of the pre-activation, or NULL */
  int32_t N, C, H, W, ksize, dilation, Cout;
  float slope;
  int32_t act_dtype;
  void* out2;            /* optional bf16 twin of an fp16 `out` (the operand conv2's weight gradient reads), or NULL */
} pb_conv_first_args;
int pb_conv_first_tc(const pb_conv_first_args* a, void* stream);

/* First layer weight + bias gradient WITHOUT the im2col tensor (csrc/tc_wgrad1.cu): autograd of conv1
 * (pytorch/CNNs.py:24, train_pytorch.py:137).  partial[split][k * Cg + co] (k = ci*9 + r*3 + s < Ca, rows >= C*9 are not
 * written) and the bias row at partial[split][Ca*Cg + co], in pb_wgrad_reduce's layout for a 1-tap contraction with
 * Ca stored / C*9 valid rows; one persistent CTA per split (ksplit <= number of SMs). */
typedef struct {
  const float* in;       /* [N, C, H, W] fp32 crops */
  const void* g;         /* [N, H, W, Cg] bf16: gradient w.r.t. conv1's pre-activation */
  float* partial;        /* [ksplit][Ca*Cg + Cg] fp32 */
  int32_t N, C, H, W, ksize, dilation, Cg, Ca, ksplit;
  int32_t act_dtype;     /* PB_BF16 */
} pb_wgrad_first_args;
int pb_wgrad_first_tc(const pb_wgrad_first_args* a, void* stream);

/* parameter tensor -> packed operand:  dst[t][i][j] = src[i*stride_i + j*stride_j + kpos[t]]
 * (rows i >= I are written as zeros up to Ipad, columns j >= J as zeros up to Jpad). */
typedef struct {
  const float* src;
  void* dst;
  int32_t ntaps, I, Ipad, J, Jpad;
  int64_t stride_i, stride_j;
  int32_t kpos[PB_MAX_TAPS];
  int32_t dst_dtype;
} pb_pack_weights_args;

int pb_pack_weights(const pb_pack_weights_args* a, void* stream);

/* The same for `count` parameter tensors in ONE launch (after every optimiser step the packed operands of all
 * layers are refreshed; 25 five-microsecond launches otherwise sit between the convolutions).  `items` is a
 * DEVICE array of pb_pack_weights_args whose src/dst pointers stay valid across steps (flat parameter buffer,
 * persistent packed tensors); `max_elems` is the largest ntaps*Ipad*Jpad among them. */
typedef struct {
  const void* items;
  int32_t count;
  int64_t max_elems;
} pb_pack_weights_multi_args;
int pb_pack_weights_multi(const pb_pack_weights_multi_args* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * Input-pipeline augmentation on the device: DefaultDataset.augment_view,
 * pytorch/Datagenerators.py:153-186 = torchvision F.affine(img, angle, translate, scale, shear=0)
 * (nearest neighbour, zero fill) followed by optional hflip / vflip, applied with the same
 * per-sample parameters to the crops and to their confidence maps.
 *   theta[b][6]: the INVERSE affine matrix (output pixel -> input pixel, centre-relative
 *                coordinates) exactly as torchvision's _get_inverse_affine_matrix returns it;
 *   flips[b]:    bit 0 horizontal flip, bit 1 vertical flip (NULL = none).
 * The sampling grid is evaluated with the same fp32 operation order as torch's affine_grid +
 * grid_sample(nearest, zeros, align_corners=False), so the result is a bit-exact copy / zero.
 * The batch gather (DataGenerator.get_next_train_batch, Datagenerators.py:43-65) and ToTensor's uint8 -> /255
 * (Datagenerators.py:133-134) are folded into the same pass.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  const void* in;           /* [Nsrc,C,H,W] fp32, or uint8 when in_u8 (ToTensor: value / 255) */
  float* out;               /* [B,C,H,W] fp32 */
  const float* theta;       /* [B,6] */
  const int32_t* flips;     /* [B] or NULL */
  const int32_t* src_index; /* [B] row of `in` each output sample is drawn from (the batch gather), or NULL = b */
  int32_t B, C, H, W;
  int32_t in_u8;
} pb_affine_nearest_args;
int pb_affine_nearest(const pb_affine_nearest_args* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * 2x2/2 max-pool followed by LeakyReLU (pytorch/CNNs.py:77,82) and its backward, which also
 * applies the LeakyReLU-backward of the producing conv layer (mask) so the result feeds wgrad.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  const void* x;         /* [N,H,W,C] */
  void* y;               /* [N,H/2,W/2,C] */
  int32_t N, H, W, C;
  float slope;
  int32_t act_dtype;
  void* y2;              /* optional bf16 twin of y when act_dtype is PB_F16 (see pb_conv_args.out2) */
} pb_pool_fwd_args;
int pb_maxpool_lrelu_fwd(const pb_pool_fwd_args* a, void* stream);

typedef struct {
  const void* x;         /* [N,H,W,C] forward input of the pool */
  const void* gy;        /* [N,H/2,W/2,C] grad wrt pool+lrelu output */
  const uint32_t* mask;  /* sign bits of the conv that produced x, or NULL */
  void* gx;              /* [N,H,W,C] grad wrt x */
  void* gx_masked;       /* [N,H,W,C] gx * lrelu'(mask), or NULL */
  int32_t N, H, W, C;
  float slope;
  int32_t act_dtype;     /* of gy / gx / gx_masked */
  int32_t x_dtype;       /* pb_dtype of x when it differs from act_dtype (PB_F16 forward activations); 0 = same */
} pb_pool_bwd_args;
int pb_maxpool_lrelu_bwd(const pb_pool_bwd_args* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * Heatmap loss: torch.nn.MSELoss()(o,t)/accumulation_steps and its gradient
 * (pytorch/train_pytorch.py:110,134-137).  o,t are NCHW fp32 [B,C,H,W].
 *   loss_sum[0] += sum (o-t)^2              (caller zeroes it; mean = loss_sum/numel)
 *   grad_nchw   = (o-t)*grad_scale          (optional; plain dL/do for the autograd path)
 *   grad_nhwc   = (o-t)*grad_scale*(o>0?1:slope), channel-padded NHWC act_dtype (optional; the
 *                 operand the last layer's dgrad/wgrad consume -- fuses LeakyReLU' of CNNs.py:155)
 * If target==NULL the target is synthesised on the fly from points (fused Gaussian, sigma).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  const float* out;
  const float* target;      /* or NULL with points != NULL */
  const float* points;      /* [B,C,2] (x,y) */
  float sigma;
  float* loss_sum;          /* 1 float (fp32 accumulation via atomics over block partials) */
  double* loss_sum_f64;     /* optional exact-order accumulation target, may be NULL */
  float* grad_nchw;
  void* grad_nhwc;
  int32_t B, C, H, W, Cpad;
  float grad_scale;         /* 2*scale/(numel*accumulation_steps) */
  float slope;
  int32_t act_dtype;
} pb_mse_args;
int pb_mse_loss_fwd_bwd(const pb_mse_args* a, void* stream);

/* NCHW fp32 upstream gradient -> NHWC act_dtype (channel padded), times LeakyReLU'(out) */
typedef struct {
  const float* grad_nchw;
  const float* out_nchw;    /* network output (sign source) or NULL for no activation backward */
  void* grad_nhwc;
  int32_t B, C, H, W, Cpad;
  float slope;
  int32_t act_dtype;
} pb_grad_ingest_args;
int pb_grad_ingest(const pb_grad_ingest_args* a, void* stream);

/* Gaussian target heatmaps: tensorflow/simple_data_generator.py:119-136.  out NCHW fp32. */
typedef struct {
  const float* points;      /* [B*C][2] (x,y) */
  float* out;               /* [B*C][H][W] */
  int32_t BC, H, W;
  float sigma;
} pb_gaussian_args;
int pb_gaussian_heatmaps(const pb_gaussian_args* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * Peaks.  argmax: Augmentor.tf_find_peaks, pytorch/Augmentor.py:105-148 -- lowest flat index on
 * ties, NaN is the maximum (first NaN).  soft: find_peaks_soft_argmax, pytorch/utils.py:47-83.
 * The heatmap element (n, y, x, c) is read at  n*stride_n + y*stride_y + x*stride_x + c*stride_c
 * so NHWC (the reference's argument layout) and NCHW (the network output) are both zero-copy.
 * peaks: [N][C][2] fp32 (x=col, y=row);  values (optional): [N][C] fp32 maximum.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  const void* heatmaps;
  float* peaks;
  float* values;
  int32_t N, C, H, W;
  int64_t stride_n, stride_c, stride_y, stride_x;
  int32_t dtype;            /* pb_dtype of heatmaps */
} pb_peaks_args;
int pb_peaks_argmax(const pb_peaks_args* a, void* stream);
int pb_peaks_softargmax(const pb_peaks_args* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused Adam over a flat fp32 parameter buffer (torch.optim.Adam defaults,
 * pytorch/train_pytorch.py:111,140): one launch for the whole model.  grad is multiplied by
 * grad_scale first (1/world, 1/loss-scale).  found_inf (optional, device int) != 0 skips the
 * update like GradScaler.step (train_pytorch.py:140).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  int64_t n;
  float lr, beta1, beta2, eps, weight_decay;
  float grad_scale;
  int32_t step;             /* 1-based */
  const int32_t* found_inf;
  /* CUDA-graph-safe form (a captured launch must not freeze per-step scalars): when step_dev is set, `step` is ignored
   * and the kernel reads t = *step_dev + 1 (bias corrections computed on the device in double); when lr_dev is set it
   * overrides `lr`; inc_step != 0 increments *step_dev behind this launch (set on the step's last Adam launch). */
  int32_t* step_dev;
  const float* lr_dev;
  int32_t inc_step;
} pb_adam_args;
int pb_adam_step(const pb_adam_args* a, void* stream);
/* adds n to the library's launch counter (pb_launch_count): a CUDA-graph replay launches the kernels counted at capture */
void pb_note_launches(int32_t n);

/* ------------------------------------------------------------------------------------------
 * ViT pieces (pytorch/pytorch_vit_encoder.py, pytorch/VITs.py)
 * ---------------------------------------------------------------------------------------- */
/* patchify :135-138  NCHW fp32 [B,C,H,W] -> [B*(H/p)*(W/p)][C*p*p] act_dtype, feature order (c,ph,pw) */
typedef struct {
  const float* img;
  void* patches;
  int32_t B, C, H, W, P;
  int32_t act_dtype;
} pb_patchify_args;
int pb_patchify(const pb_patchify_args* a, void* stream);

/* nn.LayerNorm(dim), eps 1e-5 (:19,47,90,123); y = ln(x)*gamma+beta (+ add, e.g. pos_embedding :144) */
typedef struct {
  const void* x;            /* [rows][dim] */
  const float* gamma;
  const float* beta;
  const float* add;         /* [add_rows][dim] broadcast over rows modulo add_rows, or NULL */
  void* y;
  float* mean;              /* [rows] saved for backward, or NULL */
  float* rstd;
  int32_t rows, dim, add_rows;
  float eps;
  int32_t act_dtype;
} pb_layernorm_fwd_args;
int pb_layernorm_fwd(const pb_layernorm_fwd_args* a, void* stream);

typedef struct {
  const void* x;
  const void* gy;
  const float* gamma;
  const float* mean;
  const float* rstd;
  const void* gx_add;       /* optional: gx = ln_bwd + gx_add (residual branch gradient) */
  void* gx;
  float* dgamma_partial;    /* [nblk][dim] */
  float* dbeta_partial;     /* [nblk][dim] */
  int32_t rows, dim, nblk;
  int32_t act_dtype;
  float* gx_colsum_partial; /* optional [nblk][dim]: per-block column sums of the stored gx -- the bias gradient of the
                             * nn.Linear whose output gradient gx is (to_out / the MLP's second Linear,
                             * pytorch_vit_encoder.py:20-23,55); bf16, dim == 256 only, else PB_ERR_UNSUPPORTED */
} pb_layernorm_bwd_args;
int pb_layernorm_bwd(const pb_layernorm_bwd_args* a, void* stream);

/* column sums of partial blocks: out[j] = beta*out[j] + alpha*sum_b partial[b][j]
 * (partial is fp32, or in_dtype when reducing an activation tensor, e.g. the pos_embedding grad) */
typedef struct {
  const void* partial;
  float* out;
  int32_t nblk, dim;
  float alpha, beta;
  int32_t in_dtype;
  const void* partial2;     /* optional second problem of the same shape in the same launch (LayerNorm's dbeta */
  float* out2;              /* next to its dgamma); NULL = none */
} pb_colsum_args;
int pb_colsum(const pb_colsum_args* a, void* stream);

/* softmax(q k^T * scale) v per (batch, head): Attention.forward :59-78.
 * qkv: [B*S][3*H*D] (q | k | v, each H blocks of D) ; out: [B*S][H*D].
 * probs (optional, fp32 [B][H][S][S]) is saved for the backward. */
typedef struct {
  const void* qkv;
  void* out;
  float* probs;
  int32_t B, S, H, D;
  float scale;
  int32_t act_dtype;
} pb_attention_fwd_args;
int pb_attention_fwd(const pb_attention_fwd_args* a, void* stream);

typedef struct {
  const void* qkv;
  const float* probs;
  const void* gout;         /* [B*S][H*D] */
  void* gqkv;               /* [B*S][3*H*D] */
  float* dprobs_ws;         /* fp32 [B][H][S][S] workspace */
  int32_t B, S, H, D;
  float scale;
  int32_t act_dtype;
} pb_attention_bwd_args;
int pb_attention_bwd(const pb_attention_bwd_args* a, void* stream);

/* y[b][j][i] = x[b][i][j] for `batch` row-major [rows x cols] matrices.  CNN_Decoder.forward
 * (pytorch/VITs.py:39) re-reads the (144 x 256) token matrix of a sample as a (256, 12, 12) NCHW
 * tensor; in the NHWC pipeline that reinterpretation is exactly this transpose. */
typedef struct {
  const void* x;
  void* y;
  int32_t batch, rows, cols;
  int32_t act_dtype;
} pb_transpose_args;
int pb_batched_transpose(const pb_transpose_args* a, void* stream);

/* GELU backward fused multiply: gx = gy * gelu'(pre).  With colsum_partial != NULL (bf16, [rows][dim] matrices,
 * dim a multiple of 8 and <= 2048) every block also writes the column sums of the gx rows it produced to
 * colsum_partial[block][dim] (nblk blocks): the bias gradient of the nn.Linear in front of the GELU
 * (pytorch_vit_encoder.py:20-21) without another pass over gx; fold them with pb_colsum. */
typedef struct {
  const void* pre;
  const void* gy;
  void* gx;
  int64_t n;
  int32_t act_dtype;
  int32_t dim;              /* row length (only read when colsum_partial != NULL) */
  float* colsum_partial;    /* [nblk][dim] fp32 or NULL */
  int32_t nblk;
} pb_gelu_bwd_args;
int pb_gelu_bwd(const pb_gelu_bwd_args* a, void* stream);

/* CNN_Decoder.normalize_between_0_and_1, pytorch/VITs.py:55-58: global min/max over the whole
 * NCHW fp32 tensor (NaN anywhere -> NaN everywhere; max == min -> inf/NaN, as in the reference). */
typedef struct {
  const float* x;
  float* y;
  void* minmax;             /* 16 bytes of device scratch: {u32 min key, u32 max key, float min, float max} */
  int64_t n;
} pb_minmax_norm_fwd_args;
int pb_minmax_normalize_fwd(const pb_minmax_norm_fwd_args* a, void* stream);

typedef struct {
  const float* x;
  const float* gy;
  const void* minmax;       /* as written by the forward */
  float* gx;
  void* scratch;            /* 32 bytes: {double sum gy, double sum gy*x, u64 argmin, u64 argmax} */
  int64_t n;
} pb_minmax_norm_bwd_args;
int pb_minmax_normalize_bwd(const pb_minmax_norm_bwd_args* a, void* stream);

/* Fused tail of VIT_encoder_CNN_decoder's training step: normalize_between_0_and_1 (pytorch/VITs.py:45,55-58) ->
 * MSELoss (pytorch/train_pytorch.py:110,134) -> both backwards -> LeakyReLU' of deconv4 (VITs.py:44), i.e. the chain
 * pb_minmax_normalize_fwd, pb_mse_loss_fwd_bwd, pb_minmax_normalize_bwd, pb_grad_ingest in three passes over x
 * (min/max; loss + the two sums the min/max gradient needs; gradient) instead of eleven tensor-sized transfers.
 *   y = (x - min x) / (max x - min x);  loss_sum[0] += sum (y - t)^2;  g = (y - t) * grad_scale
 *   dx = g / range, plus at the FIRST flat index holding the min:  (sum g*x - max * sum g) / range^2
 *                   and  at the FIRST flat index holding the max: -(sum g*x - min * sum g) / range^2
 *   grad_nhwc = dx * (x > 0 ? 1 : slope), channel-padded NHWC bf16 (what deconv4's dgrad / wgrad consume).
 * Same arithmetic, element for element, as the four separate entry points. */
typedef struct {
  const float* x;           /* [B,C,H,W] fp32: deconv4's output (after its LeakyReLU), before the normalisation */
  const float* target;      /* [B,C,H,W] or NULL with points != NULL */
  const float* points;      /* [B,C,2] (x,y): sigma-Gaussian targets rendered on the fly */
  float sigma;
  float* loss_sum;          /* 1 float, caller zeroes it */
  void* grad_nhwc;          /* [B,H,W,Cpad] bf16 */
  void* scratch;            /* 48 bytes of device scratch (min/max keys and values, the two sums, argmin/argmax) */
  int32_t B, C, H, W, Cpad;
  float grad_scale;         /* 2*scale/(numel*accumulation_steps) */
  float slope;
} pb_minmax_mse_args;
int pb_minmax_mse_fwd_bwd(const pb_minmax_mse_args* a, void* stream);

/* Column-block move between row-major matrices -- the "glue" of the multi-camera models
 * (torch.cat / torch.split along features, the broadcast of the joint encoding to the four views, the sum of the four
 * views' gradients; pytorch/CNNs.py:226-236, pytorch/VITs.py:295-305) as ONE strided pass instead of ATen copies:
 *   dst[r][dst_col0 + j] = (accumulate ? dst[r][dst_col0 + j] : 0)
 *                        + sum_{f < nfold} src[f * fold_stride + (r % src_rows_mod) * src_row_stride + src_col0 + j]
 * for r < rows, j < ncols (sums in fp32).  src_rows_mod = 0 means "no wrap".  Strides / offsets in ELEMENTS. */
typedef struct {
  const void* src;
  void* dst;
  int64_t rows, ncols;
  int64_t src_row_stride, dst_row_stride;
  int64_t src_col0, dst_col0;
  int64_t src_rows_mod;
  int64_t fold_stride;
  int32_t nfold;            /* >= 1 */
  int32_t accumulate;
  int32_t dtype;            /* pb_dtype of src and dst */
} pb_colblock_args;
int pb_colblock(const pb_colblock_args* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * FourCamerasDisentanglement pieces (pytorch/CNNs.py:240-352)
 * ---------------------------------------------------------------------------------------- */
/* FTL / InvFTL (CNNs.py:322-345): the reference reinterprets the NCHW tensor of a sample as consecutive groups of
 * `kin` values (no permute: groups run along the flat memory order), multiplies every group by the sample's
 * [kout x kin] matrix and reinterprets the result as NCHW again:
 *   out[b*out_batch_stride + kout*m + i] (+)= sum_j mats[b][i][j] * in[(b % in_batch_mod)*in_batch_stride + kin*m + j]
 * for m < groups.  FTL: kin 4, kout 3 (camera matrix); InvFTL: kin 3, kout 4; their backward passes are the same
 * kernel with the transposed matrices.  in / out are flat NCHW buffers (act dtype); in_batch_mod = 0 means "b". */
typedef struct {
  const void* in;
  void* out;
  const float* mats;        /* [B][kout][kin] */
  int32_t B;
  int64_t groups;
  int32_t kin, kout;
  int64_t in_batch_stride, out_batch_stride;
  int32_t in_batch_mod;
  int32_t accumulate;
  int32_t act_dtype;
} pb_ftl_args;
int pb_ftl(const pb_ftl_args* a, void* stream);

/* nn.BatchNorm2d (+ ReLU) on NHWC rows, CNNs.py:267-269,302-309.  The rows are `groups` consecutive blocks of
 * rows_per_group rows that the reference normalises in SEPARATE calls of the same module (batch_norm3 on the four
 * re-projected views): statistics are per (group, channel); running statistics are updated group after group.
 *   training: mean / biased variance of the batch; running_mean, running_var (unbiased), momentum as torch
 *   eval    : running statistics
 *   y = relu?((x - mean) * rstd * gamma + beta); channels [C, Cs) of the stored row (padding) are written as 0. */
typedef struct {
  const void* x;            /* [groups*rows_per_group][Cs] act dtype */
  void* y;
  const float* gamma;
  const float* beta;
  float* running_mean;      /* [C] (updated in training) */
  float* running_var;
  float* save_mean;         /* [groups][C] out (training) */
  float* save_rstd;
  float* partial;           /* workspace: [groups][nblk][2][C] fp32 */
  int32_t groups, rows_per_group, C, Cs, nblk;
  float eps, momentum;
  int32_t training, relu;
  int32_t act_dtype;
} pb_batchnorm_fwd_args;
int pb_batchnorm_fwd(const pb_batchnorm_fwd_args* a, void* stream);

/* backward of the above in training mode: gy is masked by (y > 0) when relu;
 *   dbeta[c] = beta_acc*dbeta[c] + sum gy;  dgamma[c] = beta_acc*dgamma[c] + sum gy * xhat   (summed over the groups)
 *   gx = gamma * rstd * (gy - mean_rows(gy) - xhat * mean_rows(gy * xhat))                  (per group) */
typedef struct {
  const void* x;
  const void* y;            /* forward output (ReLU mask source) or NULL when relu == 0 */
  const void* gy;
  void* gx;
  const float* gamma;
  const float* save_mean;
  const float* save_rstd;
  float* dgamma;
  float* dbeta;
  float* partial;           /* workspace: [groups][nblk][2][C] fp32 */
  int32_t groups, rows_per_group, C, Cs, nblk;
  float beta_acc;           /* 0 overwrite, 1 accumulate into dgamma / dbeta */
  int32_t relu;
  int32_t act_dtype;
} pb_batchnorm_bwd_args;
int pb_batchnorm_bwd(const pb_batchnorm_bwd_args* a, void* stream);

/* generic elementwise helper: out = (a + b) * (mask ? (bit ? 1 : slope) : 1)
 * (residual-gradient add and LeakyReLU backward at a module boundary; mask is the producing
 * layer's sign-bit tensor [n/C][ceil(C/32)]) */
typedef struct {
  const void* a;
  const void* b;            /* or NULL */
  void* out;
  int64_t n;
  int32_t act_dtype;
  const uint32_t* mask;     /* or NULL */
  int32_t C;                /* channels (innermost extent) when mask != NULL */
  float slope;
} pb_add_args;
int pb_add(const pb_add_args* a, void* stream);

/* self-test of the tcgen05 building blocks: D[M,N] = A[M,K] * B[N,K]^T (kmajor) or with
 * MN-major operands; used by the GPU tests to pin descriptor encodings. */
typedef struct {
  const void* a;            /* bf16: a_mn_major ? [K][M] : [M][K] */
  const void* b;            /* bf16: b_mn_major ? [K][N] : [N][K] */
  float* d;                 /* fp32 [M][N] */
  int32_t M, N, K;
  int32_t a_mn_major, b_mn_major;
  int32_t a_f16, b_f16;     /* operand holds IEEE half instead of bf16 (instruction-descriptor format fields);
                               sm_100a raises an illegal instruction unless both are equal */
} pb_gemm_selftest_args;
int pb_gemm_selftest(const pb_gemm_selftest_args* a, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* POSEB200_H */
