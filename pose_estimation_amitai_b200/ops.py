"""Thin torch-tensor wrappers over the C ABI (include/poseb200.h).

Every function takes CUDA tensors, passes raw device pointers + sizes to libposeb200.so on
torch's current stream and returns torch tensors.  Nothing here computes on the CPU or through
ATen: a CPU tensor raises, a missing library raises (``_lib.LibraryMissing``).
"""
from __future__ import annotations

import os
import ctypes
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import PB_ACT_GELU, PB_ACT_LRELU, PB_ACT_MASKMUL, PB_ACT_NONE, PB_BF16, PB_F16, PB_F32, STRUCTS

LEAKY_SLOPE = 0.1
_N_SM: Optional[int] = None      # SM count of the current device (device_info), cached


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("pose_estimation_amitai_b200: CPU tensor passed to a CUDA op (there is no CPU fallback)")
    return t.data_ptr()


def pb_dtype(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return PB_F32
    if dt == torch.bfloat16:
        return PB_BF16
    if dt == torch.float16:
        return PB_F16
    raise TypeError(f"unsupported activation dtype {dt}")


# ------------------------------------------------------------------------------------------
# contraction descriptors
# ------------------------------------------------------------------------------------------
@dataclass
class Contraction:
    """One conv-like layer of the reference and the three gather-convolutions derived from it.

    kind: 'conv'   nn.Conv2d(k, dilation d, padding d*(k-1)/2)        pytorch/CNNs.py:45-49
          'convT1' nn.ConvTranspose2d(k3, s1, p1)                       pytorch/CNNs.py:113-122
          'convT2' nn.ConvTranspose2d(k3, s2, p1, op1)                  pytorch/CNNs.py:108-110,125-128
          'linear' nn.Linear                                            pytorch/pytorch_vit_encoder.py:20-23
    """
    kind: str
    cin: int
    cout: int
    dilation: int = 1
    ksize: int = 3
    # filled by __post_init__
    kpos: List[int] = field(default_factory=list)
    fwd_dy: List[int] = field(default_factory=list)
    fwd_dx: List[int] = field(default_factory=list)

    def __post_init__(self):
        if self.kind == "linear":
            self.ksize = 1
        k = self.ksize
        self.kpos, self.fwd_dy, self.fwd_dx = [], [], []
        for r in range(k):
            for s in range(k):
                self.kpos.append(r * k + s)
                if self.kind == "conv":
                    c = (k - 1) // 2
                    self.fwd_dy.append(self.dilation * (r - c))
                    self.fwd_dx.append(self.dilation * (s - c))
                elif self.kind in ("convT1", "convT2"):
                    # out[o] += in[i] w[k] with o = i*stride - 1 + k  ->  i*stride = o + 1 - k
                    self.fwd_dy.append(1 - r)
                    self.fwd_dx.append(1 - s)
                else:
                    self.fwd_dy.append(0)
                    self.fwd_dx.append(0)

    @property
    def ntaps(self) -> int:
        return len(self.kpos)

    # strides of element (ci, co, kpos) inside the torch parameter tensor
    @property
    def stride_ci(self) -> int:
        kk = self.ksize * self.ksize
        return {"conv": kk, "convT1": self.cout * kk, "convT2": self.cout * kk, "linear": 1}[self.kind]

    @property
    def stride_co(self) -> int:
        kk = self.ksize * self.ksize
        return {"conv": self.cin * kk, "convT1": kk, "convT2": kk, "linear": self.cin}[self.kind]

    def out_hw(self, ih: int, iw: int) -> Tuple[int, int]:
        return (2 * ih, 2 * iw) if self.kind == "convT2" else (ih, iw)

    def fwd_taps(self):
        return make_taps(self.fwd_dy, self.fwd_dx, 1, 2 if self.kind == "convT2" else 1)

    def dgrad_taps(self):
        dy = [-v for v in self.fwd_dy]
        dx = [-v for v in self.fwd_dx]
        return make_taps(dy, dx, 2 if self.kind == "convT2" else 1, 1)


def make_taps(dy: Sequence[int], dx: Sequence[int], out_mul: int = 1, in_div: int = 1):
    t = STRUCTS["pb_taps"]()
    t.ntaps, t.out_mul, t.in_div = len(dy), out_mul, in_div
    for i, (a, b) in enumerate(zip(dy, dx)):
        t.dy[i], t.dx[i] = a, b
    return t


def pack_weights_args(weight: torch.Tensor, c: Contraction, role: str, dtype: torch.dtype, ipad: int = 0,
                      jpad: int = 0, dst: Optional[torch.Tensor] = None):
    """(pb_pack_weights_args, dst) for one parameter tensor; see pack_weights."""
    a = STRUCTS["pb_pack_weights_args"]()
    if role == "io":
        i_n, j_n, si, sj = c.cin, c.cout, c.stride_ci, c.stride_co
    else:
        i_n, j_n, si, sj = c.cout, c.cin, c.stride_co, c.stride_ci
    ip, jp = max(ipad, i_n), max(jpad, j_n)
    if dst is None:
        dst = torch.empty((c.ntaps, ip, jp), device=weight.device, dtype=dtype)
    assert dst.shape == (c.ntaps, ip, jp) and dst.dtype == dtype
    w = weight.detach()
    assert w.is_contiguous() and w.dtype == torch.float32
    a.src, a.dst = _ptr(w), _ptr(dst)
    a.ntaps, a.I, a.Ipad, a.J, a.Jpad = c.ntaps, i_n, ip, j_n, jp
    a.stride_i, a.stride_j = si, sj
    for t, kp in enumerate(c.kpos):
        a.kpos[t] = kp
    a.dst_dtype = pb_dtype(dtype)
    return a, dst


def pack_weights(weight: torch.Tensor, c: Contraction, role: str, dtype: torch.dtype, ipad: int = 0,
                 jpad: int = 0, dst: Optional[torch.Tensor] = None) -> torch.Tensor:
    """role 'io': dst[t][ci][co]  (simt forward operand / tcgen05 dgrad operand)
       role 'oi': dst[t][co][ci]  (simt dgrad operand / tcgen05 forward operand, K contiguous)
       dst: re-pack into an existing operand tensor (same shape / dtype) instead of allocating."""
    a, dst = pack_weights_args(weight, c, role, dtype, ipad, jpad, dst=dst)
    _lib.call("pb_pack_weights", a, _stream())
    return dst


def pack_table(items: list, device) -> Tuple[torch.Tensor, int]:
    """device copy of an array of pb_pack_weights_args (for pack_weights_multi) and its largest element count."""
    arr = (STRUCTS["pb_pack_weights_args"] * len(items))(*items)
    raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).clone()
    return raw.to(device), max(int(a.ntaps) * int(a.Ipad) * int(a.Jpad) for a in items)


def pack_weights_multi(table: torch.Tensor, count: int, max_elems: int) -> None:
    """refresh every packed operand described by `table` in one launch."""
    m = STRUCTS["pb_pack_weights_multi_args"]()
    m.items, m.count, m.max_elems = _ptr(table), count, max_elems
    _lib.call("pb_pack_weights_multi", m, _stream())


def conv(impl: str, x: torch.Tensor, w: torch.Tensor, taps, n: int, ih: int, iw: int, cin: int, oh: int, ow: int,
         cout: int, *, bias: Optional[torch.Tensor] = None, act: int = PB_ACT_NONE, slope: float = LEAKY_SLOPE,
         add0: Optional[torch.Tensor] = None, add1: Optional[torch.Tensor] = None,
         pre_out: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
         mask_out: Optional[torch.Tensor] = None, mask_in: Optional[torch.Tensor] = None,
         act_dtype: torch.dtype = torch.float32, in_nchw: bool = False, out_nchw: bool = False,
         prof_cin: Optional[int] = None, out2: Optional[torch.Tensor] = None,
         pool_out: Optional[torch.Tensor] = None, pool_only: bool = False) -> torch.Tensor:
    """out = epilogue(gather_conv(x, w)); see pb_conv_args in include/poseb200.h.
    out2: bf16 twin of an fp16 NHWC `out`, written by the same epilogue.
    pool_out [n, oh/2, ow/2, cout]: also emit lrelu(maxpool2x2(out)) from the same epilogue (tensor-core path);
    pool_only: `out` itself is not written (returns pool_out)."""
    if pool_out is not None:
        assert impl == "tc" and not out_nchw and pool_out.shape == (n, oh // 2, ow // 2, cout) and pool_out.dtype == act_dtype
    if pool_only:
        assert pool_out is not None and out is None
        out = pool_out          # never written: only keeps the argument checks uniform
    if out is None:
        if out_nchw:
            out = torch.empty((n, cout, oh, ow), device=x.device, dtype=torch.float32)
        else:
            out = torch.empty((n, oh, ow, cout), device=x.device, dtype=act_dtype)
    a = STRUCTS["pb_conv_args"]()
    setattr(a, "in", _ptr(x))
    a.w, a.bias, a.add0, a.add1 = _ptr(w), _ptr(bias), _ptr(add0), _ptr(add1)
    a.pre_out, a.out, a.mask_out, a.mask_in = _ptr(pre_out), _ptr(out), _ptr(mask_out), _ptr(mask_in)
    if out2 is not None:
        assert act_dtype == torch.float16 and not out_nchw and out2.shape == out.shape and out2.dtype == torch.bfloat16
        a.out2 = _ptr(out2)
    a.N, a.IH, a.IW, a.Cin, a.OH, a.OW, a.Cout = n, ih, iw, cin, oh, ow, cout
    a.act, a.slope = act, slope
    a.act_dtype = pb_dtype(act_dtype)
    a.out_nchw_f32, a.in_nchw_f32 = int(out_nchw), int(in_nchw)
    a.taps = taps
    a.pool_out, a.pool_only = _ptr(pool_out), int(pool_only)
    fn = "pb_conv_tc" if impl == "tc" else "pb_conv_simt"
    if _PROFILE is None:
        _lib.call(fn, a, _stream())
    else:
        macs = n * oh * ow * taps.ntaps * (prof_cin or cin) * cout // (taps.in_div ** 2)  # no padding counted
        _timed(fn, 2.0 * macs, lambda: _lib.call(fn, a, _stream()))
    return out


def _head_args(x: torch.Tensor, w: torch.Tensor, taps, n: int, ih: int, iw: int, cin: int, cout: int,
               bias: Optional[torch.Tensor], slope: float):
    h = STRUCTS["pb_head_fused_args"]()
    a = h.conv
    setattr(a, "in", _ptr(x))
    a.w, a.bias = _ptr(w), _ptr(bias)
    a.N, a.IH, a.IW, a.Cin, a.OH, a.OW, a.Cout = n, ih, iw, cin, 2 * ih, 2 * iw, cout
    a.act, a.slope, a.act_dtype, a.out_nchw_f32 = PB_ACT_LRELU, slope, pb_dtype(x.dtype), 1
    a.taps = taps
    return h


def head_folded_supported(cin: int, cout: int, dtype: torch.dtype) -> bool:
    """mirror of csrc/tc_head.cu's shape test (head_tc): the folded-parity head kernel takes this layer."""
    nt = (cout + 15) // 16 * 16
    w_bytes = 9 * (cin // 64) * nt * 128
    return (cin % 64 == 0 and nt <= 48 and dtype in (torch.bfloat16, torch.float16)
            and w_bytes + 2 * 20480 <= 220 * 1024 - 1024 and os.environ.get("POSEB200_HEAD_V2", "1") != "0")


def head_dbias_buffer(cout: int, device) -> torch.Tensor:
    """zeroed per-CTA partial buffer for pb_head_fused_args.dbias; fold it with vit_ops.colsum(buf, db, rows, cout)."""
    global _N_SM
    if _N_SM is None:
        _N_SM = device_info()[2]
    return torch.zeros((_N_SM, cout), device=device, dtype=torch.float32)


def head_argmax_fused(x: torch.Tensor, w: torch.Tensor, taps, n: int, ih: int, iw: int, cin: int, cout: int, *,
                      bias: Optional[torch.Tensor] = None, slope: float = LEAKY_SLOPE, want_values: bool = False):
    """last layer (stride-2 transposed conv + LeakyReLU) + per-map arg-max in ONE kernel: (n, cout, 2) [x, y] peaks
    (and the maxima); the heatmaps are never written (pb_convT_argmax_fused)."""
    h = _head_args(x, w, taps, n, ih, iw, cin, cout, bias, slope)
    peaks = torch.empty((n, cout, 2), device=x.device, dtype=torch.float32)
    values = torch.empty((n, cout), device=x.device, dtype=torch.float32) if want_values else None
    h.peaks, h.values = _ptr(peaks), _ptr(values)
    if _PROFILE is None:
        _lib.call("pb_convT_argmax_fused", h, _stream())
    else:
        _timed("pb_conv_tc", 2.0 * n * ih * iw * taps.ntaps * cin * cout,
               lambda: _lib.call("pb_convT_argmax_fused", h, _stream()))
    return (peaks, values) if want_values else peaks


def head_mse_fused(x: torch.Tensor, w: torch.Tensor, taps, n: int, ih: int, iw: int, cin: int, cout: int, *,
                   bias: Optional[torch.Tensor] = None, slope: float = LEAKY_SLOPE,
                   target: Optional[torch.Tensor] = None, points: Optional[torch.Tensor] = None, sigma: float = 3.0,
                   accumulation_steps: int = 1, loss_scale: float = 1.0, grad_out: Optional[torch.Tensor] = None,
                   dbias_out: Optional[torch.Tensor] = None):
    """last layer + MSELoss + the gradient w.r.t. its pre-activation in ONE kernel (pb_convT_mse_fused):
    returns (loss_sum tensor[1], grad_nhwc bf16 [n, 2ih, 2iw, cpad]); mean loss = loss_sum / (n*cout*4*ih*iw) /
    accumulation_steps, as ops.mse_loss_fwd_bwd."""
    h = _head_args(x, w, taps, n, ih, iw, cin, cout, bias, slope)
    cpad = (cout + 15) // 16 * 16
    loss_sum = torch.zeros(1, device=x.device, dtype=torch.float32)
    grad = grad_out if grad_out is not None else torch.empty((n, 2 * ih, 2 * iw, cpad), device=x.device, dtype=torch.bfloat16)
    assert grad.shape == (n, 2 * ih, 2 * iw, cpad) and grad.dtype == torch.bfloat16 and grad.is_contiguous()
    if target is not None:
        assert target.shape == (n, cout, 2 * ih, 2 * iw) and target.dtype == torch.float32 and target.is_contiguous()
    else:
        assert points is not None and points.shape == (n, cout, 2) and points.dtype == torch.float32 and points.is_contiguous()
    h.target, h.points, h.sigma = _ptr(target), _ptr(points), sigma
    h.loss_sum, h.grad_nhwc, h.Cpad = _ptr(loss_sum), _ptr(grad), cpad
    h.grad_scale = 2.0 * loss_scale / (n * cout * 4 * ih * iw * accumulation_steps)
    if dbias_out is not None:
        # zeroed [rows >= SMs][cout] fp32: row b receives CTA b's share of the head's bias gradient (column sum = the
        # gradient; see head_dbias_buffer)
        assert dbias_out.dtype == torch.float32 and dbias_out.dim() == 2 and dbias_out.shape[1] == cout and dbias_out.is_contiguous()
        h.dbias, h.dbias_rows = _ptr(dbias_out), int(dbias_out.shape[0])
    if _PROFILE is None:
        _lib.call("pb_convT_mse_fused", h, _stream())
    else:
        _timed("pb_conv_tc", 2.0 * n * ih * iw * taps.ntaps * cin * cout,
               lambda: _lib.call("pb_convT_mse_fused", h, _stream()))
    return loss_sum, grad


# ------------------------------------------------------------------------------------------
# per-launch timing of the contraction kernels (bench.py roofline leg): CUDA events on the
# launching stream around each call, collected only while profiling is switched on
# ------------------------------------------------------------------------------------------
_PROFILE: Optional[list] = None


def profile_begin() -> None:
    global _PROFILE
    _PROFILE = []


def profile_end() -> List[Tuple[str, float, float]]:
    """returns [(entry point, algorithmic FLOPs, milliseconds)] and switches profiling off."""
    global _PROFILE
    rec, _PROFILE = _PROFILE or [], None
    torch.cuda.synchronize()
    return [(name, flops, e0.elapsed_time(e1)) for name, flops, e0, e1 in rec]


def _timed(name: str, flops: float, launch) -> None:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    launch()
    e1.record()
    _PROFILE.append((name, flops, e0, e1))


def im2col_first(x_nchw: torch.Tensor, ksize: int, dilation: int, kpad: int, dtype: torch.dtype) -> torch.Tensor:
    """(N,C,H,W) fp32 -> (N,H,W,kpad): k = ci*ksize^2 + r*ksize + s (see pb_im2col_args)."""
    n, c, h, w = x_nchw.shape
    out = torch.empty((n, h, w, kpad), device=x_nchw.device, dtype=dtype)
    a = STRUCTS["pb_im2col_args"]()
    setattr(a, "in", _ptr(x_nchw))
    a.out = _ptr(out)
    a.N, a.C, a.H, a.W, a.ksize, a.dilation, a.Kpad = n, c, h, w, ksize, dilation, kpad
    a.act_dtype = pb_dtype(dtype)
    _lib.call("pb_im2col_first", a, _stream())
    return out


def conv_first_supported(cin: int, ksize: int, cout: int) -> bool:
    """shapes csrc/tc_conv1.cu tiles (pb_conv_first_tc)."""
    import os
    return (1 <= cin <= 4 and ksize == 3 and cout in (32, 64, 128)
            and os.environ.get("POSEB200_NO_CONV1_DIRECT", "0") != "1")


def conv_first(x_nchw: torch.Tensor, w_packed: torch.Tensor, bias: Optional[torch.Tensor], cout: int, dilation: int,
               act_dtype: torch.dtype, *, slope: float = LEAKY_SLOPE, mask_out: Optional[torch.Tensor] = None,
               ksize: int = 3, out: Optional[torch.Tensor] = None, out2: Optional[torch.Tensor] = None) -> torch.Tensor:
    """LeakyReLU(conv1(x) + bias) straight from the NCHW fp32 crops, NHWC `act_dtype` out (pb_conv_first_tc).
    out2: bf16 twin of an fp16 `out`, written by the same epilogue."""
    n, c, h, w = x_nchw.shape
    assert x_nchw.dtype == torch.float32 and x_nchw.is_contiguous()
    assert w_packed.dtype == act_dtype and w_packed.numel() == cout * 64
    if out is None:
        out = torch.empty((n, h, w, cout), device=x_nchw.device, dtype=act_dtype)
    assert out.shape == (n, h, w, cout) and out.dtype == act_dtype and out.is_contiguous()
    a = STRUCTS["pb_conv_first_args"]()
    setattr(a, "in", _ptr(x_nchw))
    a.w, a.bias, a.out, a.mask_out = _ptr(w_packed), _ptr(bias), _ptr(out), _ptr(mask_out)
    if out2 is not None:
        assert act_dtype == torch.float16 and out2.shape == out.shape and out2.dtype == torch.bfloat16 and out2.is_contiguous()
        a.out2 = _ptr(out2)
    a.N, a.C, a.H, a.W, a.ksize, a.dilation, a.Cout = n, c, h, w, ksize, dilation, cout
    a.slope, a.act_dtype = slope, pb_dtype(act_dtype)
    if _PROFILE is None:
        _lib.call("pb_conv_first_tc", a, _stream())
    else:
        _timed("pb_conv_tc", 2.0 * n * h * w * ksize * ksize * c * cout, lambda: _lib.call("pb_conv_first_tc", a, _stream()))
    return out


def wgrad_first_supported(cin: int, ksize: int, cout: int, dtype: torch.dtype) -> bool:
    """shapes csrc/tc_wgrad1.cu takes (pb_wgrad_first_tc)."""
    import os
    return (1 <= cin <= 4 and ksize == 3 and cout == 64 and dtype == torch.bfloat16
            and os.environ.get("POSEB200_NO_WGRAD1_DIRECT", "0") != "1")


def wgrad_first(x_nchw: torch.Tensor, g: torch.Tensor, dw: torch.Tensor, dbias: Optional[torch.Tensor], dilation: int, *,
                beta: float = 0.0, alpha: float = 1.0, workspace: Optional[torch.Tensor] = None) -> None:
    """conv1's weight / bias gradient straight from the NCHW fp32 crops (no im2col tensor): dw [cout, cin, 3, 3] =
    beta*dw + alpha * d(loss)/d(weight); g [n, h, w, cout] bf16 is the gradient w.r.t. conv1's pre-activation."""
    global _N_SM
    n, cin, h, w = x_nchw.shape
    cout = int(g.shape[-1])
    assert x_nchw.dtype == torch.float32 and x_nchw.is_contiguous() and g.dtype == torch.bfloat16 and g.is_contiguous()
    if _N_SM is None:
        _N_SM = device_info()[2]
    tiles = n * ((h + 3) // 4) * ((w + 31) // 32)
    ks = max(1, min(_N_SM, tiles))
    ca = 64
    L = ca * cout + cout
    if workspace is None or workspace.numel() < ks * L:
        workspace = torch.empty(ks * L, device=g.device, dtype=torch.float32)
    a = STRUCTS["pb_wgrad_first_args"]()
    setattr(a, "in", _ptr(x_nchw))
    a.g, a.partial = _ptr(g), _ptr(workspace)
    a.N, a.C, a.H, a.W, a.ksize, a.dilation, a.Cg, a.Ca, a.ksplit = n, cin, h, w, 3, dilation, cout, ca, ks
    a.act_dtype = pb_dtype(g.dtype)
    if _PROFILE is None:
        _lib.call("pb_wgrad_first_tc", a, _stream())
    else:
        _timed("pb_wgrad_tc", 2.0 * n * h * w * 9 * cin * cout, lambda: _lib.call("pb_wgrad_first_tc", a, _stream()))
    c = Contraction("linear", cin * 9, cout)
    r = STRUCTS["pb_wgrad_reduce_args"]()
    r.partial, r.dw, r.dbias = _ptr(workspace), _ptr(dw), _ptr(dbias)
    r.ksplit, r.ntaps, r.Ca, r.Cg, r.Ca_valid = ks, 1, ca, cout, cin * 9
    r.stride_a, r.stride_g = c.stride_ci, c.stride_co
    r.kpos[0] = c.kpos[0]
    r.beta, r.alpha = beta, alpha
    _lib.call("pb_wgrad_reduce", r, _stream())


def wgrad_workspace_len(c: Contraction, ca_stored: int = 0) -> int:
    return c.ntaps * max(c.cin, ca_stored) * c.cout + c.cout


def wgrad_v2_eligible(c: Contraction, ph: int, pw: int) -> bool:
    """mirror of csrc/tc_wgrad2.cu's shape test: conv / stride-1 transposed conv / 1-tap contractions on real
    images (>= one 16 x 8 pixel tile)."""
    import os
    if os.environ.get("POSEB200_WGRAD_V1", "0") == "1":
        return False
    return c.kind in ("conv", "convT1", "linear") and ph >= 16 and pw >= 8


def wgrad_up_eligible(c: Contraction, ph: int, pw: int, ca_stored: int = 0) -> bool:
    """mirror of csrc/tc_wgrad_up.cu's shape test: all nine taps of a narrow stride-2 transposed conv in one CTA."""
    import os
    ca = max(c.cin, ca_stored)
    return (c.kind == "convT2" and c.ksize == 3 and ca % 128 == 0 and 9 * ((c.cout + 15) // 16 * 16) <= 512
            and ph >= 16 and pw >= 8 and os.environ.get("POSEB200_WGRAD_UP", "1") != "0")


def choose_ksplit(c: Contraction, pixels: int, n_sm: int = 148, impl: str = "simt", ph: int = 0, pw: int = 0,
                  ca_stored: int = 0) -> int:
    """number of pixel-range splits of a weight-gradient contraction (each split writes one fp32
    partial tile that pb_wgrad_reduce folds)."""
    if impl == "tc" and wgrad_up_eligible(c, ph, pw, ca_stored):
        units = max(c.cin, ca_stored) // 128                               # one CTA per 128-channel block and split
        return int(max(1, min(n_sm // units, max(1, pixels // 128))))
    if impl == "tc" and wgrad_v2_eligible(c, ph, pw):
        cib = (max(c.cin, ca_stored) + 63) // 64
        if c.cout % 128 == 0 and os.environ.get("POSEB200_WGRAD_NARROW", "0") != "1":
            # wide mode (tc_wgrad2.cu): 64 x 128 blocks; with more than 8 taps a second kind of CTA takes the last
            # tap pair over WG2_B_RATIO = 3 splits each, so one wave holds units * (ks + ks / 3) CTAs
            units = cib * (c.cout // 128)
            ks = n_sm // units if c.ntaps <= 8 else (3 * n_sm) // (4 * units)
            return int(max(1, min(ks, max(1, pixels // 128))))
        units = cib * ((c.cout + 63) // 64)                                # one CTA per 64x64 block of dW
        return int(max(1, min(n_sm // units, max(1, pixels // 128))))      # one wave of one-CTA-per-SM items
    if impl == "tc":
        units = max(1, (c.ntaps + 1) // 2 if c.cin == 64 else c.ntaps * (c.cin // 128)) * max(1, c.cout // (256 if c.cout % 256 == 0 else 128) if c.cout > 256 else 1)
        ks = max(1, (2 * n_sm) // units)          # ~2 waves of one-CTA-per-SM work items
        return int(max(1, min(ks, max(1, pixels // 512))))
    tiles = c.ntaps * ((c.cin + 63) // 64) * ((c.cout + 63) // 64)
    ks = max(1, (4 * n_sm + tiles - 1) // tiles)
    return int(max(1, min(ks, max(1, pixels // 256))))


def wgrad(impl: str, c: Contraction, a_in: torch.Tensor, g: torch.Tensor, n: int, ih: int, iw: int,
          dw: torch.Tensor, dbias: Optional[torch.Tensor], *, act_dtype: torch.dtype, a_nchw: bool = False,
          beta: float = 0.0, alpha: float = 1.0, workspace: Optional[torch.Tensor] = None) -> None:
    """dw (parameter-shaped, fp32) = beta*dw + alpha * d(loss)/d(weight); same for dbias.
    a_in: layer input [n, ih, iw, cin]; g: grad wrt the layer's pre-activation [n, oh, ow, cout]; both `act_dtype`
    (in the "fp16" precision a_in is the bf16 twin the forward epilogue stored next to the fp16 activation)."""
    oh, ow = c.out_hw(ih, iw)
    w = STRUCTS["pb_wgrad_args"]()
    w.a, w.g = _ptr(a_in), _ptr(g)
    w.N = n
    if c.kind == "convT2":
        # sum over input pixels i: a[i] (x) g[2i - 1 + k]
        w.PH, w.PW = ih, iw
        w.mul_a, w.mul_g = 1, 2
        for t in range(c.ntaps):
            w.dya[t] = w.dxa[t] = 0
            w.dyg[t], w.dxg[t] = -c.fwd_dy[t], -c.fwd_dx[t]
    else:
        # sum over output pixels o: a[o + off] (x) g[o]
        w.PH, w.PW = oh, ow
        w.mul_a, w.mul_g = 1, 1
        for t in range(c.ntaps):
            w.dya[t], w.dxa[t] = c.fwd_dy[t], c.fwd_dx[t]
            w.dyg[t] = w.dxg[t] = 0
    ca = c.cin if a_nchw else int(a_in.shape[-1])  # channels as stored (>= cin when zero padded)
    w.AH, w.AW, w.Ca, w.GH, w.GW, w.Cg = ih, iw, ca, oh, ow, c.cout
    w.ntaps = c.ntaps
    pixels = n * w.PH * w.PW
    ks = choose_ksplit(c, pixels, impl=impl, ph=int(w.PH), pw=int(w.PW), ca_stored=ca)
    L = wgrad_workspace_len(c, ca)
    if workspace is None or workspace.numel() < ks * L:
        workspace = torch.empty(ks * L, device=g.device, dtype=torch.float32)
    w.partial = _ptr(workspace)
    w.ksplit = ks
    w.act_dtype = pb_dtype(act_dtype)
    if not a_nchw and a_in.dtype != g.dtype:
        raise TypeError(f"wgrad: activations ({a_in.dtype}) and gradients ({g.dtype}) must share one format "
                        "(tcgen05.mma kind::f16 takes A and B in the same 16-bit type)")
    w.a_nchw_f32 = int(a_nchw)
    w.want_bias = int(dbias is not None)
    w.g_cstride = g.shape[-1]
    fn = "pb_wgrad_tc" if impl == "tc" else "pb_wgrad_simt"
    if _PROFILE is None:
        _lib.call(fn, w, _stream())
    else:
        _timed(fn, 2.0 * n * ih * iw * c.ntaps * c.cin * c.cout, lambda: _lib.call(fn, w, _stream()))
    r = STRUCTS["pb_wgrad_reduce_args"]()
    r.partial, r.dw, r.dbias = _ptr(workspace), _ptr(dw), _ptr(dbias)
    r.ksplit, r.ntaps, r.Ca, r.Cg, r.Ca_valid = ks, c.ntaps, ca, c.cout, c.cin
    r.stride_a, r.stride_g = c.stride_ci, c.stride_co
    for t, kp in enumerate(c.kpos):
        r.kpos[t] = kp
    r.beta, r.alpha = beta, alpha
    _lib.call("pb_wgrad_reduce", r, _stream())


def maxpool_lrelu_fwd(x: torch.Tensor, slope: float = LEAKY_SLOPE, twin: bool = False):
    """y = lrelu(maxpool2x2(x)); with `twin` (fp16 x) returns (y, bf16 copy of y) from the same launch."""
    n, h, w, c = x.shape
    y = torch.empty((n, h // 2, w // 2, c), device=x.device, dtype=x.dtype)
    a = STRUCTS["pb_pool_fwd_args"]()
    a.x, a.y = _ptr(x), _ptr(y)
    a.N, a.H, a.W, a.C, a.slope, a.act_dtype = n, h, w, c, slope, pb_dtype(x.dtype)
    y2 = None
    if twin:
        assert x.dtype == torch.float16
        y2 = torch.empty(y.shape, device=x.device, dtype=torch.bfloat16)
        a.y2 = _ptr(y2)
    _lib.call("pb_maxpool_lrelu_fwd", a, _stream())
    return (y, y2) if twin else y


def maxpool_lrelu_bwd(x: torch.Tensor, gy: torch.Tensor, mask: Optional[torch.Tensor],
                      slope: float = LEAKY_SLOPE) -> Tuple[torch.Tensor, torch.Tensor]:
    n, h, w, c = x.shape
    gx = torch.empty(x.shape, device=x.device, dtype=gy.dtype)
    gxm = torch.empty_like(gx)
    a = STRUCTS["pb_pool_bwd_args"]()
    a.x, a.gy, a.mask, a.gx, a.gx_masked = _ptr(x), _ptr(gy), _ptr(mask), _ptr(gx), _ptr(gxm)
    a.N, a.H, a.W, a.C, a.slope, a.act_dtype = n, h, w, c, slope, pb_dtype(gy.dtype)
    if x.dtype != gy.dtype:
        a.x_dtype = pb_dtype(x.dtype)     # fp16 forward activations against bf16 gradients
    _lib.call("pb_maxpool_lrelu_bwd", a, _stream())
    return gx, gxm


def mse_loss_fwd_bwd(out: torch.Tensor, target: Optional[torch.Tensor], *, points: Optional[torch.Tensor] = None,
                     sigma: float = 3.0, accumulation_steps: int = 1, loss_scale: float = 1.0,
                     want_grad_nchw: bool = False, grad_nhwc_dtype: Optional[torch.dtype] = None, cpad: int = 0,
                     slope: float = LEAKY_SLOPE):
    """returns (loss_sum tensor[1] (sum of squared errors), grad_nchw|None, grad_nhwc|None).
    mean loss = loss_sum / out.numel() / accumulation_steps   (pytorch/train_pytorch.py:134-135)."""
    b, c, h, w = out.shape
    assert out.dtype == torch.float32 and out.is_contiguous()
    loss_sum = torch.zeros(1, device=out.device, dtype=torch.float32)
    g_nchw = torch.empty_like(out) if want_grad_nchw else None
    g_nhwc = None
    cp = max(cpad, c)
    if grad_nhwc_dtype is not None:
        g_nhwc = torch.empty((b, h, w, cp), device=out.device, dtype=grad_nhwc_dtype)
    a = STRUCTS["pb_mse_args"]()
    a.out, a.target, a.points = _ptr(out), _ptr(target), _ptr(points)
    if target is not None:
        assert target.shape == out.shape and target.dtype == torch.float32 and target.is_contiguous()
    a.sigma = sigma
    a.loss_sum, a.loss_sum_f64, a.grad_nchw, a.grad_nhwc = _ptr(loss_sum), None, _ptr(g_nchw), _ptr(g_nhwc)
    a.B, a.C, a.H, a.W, a.Cpad = b, c, h, w, cp
    a.grad_scale = 2.0 * loss_scale / (out.numel() * accumulation_steps)
    a.slope = slope
    a.act_dtype = pb_dtype(grad_nhwc_dtype or torch.float32)
    _lib.call("pb_mse_loss_fwd_bwd", a, _stream())
    return loss_sum, g_nchw, g_nhwc


def minmax_mse_eligible(c: int, h: int, w: int, dtype: torch.dtype, cpad: int) -> bool:
    """shapes pb_minmax_mse_fwd_bwd takes (the bf16 NHWC gradient tile walk of the MSE kernel)."""
    cp = max(cpad, c)
    return dtype == torch.bfloat16 and cp % 8 == 0 and cp <= 64 and (h * w) % 128 == 0


def minmax_mse_fwd_bwd(x: torch.Tensor, target: Optional[torch.Tensor], *, points: Optional[torch.Tensor] = None,
                       sigma: float = 3.0, accumulation_steps: int = 1, loss_scale: float = 1.0, cpad: int = 0,
                       slope: float = LEAKY_SLOPE, numel: Optional[int] = None, loss_sum: Optional[torch.Tensor] = None,
                       grad_out: Optional[torch.Tensor] = None):
    """normalize_between_0_and_1 (pytorch/VITs.py:55-58) + MSELoss + both backwards + LeakyReLU' of the layer that
    produced x, fused: returns (loss_sum tensor[1], grad_nhwc bf16 [B,H,W,cpad]) for the PRE-normalisation heatmaps
    x [B,C,H,W] fp32; mean loss = loss_sum / x.numel() / accumulation_steps."""
    b, c, h, w = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    cp = max(cpad, c)
    # numel / loss_sum / grad_out: x is one GROUP of a larger batch whose loss is a mean over `numel` elements (the four
    # per-view decoder calls of VIT4CamerasBaseLine each normalise their own tensor, VITs.py:301-304); the sum of squared
    # errors accumulates into the shared loss_sum and the gradient lands in the caller's slice
    if loss_sum is None:
        loss_sum = torch.zeros(1, device=x.device, dtype=torch.float32)
    g = torch.empty((b, h, w, cp), device=x.device, dtype=torch.bfloat16) if grad_out is None else grad_out
    assert g.shape == (b, h, w, cp) and g.dtype == torch.bfloat16 and g.is_contiguous()
    scratch = torch.empty(16, device=x.device, dtype=torch.int32)      # 64 bytes, 16-byte aligned
    a = STRUCTS["pb_minmax_mse_args"]()
    if target is not None:
        assert target.shape == x.shape and target.dtype == torch.float32 and target.is_contiguous()
    a.x, a.target, a.points, a.sigma = _ptr(x), _ptr(target), _ptr(points), sigma
    a.loss_sum, a.grad_nhwc, a.scratch = _ptr(loss_sum), _ptr(g), _ptr(scratch)
    a.B, a.C, a.H, a.W, a.Cpad = b, c, h, w, cp
    a.grad_scale = 2.0 * loss_scale / ((numel if numel is not None else x.numel()) * accumulation_steps)
    a.slope = slope
    _lib.call("pb_minmax_mse_fwd_bwd", a, _stream())
    return loss_sum, g


def grad_ingest(grad_nchw: torch.Tensor, out_nchw: Optional[torch.Tensor], dtype: torch.dtype, cpad: int = 0,
                slope: float = LEAKY_SLOPE) -> torch.Tensor:
    b, c, h, w = grad_nchw.shape
    cp = max(cpad, c)
    g = torch.empty((b, h, w, cp), device=grad_nchw.device, dtype=dtype)
    a = STRUCTS["pb_grad_ingest_args"]()
    gc = grad_nchw.contiguous().float()
    a.grad_nchw, a.out_nchw, a.grad_nhwc = _ptr(gc), _ptr(out_nchw), _ptr(g)
    a.B, a.C, a.H, a.W, a.Cpad, a.slope, a.act_dtype = b, c, h, w, cp, slope, pb_dtype(dtype)
    _lib.call("pb_grad_ingest", a, _stream())
    return g


def gaussian_heatmaps(points: torch.Tensor, sigma: float = 3.0, size: Tuple[int, int] = (192, 192)) -> torch.Tensor:
    """(B,C,2) [x,y] float32 CUDA -> (B,C,H,W) float32 (tensorflow/simple_data_generator.py:119-136)."""
    b, c, _ = points.shape
    out = torch.empty((b, c, size[0], size[1]), device=points.device, dtype=torch.float32)
    a = STRUCTS["pb_gaussian_args"]()
    p = points.contiguous().float()
    a.points, a.out, a.BC, a.H, a.W, a.sigma = _ptr(p), _ptr(out), b * c, size[0], size[1], sigma
    _lib.call("pb_gaussian_heatmaps", a, _stream())
    return out


def _peaks(fn: str, hm: torch.Tensor, layout: str, want_values: bool):
    if layout == "nhwc":
        n, h, w, c = hm.shape
        sn, sy, sx, sc = hm.stride()
    else:
        n, c, h, w = hm.shape
        sn, sc, sy, sx = hm.stride()
    peaks = torch.empty((n, c, 2), device=hm.device, dtype=torch.float32)
    values = torch.empty((n, c), device=hm.device, dtype=torch.float32) if want_values else None
    a = STRUCTS["pb_peaks_args"]()
    a.heatmaps, a.peaks, a.values = _ptr(hm), _ptr(peaks), _ptr(values)
    a.N, a.C, a.H, a.W = n, c, h, w
    a.stride_n, a.stride_c, a.stride_y, a.stride_x = sn, sc, sy, sx
    a.dtype = pb_dtype(hm.dtype)
    _lib.call(fn, a, _stream())
    return (peaks, values) if want_values else peaks


def peaks_argmax(hm: torch.Tensor, layout: str = "nchw", want_values: bool = False):
    """Augmentor.tf_find_peaks (pytorch/Augmentor.py:105-148) on a CUDA tensor; zero-copy for both
    (N,H,W,C) and (N,C,H,W) (any strides with unit stride along x or c)."""
    if not (hm.stride(-1) == 1 or (layout == "nchw" and hm.stride(1) == 1) or (layout == "nhwc" and hm.stride(2) == 1)):
        hm = hm.contiguous()
    return _peaks("pb_peaks_argmax", hm, layout, want_values)


def peaks_softargmax(hm: torch.Tensor, layout: str = "nchw"):
    """find_peaks_soft_argmax (pytorch/utils.py:47-83) on a CUDA tensor."""
    return _peaks("pb_peaks_softargmax", hm, layout, False)


def affine_nearest(x: torch.Tensor, theta: torch.Tensor, flips: Optional[torch.Tensor] = None,
                   src_index: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """torchvision F.affine (nearest, zero fill) + optional h/v flips (pytorch/Datagenerators.py:170-182) on
    x[Nsrc,C,H,W] (fp32, or uint8 = ToTensor's /255 folded in) with per-sample inverse matrices theta[B,6]
    (torchvision's _get_inverse_affine_matrix convention), flips[B] (bit 0 h, bit 1 v) and the dataset row of
    every output sample src_index[B] (None: B == Nsrc, identity).  Returns fp32 [B,C,H,W]."""
    _ptr(x), _ptr(theta)   # CPU tensors raise here: there is no CPU fallback
    assert x.dtype in (torch.float32, torch.uint8) and x.is_contiguous() and x.dim() == 4
    nsrc, c, h, w = x.shape
    b = theta.shape[0]
    assert theta.shape == (b, 6) and theta.dtype == torch.float32 and theta.is_contiguous() and theta.is_cuda
    if src_index is None:
        assert b == nsrc, "affine_nearest: theta rows must match the batch when no src_index is given"
    else:
        assert src_index.shape == (b,) and src_index.dtype == torch.int32 and src_index.is_cuda
    if flips is not None:
        assert flips.shape == (b,) and flips.dtype == torch.int32 and flips.is_cuda
    if out is None:
        out = torch.empty((b, c, h, w), device=x.device, dtype=torch.float32)
    else:
        assert out.shape == (b, c, h, w) and out.dtype == torch.float32 and out.is_contiguous() and out.is_cuda
    if b == 0:
        return out
    a = STRUCTS["pb_affine_nearest_args"]()
    setattr(a, "in", _ptr(x))
    a.out, a.theta, a.flips, a.src_index = _ptr(out), _ptr(theta), _ptr(flips), _ptr(src_index)
    a.B, a.C, a.H, a.W, a.in_u8 = b, c, h, w, int(x.dtype == torch.uint8)
    _lib.call("pb_affine_nearest", a, _stream())
    return out


# every C-ABI call that writes parameter memory bumps this; cached packed operands (engine.Layer.packed) carry the
# value they were built at, because writes through the C ABI or through ``param.data`` do not move torch's version
# counter.  Code that writes weights behind the package's back calls model.invalidate_packed_weights().
_WEIGHTS_GENERATION = 0


def weights_generation() -> int:
    return _WEIGHTS_GENERATION


def bump_weights_generation() -> int:
    global _WEIGHTS_GENERATION
    _WEIGHTS_GENERATION += 1
    return _WEIGHTS_GENERATION


def adam_step(param: torch.Tensor, grad: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, step: int,
              lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
              grad_scale: float = 1.0, found_inf: Optional[torch.Tensor] = None,
              step_dev: Optional[torch.Tensor] = None, lr_dev: Optional[torch.Tensor] = None,
              inc_step: bool = False) -> None:
    """step_dev (int32[1], the number of COMPLETED steps) / lr_dev (float32[1]): the CUDA-graph-safe form -- the launch
    reads the step number and the learning rate from device memory instead of freezing them as kernel arguments;
    inc_step increments step_dev behind this launch (the last Adam launch of a step)."""
    a = STRUCTS["pb_adam_args"]()
    if step_dev is not None:
        assert step_dev.dtype == torch.int32 and step_dev.numel() == 1
        a.step_dev, a.inc_step = _ptr(step_dev), int(inc_step)
    if lr_dev is not None:
        assert lr_dev.dtype == torch.float32 and lr_dev.numel() == 1
        a.lr_dev = _ptr(lr_dev)
    a.param, a.grad, a.exp_avg, a.exp_avg_sq = _ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq)
    a.n = param.numel()
    a.lr, a.beta1, a.beta2, a.eps, a.weight_decay = lr, betas[0], betas[1], eps, weight_decay
    a.grad_scale, a.step, a.found_inf = grad_scale, step, _ptr(found_inf)
    _lib.call("pb_adam_step", a, _stream())
    bump_weights_generation()


def add(a_t: torch.Tensor, b_t: Optional[torch.Tensor], out: Optional[torch.Tensor] = None,
        mask: Optional[torch.Tensor] = None, slope: float = LEAKY_SLOPE) -> torch.Tensor:
    """out = (a + b) * lrelu'(mask)  -- elementwise, NHWC (last dim = channels)."""
    out = torch.empty_like(a_t) if out is None else out
    a = STRUCTS["pb_add_args"]()
    a.a, a.b, a.out, a.n, a.act_dtype = _ptr(a_t), _ptr(b_t), _ptr(out), a_t.numel(), pb_dtype(a_t.dtype)
    a.mask, a.C, a.slope = _ptr(mask), a_t.shape[-1], slope
    _lib.call("pb_add", a, _stream())
    return out


def ftl(x: torch.Tensor, mats: torch.Tensor, out: torch.Tensor, *, kin: int, kout: int, groups: int,
        in_batch_stride: int, out_batch_stride: int, in_batch_mod: int = 0, accumulate: bool = False) -> torch.Tensor:
    """FTL / InvFTL (pytorch/CNNs.py:322-345) on flat NCHW buffers: out[b][kout*m + i] (+)= sum_j mats[b][i][j] *
    x[b % in_batch_mod][kin*m + j]; mats [B, kout, kin] fp32.  See pb_ftl."""
    assert mats.dtype == torch.float32 and mats.is_contiguous() and mats.shape[1:] == (kout, kin)
    assert x.dtype == out.dtype and x.is_contiguous() and out.is_contiguous()
    a = STRUCTS["pb_ftl_args"]()
    setattr(a, "in", _ptr(x))
    a.out, a.mats, a.B, a.groups, a.kin, a.kout = _ptr(out), _ptr(mats), mats.shape[0], groups, kin, kout
    a.in_batch_stride, a.out_batch_stride, a.in_batch_mod = in_batch_stride, out_batch_stride, in_batch_mod
    a.accumulate, a.act_dtype = int(accumulate), pb_dtype(x.dtype)
    _lib.call("pb_ftl", a, _stream())
    return out


def _bn_nblk(rows_per_group: int) -> int:
    return max(1, min(64, rows_per_group // 256))


def batchnorm_fwd(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, running_mean: torch.Tensor,
                  running_var: torch.Tensor, *, groups: int, channels: int, training: bool, relu: bool = True,
                  eps: float = 1e-5, momentum: float = 0.1):
    """nn.BatchNorm2d (+ ReLU) on NHWC rows x [groups*rows_per_group, Cs] (Cs >= channels: zero padding), statistics per
    (group, channel); returns (y, save_mean [groups, C], save_rstd).  Training updates the running statistics in
    place, group after group (pytorch/CNNs.py:302-309: one module called on four tensors)."""
    rows, cs = x.shape[0], x.shape[1]
    assert x.dim() == 2 and x.is_contiguous() and rows % groups == 0
    rpg = rows // groups
    y = torch.empty_like(x)
    mean = torch.empty((groups, channels), device=x.device, dtype=torch.float32)
    rstd = torch.empty_like(mean)
    nblk = _bn_nblk(rpg)
    a = STRUCTS["pb_batchnorm_fwd_args"]()
    a.x, a.y, a.gamma, a.beta = _ptr(x), _ptr(y), _ptr(gamma), _ptr(beta)
    a.running_mean, a.running_var, a.save_mean, a.save_rstd = _ptr(running_mean), _ptr(running_var), _ptr(mean), _ptr(rstd)
    ws = torch.empty(groups * nblk * 2 * channels, device=x.device, dtype=torch.float32) if training else None
    a.partial = _ptr(ws)
    a.groups, a.rows_per_group, a.C, a.Cs, a.nblk = groups, rpg, channels, cs, nblk
    a.eps, a.momentum, a.training, a.relu, a.act_dtype = eps, momentum, int(training), int(relu), pb_dtype(x.dtype)
    _lib.call("pb_batchnorm_fwd", a, _stream())
    return y, mean, rstd


def batchnorm_bwd(x: torch.Tensor, y: Optional[torch.Tensor], gy: torch.Tensor, gamma: torch.Tensor, mean: torch.Tensor,
                  rstd: torch.Tensor, dgamma: torch.Tensor, dbeta: torch.Tensor, *, groups: int, channels: int,
                  relu: bool = True, beta_acc: float = 0.0) -> torch.Tensor:
    """training-mode BatchNorm (+ ReLU) backward; dgamma / dbeta = beta_acc * old + the sum over all groups."""
    rows, cs = x.shape[0], x.shape[1]
    rpg = rows // groups
    gx = torch.empty_like(x)
    nblk = _bn_nblk(rpg)
    ws = torch.empty(groups * nblk * 2 * channels, device=x.device, dtype=torch.float32)
    a = STRUCTS["pb_batchnorm_bwd_args"]()
    a.x, a.y, a.gy, a.gx, a.gamma = _ptr(x), _ptr(y), _ptr(gy), _ptr(gx), _ptr(gamma)
    a.save_mean, a.save_rstd, a.dgamma, a.dbeta, a.partial = _ptr(mean), _ptr(rstd), _ptr(dgamma), _ptr(dbeta), _ptr(ws)
    a.groups, a.rows_per_group, a.C, a.Cs, a.nblk = groups, rpg, channels, cs, nblk
    a.beta_acc, a.relu, a.act_dtype = beta_acc, int(relu), pb_dtype(x.dtype)
    _lib.call("pb_batchnorm_bwd", a, _stream())
    return gx


def note_launches(n: int) -> None:
    """a CUDA-graph replay launched `n` kernels of this library (counted when the graph was captured)."""
    _lib.load().pb_note_launches(ctypes.c_int32(int(n)))


def profiling_active() -> bool:
    return _PROFILE is not None


def launch_count() -> int:
    """CUDA kernels launched by libposeb200.so in this process so far."""
    lib = _lib.load()
    n = ctypes.c_ulonglong(0)
    lib.pb_launch_count(ctypes.byref(n))
    return int(n.value)


def device_info() -> Tuple[str, int, int]:
    lib = _lib.load()
    name = ctypes.create_string_buffer(128)
    sm, n_sm = ctypes.c_int(0), ctypes.c_int(0)
    rc = lib.pb_device_info(name, 128, ctypes.byref(sm), ctypes.byref(n_sm))
    if rc != 0:
        raise _lib.PoseB200Error("pb_device_info", rc, _lib.last_error())
    return name.value.decode(), sm.value, n_sm.value
