"""torch-tensor wrappers of the ViT-side C-ABI entry points (patchify, LayerNorm, attention, GELU',
min/max normalise, batched transpose).  CUDA tensors only."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import STRUCTS
from .ops import _ptr, _stream, pb_dtype


def patchify(img: torch.Tensor, patch: int, dtype: torch.dtype) -> torch.Tensor:
    """(B,C,H,W) fp32 -> (B*(H/p)*(W/p), C*p*p), feature order (c,ph,pw)  [pytorch_vit_encoder.py:135-138]"""
    b, c, h, w = img.shape
    out = torch.empty((b * (h // patch) * (w // patch), c * patch * patch), device=img.device, dtype=dtype)
    a = STRUCTS["pb_patchify_args"]()
    a.img, a.patches = _ptr(img.contiguous().float()), _ptr(out)
    a.B, a.C, a.H, a.W, a.P, a.act_dtype = b, c, h, w, patch, pb_dtype(dtype)
    _lib.call("pb_patchify", a, _stream())
    return out


def layernorm_fwd(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, add: Optional[torch.Tensor] = None,
                  eps: float = 1e-5, save: bool = True):
    rows, dim = x.shape
    y = torch.empty_like(x)
    mean = torch.empty(rows, device=x.device, dtype=torch.float32) if save else None
    rstd = torch.empty(rows, device=x.device, dtype=torch.float32) if save else None
    a = STRUCTS["pb_layernorm_fwd_args"]()
    a.x, a.gamma, a.beta, a.add, a.y, a.mean, a.rstd = _ptr(x), _ptr(gamma), _ptr(beta), _ptr(add), _ptr(y), _ptr(
        mean), _ptr(rstd)
    a.rows, a.dim, a.add_rows = rows, dim, (add.numel() // dim if add is not None else 0)
    a.eps, a.act_dtype = eps, pb_dtype(x.dtype)
    _lib.call("pb_layernorm_fwd", a, _stream())
    return y, mean, rstd


def colsum(partial: torch.Tensor, out: torch.Tensor, nblk: int, dim: int, alpha: float = 1.0, beta: float = 0.0,
           partial2: Optional[torch.Tensor] = None, out2: Optional[torch.Tensor] = None):
    """out[j] = beta*out[j] + alpha * sum_b partial[b][j]; (partial2, out2): a second problem of the same shape in
    the same launch."""
    a = STRUCTS["pb_colsum_args"]()
    a.partial, a.out, a.nblk, a.dim, a.alpha, a.beta = _ptr(partial), _ptr(out), nblk, dim, alpha, beta
    a.partial2, a.out2 = _ptr(partial2), _ptr(out2)
    a.in_dtype = pb_dtype(partial.dtype)
    _lib.call("pb_colsum", a, _stream())


def layernorm_bwd(x: torch.Tensor, gy: torch.Tensor, gamma: torch.Tensor, mean: torch.Tensor, rstd: torch.Tensor,
                  dgamma: torch.Tensor, dbeta: torch.Tensor, gx_add: Optional[torch.Tensor] = None,
                  beta: float = 0.0, want_colsum: bool = False):
    """returns gx (+ gx_add); dgamma/dbeta = beta*old + new.
    want_colsum: returns (gx, (partial [nblk, dim] fp32, nblk) or None) -- per-block column sums of gx written by the
    same kernel (bf16, dim 256): the bias gradient of the nn.Linear whose output gradient gx is; None when the shape
    is outside that kernel."""
    rows, dim = x.shape
    nblk = max(1, min(296, (rows + 63) // 64))
    gx = torch.empty_like(x)
    pg = torch.empty((nblk, dim), device=x.device, dtype=torch.float32)
    pb = torch.empty((nblk, dim), device=x.device, dtype=torch.float32)
    a = STRUCTS["pb_layernorm_bwd_args"]()
    a.x, a.gy, a.gamma, a.mean, a.rstd, a.gx_add, a.gx = _ptr(x), _ptr(gy), _ptr(gamma), _ptr(mean), _ptr(rstd), _ptr(
        gx_add), _ptr(gx)
    a.dgamma_partial, a.dbeta_partial = _ptr(pg), _ptr(pb)
    a.rows, a.dim, a.nblk, a.act_dtype = rows, dim, nblk, pb_dtype(x.dtype)
    part = None
    if want_colsum and x.dtype == torch.bfloat16 and dim == 256 and all(
            t is None or t.data_ptr() % 16 == 0 for t in (x, gy, gx, gx_add)):
        px = torch.empty((nblk, dim), device=x.device, dtype=torch.float32)
        a.gx_colsum_partial = _ptr(px)
        part = (px, nblk)
    _lib.call("pb_layernorm_bwd", a, _stream())
    colsum(pg, dgamma, nblk, dim, beta=beta, partial2=pb, out2=dbeta)
    return (gx, part) if want_colsum else gx


def attention_fwd(qkv: torch.Tensor, b: int, s: int, h: int, d: int, scale: float):
    """qkv [B*S, 3*H*D] -> (out [B*S, H*D], probs fp32 [B,H,S,S])   [pytorch_vit_encoder.py:59-78]"""
    out = torch.empty((b * s, h * d), device=qkv.device, dtype=qkv.dtype)
    probs = torch.empty((b, h, s, s), device=qkv.device, dtype=torch.float32)
    a = STRUCTS["pb_attention_fwd_args"]()
    a.qkv, a.out, a.probs = _ptr(qkv), _ptr(out), _ptr(probs)
    a.B, a.S, a.H, a.D, a.scale, a.act_dtype = b, s, h, d, scale, pb_dtype(qkv.dtype)
    _lib.call("pb_attention_fwd", a, _stream())
    return out, probs


def attention_bwd(qkv: torch.Tensor, probs: torch.Tensor, gout: torch.Tensor, b: int, s: int, h: int, d: int,
                  scale: float) -> torch.Tensor:
    gqkv = torch.empty_like(qkv)
    ws = torch.empty_like(probs)
    a = STRUCTS["pb_attention_bwd_args"]()
    a.qkv, a.probs, a.gout, a.gqkv, a.dprobs_ws = _ptr(qkv), _ptr(probs), _ptr(gout), _ptr(gqkv), _ptr(ws)
    a.B, a.S, a.H, a.D, a.scale, a.act_dtype = b, s, h, d, scale, pb_dtype(qkv.dtype)
    _lib.call("pb_attention_bwd", a, _stream())
    return gqkv


def gelu_bwd(pre: torch.Tensor, gy: torch.Tensor, want_colsum: bool = False):
    """gx = gy * gelu'(pre).  want_colsum (bf16 [rows, dim]): returns (gx, (partial [nblk, dim] fp32, nblk)) whose
    column sums -- fold with ``colsum`` -- are the bias gradient of the nn.Linear in front of the GELU; None when the
    shape is outside that variant."""
    gx = torch.empty_like(pre)
    a = STRUCTS["pb_gelu_bwd_args"]()
    a.pre, a.gy, a.gx, a.n, a.act_dtype = _ptr(pre), _ptr(gy), _ptr(gx), pre.numel(), pb_dtype(pre.dtype)
    part = None
    if want_colsum and pre.dtype == torch.bfloat16 and pre.dim() == 2 and pre.shape[1] % 8 == 0 and pre.shape[1] <= 2048:
        rows, dim = pre.shape
        nblk = max(1, min(296, (rows + 15) // 16))
        buf = torch.empty((nblk, dim), device=pre.device, dtype=torch.float32)
        a.dim, a.colsum_partial, a.nblk = dim, _ptr(buf), nblk
        part = (buf, nblk)
    _lib.call("pb_gelu_bwd", a, _stream())
    return (gx, part) if want_colsum else gx


def minmax_normalize_fwd(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(x - x.min()) / (x.max() - x.min()) over the WHOLE tensor  [pytorch/VITs.py:55-58]"""
    assert x.dtype == torch.float32 and x.is_contiguous()
    y = torch.empty_like(x) if out is None else out
    assert y.shape == x.shape and y.dtype == torch.float32 and y.is_contiguous()
    scratch = torch.empty(4, device=x.device, dtype=torch.int32)
    a = STRUCTS["pb_minmax_norm_fwd_args"]()
    a.x, a.y, a.minmax, a.n = _ptr(x), _ptr(y), _ptr(scratch), x.numel()
    _lib.call("pb_minmax_normalize_fwd", a, _stream())
    return y, scratch


def minmax_normalize_bwd(x: torch.Tensor, gy: torch.Tensor, scratch: torch.Tensor,
                         out: Optional[torch.Tensor] = None) -> torch.Tensor:
    gx = torch.empty_like(x) if out is None else out
    assert gx.shape == x.shape and gx.dtype == torch.float32 and gx.is_contiguous()
    ws = torch.empty(4, device=x.device, dtype=torch.int64)
    a = STRUCTS["pb_minmax_norm_bwd_args"]()
    a.x, a.gy, a.minmax, a.gx, a.scratch, a.n = _ptr(x), _ptr(gy.contiguous().float()), _ptr(scratch), _ptr(gx), _ptr(
        ws), x.numel()
    _lib.call("pb_minmax_normalize_bwd", a, _stream())
    return gx


class _MinMaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        xc = x.contiguous().float()
        y, scratch = minmax_normalize_fwd(xc)
        ctx.save_for_backward(xc, scratch)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, scratch = ctx.saved_tensors
        return minmax_normalize_bwd(x, gy, scratch)


def minmax_normalize(x: torch.Tensor) -> torch.Tensor:
    """differentiable normalize_between_0_and_1 (CNNs.py:131-134, VITs.py:55-58) on a CUDA tensor."""
    if not x.is_cuda:
        raise RuntimeError("minmax_normalize: CPU tensor (there is no CPU fallback)")
    return _MinMaxFn.apply(x)


def batched_transpose(x: torch.Tensor, batch: int, rows: int, cols: int) -> torch.Tensor:
    """out[b][j][i] = x[b][i][j] for `batch` row-major [rows x cols] matrices (the raw
    (B,144,256)->(B,256,12,12) reinterpretation of CNN_Decoder.forward, VITs.py:39, in NHWC terms)."""
    out = torch.empty_like(x)
    a = STRUCTS["pb_transpose_args"]()
    a.x, a.y, a.batch, a.rows, a.cols, a.act_dtype = _ptr(x), _ptr(out), batch, rows, cols, pb_dtype(x.dtype)
    _lib.call("pb_batched_transpose", a, _stream())
    return out


def colblock(src: torch.Tensor, dst: torch.Tensor, *, rows: int, ncols: int, src_row_stride: int, dst_row_stride: int,
             src_col0: int = 0, dst_col0: int = 0, src_rows_mod: int = 0, nfold: int = 1, fold_stride: int = 0,
             accumulate: bool = False) -> torch.Tensor:
    """dst[r, dst_col0 + j] (+)= sum_f src[f*fold_stride + (r % src_rows_mod)*src_row_stride + src_col0 + j]
    (pb_colblock): concatenation / split along features, view broadcast and view sums of the multi-camera models in
    one strided pass.  src / dst are flat views of the same dtype; strides and offsets in elements."""
    assert src.dtype == dst.dtype and src.is_contiguous() and dst.is_contiguous()
    a = STRUCTS["pb_colblock_args"]()
    a.src, a.dst = _ptr(src), _ptr(dst)
    a.rows, a.ncols = rows, ncols
    a.src_row_stride, a.dst_row_stride, a.src_col0, a.dst_col0 = src_row_stride, dst_row_stride, src_col0, dst_col0
    a.src_rows_mod, a.fold_stride, a.nfold, a.accumulate = src_rows_mod, fold_stride, nfold, int(accumulate)
    a.dtype = pb_dtype(src.dtype)
    _lib.call("pb_colblock", a, _stream())
    return dst
