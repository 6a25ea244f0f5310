"""Execution engine of the ViT-encoder heatmap model (pytorch/pytorch_vit_encoder.py,
pytorch/VITs.py:13-58,197-229): schedules the C-ABI kernels for forward and backward.

Tokens are a row-major [B*S, dim] matrix in ``act_dtype``; every nn.Linear is the 1-tap
gather-convolution (tcgen05 GEMM in bf16 mode where the shape tiles) with bias / GELU /
residual-add fused into its epilogue; LayerNorm, attention, GELU' and the batch-global min/max
normalisation are the kernels of csrc/vit.cu.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

import torch
from torch import nn

from . import ops, vit_ops
from .engine import ConvStack, Layer
from .ops import Contraction, PB_ACT_GELU, PB_ACT_NONE

ParamSink = Callable[[str, torch.Tensor], Tuple[torch.Tensor, float]]


_FP16_MSG = ("the ViT models run in 'bf16' or 'fp32' (their normalised [0, 1] heatmaps meet the bf16 parity gate; the "
             "'fp16' forward exists for the conv heatmap networks)")


class TokenStack(ConvStack):
    """Shared machinery of the token-matrix engines: nn.Linear as 1-tap contractions, LayerNorm, and the reference's
    ``Transformer`` (pytorch/pytorch_vit_encoder.py:81-105: depth x [pre-LN MHA + res, pre-LN MLP + res] -> LN) for
    any width -- 256 in the encoder, 1280 in VIT4CamerasBaseLine's cross-attention blocks."""

    def __init__(self, precision: str):
        if precision == "fp16":
            raise ValueError(_FP16_MSG)
        super().__init__(precision)

    # ---- linear helpers (rows x cin) -> (rows x cout) ----------------------------------------
    def _lin(self, name: str, x: torch.Tensor, *, act: int = PB_ACT_NONE, add1=None, pre_out=None) -> torch.Tensor:
        layer = self.layers[name]
        s = layer.spec
        rows = x.shape[0]
        impl = self.impl_for(s, "fwd")
        if impl == "tc":
            from . import tc_support
            w = layer.packed("oi", torch.bfloat16, tc_support.pad_n(s.cout))
        else:
            w = layer.packed("io", torch.float32)
        y = ops.conv(impl, x, w, s.fwd_taps(), 1, 1, rows, s.cin, 1, rows, s.cout, bias=layer.module.bias, act=act,
                     add1=add1, pre_out=pre_out, act_dtype=self.act_dtype)
        return y.view(rows, s.cout)

    def _lin_dgrad(self, name: str, g: torch.Tensor, add0=None) -> torch.Tensor:
        layer = self.layers[name]
        s = layer.spec
        rows = g.shape[0]
        impl = self.impl_for(s, "dgrad")
        w = layer.packed("io", torch.bfloat16) if impl == "tc" else layer.packed("oi", torch.float32)
        y = ops.conv(impl, g, w, s.dgrad_taps(), 1, 1, rows, s.cout, 1, rows, s.cin, add0=add0,
                     act_dtype=self.act_dtype)
        return y.view(rows, s.cin)

    def _lin_wgrad(self, name: str, a_in: torch.Tensor, g: torch.Tensor, sink: ParamSink, bias_partial=None) -> None:
        """bias_partial = (partial [nblk, cout] fp32, nblk): column sums of g already produced by the kernel that
        wrote g (GELU backward) -- the bias gradient is then one small fold instead of another pass over g."""
        layer = self.layers[name]
        s = layer.spec
        rows = g.shape[0]
        impl = self.impl_for(s, "wgrad")
        dw, beta = sink(name + ".weight", layer.module.weight)
        db = None
        if layer.module.bias is not None:
            db, _ = sink(name + ".bias", layer.module.bias)
        if db is not None and bias_partial is not None:
            vit_ops.colsum(bias_partial[0], db, bias_partial[1], s.cout, beta=beta)
            ops.wgrad(impl, s, a_in, g, 1, 1, rows, dw, None, act_dtype=self.act_dtype, beta=beta,
                      workspace=self._workspace(s, rows, g.device))
        else:
            ops.wgrad(impl, s, a_in, g, 1, 1, rows, dw, db, act_dtype=self.act_dtype, beta=beta,
                      workspace=self._workspace(s, rows, g.device))
        done = getattr(sink, "done", None)
        if done is not None:
            done(name + ".weight")
            if db is not None:
                done(name + ".bias")

    def _ln_bwd(self, sink: ParamSink, prefix: str, ln: nn.LayerNorm, x, gy, mean, rstd, gx_add=None,
                want_colsum: bool = False):
        """want_colsum: returns (gx, column-sum partials of gx or None), see vit_ops.layernorm_bwd."""
        dg, beta = sink(prefix + ".weight", ln.weight)
        dbt, _ = sink(prefix + ".bias", ln.bias)
        gx = vit_ops.layernorm_bwd(x, gy, ln.weight, mean, rstd, dg, dbt, gx_add=gx_add, beta=beta,
                                   want_colsum=want_colsum)
        done = getattr(sink, "done", None)
        if done is not None:
            done(prefix + ".weight")
            done(prefix + ".bias")
        return gx

    # ---- the reference's Transformer ------------------------------------------------------------
    def tr_register(self, prefix: str, transformer: nn.Module) -> None:
        attn0 = transformer.layers[0][0]
        dim = attn0.to_qkv.in_features
        inner = attn0.to_qkv.out_features // 3
        for l, (attn, ff) in enumerate(transformer.layers):
            p = f"{prefix}layers.{l}."
            self.layers[p + "0.to_qkv"] = Layer(p + "0.to_qkv", attn.to_qkv, Contraction("linear", dim, 3 * inner))
            self.layers[p + "0.to_out.0"] = Layer(p + "0.to_out.0", attn.to_out[0], Contraction("linear", inner, dim))
            hid = ff.net[1].out_features
            self.layers[p + "1.net.1"] = Layer(p + "1.net.1", ff.net[1], Contraction("linear", dim, hid))
            self.layers[p + "1.net.4"] = Layer(p + "1.net.4", ff.net[4], Contraction("linear", hid, dim))

    def tr_forward(self, prefix: str, transformer: nn.Module, t: torch.Tensor, b: int, s_tok: int, save: bool):
        """t [b*s_tok, dim] -> (LN(blocks(t)), saved)."""
        saved = {"layers": []}
        for l, (attn, ff) in enumerate(transformer.layers):
            p = f"{prefix}layers.{l}."
            heads = attn.heads
            dh = attn.to_qkv.out_features // 3 // heads
            h, m1, r1 = vit_ops.layernorm_fwd(t, attn.norm.weight, attn.norm.bias, save=save)
            qkv = self._lin(p + "0.to_qkv", h)
            o, probs = vit_ops.attention_fwd(qkv, b, s_tok, heads, dh, attn.scale)
            t_mid = self._lin(p + "0.to_out.0", o, add1=t)
            h2, m2, r2 = vit_ops.layernorm_fwd(t_mid, ff.net[0].weight, ff.net[0].bias, save=save)
            u_pre = torch.empty((h2.shape[0], ff.net[1].out_features), device=h2.device, dtype=self.act_dtype) if save else None
            u = self._lin(p + "1.net.1", h2, act=PB_ACT_GELU, pre_out=u_pre)
            t_out = self._lin(p + "1.net.4", u, add1=t_mid)
            if save:
                saved["layers"].append((t, m1, r1, h, qkv, probs, o, t_mid, m2, r2, h2, u_pre, u))
            t = t_out
        tokens, mf, rf = vit_ops.layernorm_fwd(t, transformer.norm.weight, transformer.norm.bias, save=save)
        saved["final"] = (t, mf, rf)
        return tokens, (saved if save else None)

    def tr_backward(self, prefix: str, transformer: nn.Module, saved: dict, g_tokens: torch.Tensor, b: int, s_tok: int,
                    sink: ParamSink) -> torch.Tensor:
        """gradient w.r.t. the Transformer's input tokens."""
        # the residual-stream gradients g_t / g_tmid are the output gradients of the Linear that closes each block; the
        # LayerNorm backward that writes them also emits their column sums = those layers' bias gradients
        t, mf, rf = saved["final"]
        g_t, g_t_sums = self._ln_bwd(sink, prefix + "norm", transformer.norm, t, g_tokens, mf, rf, want_colsum=True)
        for l in range(len(transformer.layers) - 1, -1, -1):
            attn, ff = transformer.layers[l]
            p = f"{prefix}layers.{l}."
            heads = attn.heads
            dh = attn.to_qkv.out_features // 3 // heads
            t_in, m1, r1, h, qkv, probs, o, t_mid, m2, r2, h2, u_pre, u = saved["layers"][l]
            # t_out = fc2(u) + t_mid
            self._lin_wgrad(p + "1.net.4", u, g_t, sink, bias_partial=g_t_sums)
            g_u = self._lin_dgrad(p + "1.net.4", g_t)
            g_upre, bias_part = vit_ops.gelu_bwd(u_pre, g_u, want_colsum=True)
            self._lin_wgrad(p + "1.net.1", h2, g_upre, sink, bias_partial=bias_part)
            g_h2 = self._lin_dgrad(p + "1.net.1", g_upre)
            g_tmid, g_tmid_sums = self._ln_bwd(sink, p + "1.net.0", ff.net[0], t_mid, g_h2, m2, r2, gx_add=g_t,
                                               want_colsum=True)
            # t_mid = to_out(o) + t_in
            self._lin_wgrad(p + "0.to_out.0", o, g_tmid, sink, bias_partial=g_tmid_sums)
            g_o = self._lin_dgrad(p + "0.to_out.0", g_tmid)
            g_qkv = vit_ops.attention_bwd(qkv, probs, g_o, b, s_tok, heads, dh, attn.scale)
            self._lin_wgrad(p + "0.to_qkv", h, g_qkv, sink)
            g_h = self._lin_dgrad(p + "0.to_qkv", g_qkv)
            g_t, g_t_sums = self._ln_bwd(sink, p + "0.norm", attn.norm, t_in, g_h, m1, r1, gx_add=g_tmid,
                                         want_colsum=True)
            saved["layers"][l] = None
        return g_t


class VitEncoderEngine(TokenStack):
    """CustomViT: patchify -> Linear -> LN (+pos) -> depth x [pre-LN MHA + res, pre-LN MLP + res] -> LN."""

    def __init__(self, module: nn.Module, precision: str):
        super().__init__(precision)
        self.m = module
        self.dim = module.dim
        self.patch = module.patch_size
        self.depth = len(module.transformer.layers)
        self.layers["patch_to_embedding"] = Layer("patch_to_embedding", module.patch_to_embedding,
                                                  Contraction("linear", module.patch_dim, self.dim))
        self.tr_register("transformer.", module.transformer)

    # ---- forward -----------------------------------------------------------------------------
    def forward(self, img: torch.Tensor, save: bool):
        m = self.m
        b = img.shape[0]
        patches = vit_ops.patchify(img, self.patch, self.act_dtype)
        s_tok = patches.shape[0] // b
        saved: dict = {"b": b, "s": s_tok, "patches": patches}
        e = self._lin("patch_to_embedding", patches)
        pos = m.pos_embedding[0, :s_tok].contiguous()
        t, mean, rstd = vit_ops.layernorm_fwd(e, m.norm.weight, m.norm.bias, add=pos, save=save)
        saved["embed"] = (e, mean, rstd)
        tokens, s_tr = self.tr_forward("transformer.", m.transformer, t, b, s_tok, save)
        saved["tr"] = s_tr
        return tokens, (saved if save else None)

    # ---- backward ----------------------------------------------------------------------------
    def backward(self, saved: dict, g_tokens: torch.Tensor, sink: ParamSink) -> None:
        m = self.m
        b, s_tok = saved["b"], saved["s"]
        g_t = self.tr_backward("transformer.", m.transformer, saved["tr"], g_tokens, b, s_tok, sink)
        # t0 = LN(e) + pos_embedding (broadcast over the batch)
        e, mean, rstd = saved["embed"]
        dpos, beta = sink("pos_embedding", m.pos_embedding)
        vit_ops.colsum(g_t, dpos, b, s_tok * self.dim, beta=beta)
        done = getattr(sink, "done", None)
        if done is not None:
            done("pos_embedding")
        g_e = self._ln_bwd(sink, "norm", m.norm, e, g_t, mean, rstd)
        self._lin_wgrad("patch_to_embedding", saved["patches"], g_e, sink)


class CrossAttentionEngine(TokenStack):
    """CrossAttention (pytorch/VITs.py:235-250): Transformer(5*dim, depth 1, 4 heads, dim_head = mlp_dim = dim) ->
    LayerNorm(5*dim) -> Linear(5*dim -> dim) -> GELU; the caller's residual ``+ enc`` (VITs.py:297-300) rides in the
    last Linear's epilogue."""

    def __init__(self, module: nn.Module, precision: str):
        super().__init__(precision)
        self.m = module
        seq = module.layers
        self.tr, self.ln, self.lin = seq[0], seq[1], seq[2]
        self.tr_register("layers.0.", self.tr)
        self.layers["layers.2"] = Layer("layers.2", self.lin, Contraction("linear", self.lin.in_features,
                                                                          self.lin.out_features))

    def forward(self, x: torch.Tensor, residual: torch.Tensor, b: int, s_tok: int, save: bool):
        """x [b*s_tok, 5*dim] -> gelu(Linear(LN(Transformer(x)))) + residual."""
        t, s_tr = self.tr_forward("layers.0.", self.tr, x, b, s_tok, save)
        h, mean, rstd = vit_ops.layernorm_fwd(t, self.ln.weight, self.ln.bias, save=save)
        pre = torch.empty((h.shape[0], self.lin.out_features), device=h.device, dtype=self.act_dtype) if save else None
        y = self._lin("layers.2", h, act=PB_ACT_GELU, add1=residual, pre_out=pre)
        return y, ({"tr": s_tr, "ln": (t, mean, rstd), "h": h, "pre": pre} if save else None)

    def backward(self, saved: dict, g_y: torch.Tensor, b: int, s_tok: int, sink: ParamSink) -> torch.Tensor:
        """g_y: gradient w.r.t. the block's output (the residual branch's share is the caller's); returns the
        gradient w.r.t. x [b*s_tok, 5*dim]."""
        g_pre, bias_part = vit_ops.gelu_bwd(saved["pre"], g_y, want_colsum=True)
        self._lin_wgrad("layers.2", saved["h"], g_pre, sink, bias_partial=bias_part)
        g_h = self._lin_dgrad("layers.2", g_pre)
        t, mean, rstd = saved["ln"]
        g_t = self._ln_bwd(sink, "layers.1", self.ln, t, g_h, mean, rstd)
        return self.tr_backward("layers.0.", self.tr, saved["tr"], g_t, b, s_tok, sink)


class VitDecoderEngine(ConvStack):
    """CNN_Decoder (pytorch/VITs.py:13-58)."""

    names = ["deconv1", "deconv2", "deconv3", "deconv4"]

    def __init__(self, module: nn.Module, precision: str):
        if precision == "fp16":
            raise ValueError(_FP16_MSG)
        super().__init__(precision)
        self.m = module
        dim, cout = module.projection_dim, module.num_output_channels
        for i, name in enumerate(self.names):
            self.layers[name] = Layer(name, getattr(module, name),
                                      Contraction("convT2", dim, dim if i < 3 else cout, ksize=module.kernel_size))
        if module.kernel_size != 3:
            raise ValueError("CNN_Decoder: kernel size 3 is the only geometry padding=1/output_padding=1 supports")

    def out_cpad(self) -> int:
        last = self.layers["deconv4"]
        if self.impl_for(last.spec, "dgrad") == "tc":
            from . import tc_support
            return tc_support.pad_n(last.spec.cout)
        return last.spec.cout

    def forward(self, tokens: torch.Tensor, b: int, save: bool, normalize: bool = True, groups: int = 1):
        """normalize=False returns deconv4's output BEFORE normalize_between_0_and_1 (the fused train step folds the
        normalisation into its loss kernels, ops.minmax_mse_fwd_bwd).  groups: the batch is `groups` consecutive
        sub-batches that the reference pushes through the decoder in separate calls (the four views of
        VIT4CamerasBaseLine, VITs.py:301-304): the min/max normalisation is taken per sub-batch."""
        dim = self.m.projection_dim
        s_tok = tokens.shape[0] // b
        side = int(round(s_tok ** 0.5))
        # raw reinterpretation (B,144,256) -> (B,256,12,12) of VITs.py:39, expressed in NHWC
        x = vit_ops.batched_transpose(tokens.contiguous(), b, dim, s_tok).view(b, side, side, dim)
        saved: dict = {"b": b, "s": s_tok, "acts": []}
        ih = iw = side
        for i, name in enumerate(self.names):
            last = i == 3
            y, mask, _ = self.fwd_layer(self.layers[name], x, b, ih, iw, save=save and not last, out_nchw=last)
            if save:
                saved["acts"].append((x, mask, ih, iw))
            x = y
            ih, iw = 2 * ih, 2 * iw
        if not normalize:
            return x, (saved if save else None)
        if groups == 1:
            out, scratch = vit_ops.minmax_normalize_fwd(x)
            scratches = [scratch]
        else:
            out, per, scratches = torch.empty_like(x), b // groups, []
            for gi in range(groups):
                _, sc = vit_ops.minmax_normalize_fwd(x[gi * per:(gi + 1) * per], out=out[gi * per:(gi + 1) * per])
                scratches.append(sc)
        if save:
            saved["pre_norm"], saved["scratch"] = x, scratches
        return out, (saved if save else None)

    def backward(self, saved: dict, g_out: Optional[torch.Tensor], sink, need_input_grad: bool = True,
                 dc: Optional[torch.Tensor] = None):
        """g_out: gradient w.r.t. the normalised heatmaps (NCHW fp32); or dc: gradient w.r.t. deconv4's
        pre-activation, channel-padded NHWC, when the caller already went through the normalisation backward."""
        b = saved["b"]
        if dc is None:
            pre, scratches = saved["pre_norm"], saved["scratch"]
            g_out = g_out.contiguous().float()
            if len(scratches) == 1:
                g_pre = vit_ops.minmax_normalize_bwd(pre, g_out, scratches[0])
            else:
                g_pre, per = torch.empty_like(pre), b // len(scratches)
                for gi, sc in enumerate(scratches):
                    sl = slice(gi * per, (gi + 1) * per)
                    vit_ops.minmax_normalize_bwd(pre[sl], g_out[sl], sc, out=g_pre[sl])
            dc = ops.grad_ingest(g_pre, pre, self.act_dtype, cpad=self.out_cpad())
        g_in = None
        for i in (3, 2, 1, 0):
            layer = self.layers[self.names[i]]
            x, _mask, ih, iw = saved["acts"][i]
            self.wgrad_layer(layer, x, dc, b, ih, iw, sink)
            if i > 0:
                mask_prev = saved["acts"][i - 1][1]
                _, dc = self.dgrad_layer(layer, dc, b, ih, iw, want_g=False, mask_prev=mask_prev)
            elif need_input_grad:
                _, g_in = self.dgrad_layer(layer, dc, b, ih, iw, want_g=False, mask_prev=None)
        if g_in is None:
            return None
        dim = self.m.projection_dim
        return vit_ops.batched_transpose(g_in.view(b, saved["s"], dim), b, saved["s"], dim).view(b * saved["s"], dim)
