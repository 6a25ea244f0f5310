"""Drop-in ViT-encoder / conv-transpose-decoder heatmap model (pytorch/VITs.py:13-58,197-229).

Same constructors, attribute names and state_dict keys as the reference (``vit_encoder.*``,
``cnn_decoder.deconv{1..4}.*``; 104 entries), plus the four-camera model ``VIT4CamerasBaseLine`` with its
``CrossAttention`` blocks (pytorch/VITs.py:235-306).  The dead classes of the reference file
(PositionalEncoding, TransformerBlock, ViTEncoder, TransformerDecoder, VIT_encoder_decoder -- not reachable from
Network.py) are not reproduced.
"""
from __future__ import annotations

import os

from typing import Dict, Optional, Tuple

import torch
from torch import nn

from . import ops, vit_ops
from .pytorch_vit_encoder import Attention, CustomViT, FeedForward, Transformer  # noqa: F401 (reference re-exports)


class _VitDecoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, need, tokens, *params):
        eng = module._engine()
        b = tokens.shape[0]
        t2d = tokens.reshape(-1, tokens.shape[-1]).contiguous().to(eng.act_dtype)
        out, saved = eng.forward(t2d, b, save=need)
        ctx.module, ctx.saved, ctx.need_x, ctx.shape = module, saved, tokens.requires_grad, tokens.shape
        return out

    @staticmethod
    def backward(ctx, g):
        module = ctx.module
        eng = module._engine()
        store: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}

        def sink(name: str):
            m = getattr(module, name)
            store[name] = (torch.empty_like(m.weight), torch.empty_like(m.bias))
            return store[name][0], store[name][1], 0.0

        g_tok = eng.backward(ctx.saved, g.contiguous().float(), sink, need_input_grad=ctx.need_x)
        ctx.saved = None
        grads = []
        for name in eng.names:
            grads.extend(store[name])
        gx = g_tok.view(ctx.shape) if g_tok is not None else None
        return (None, None, gx, *grads)


class CNN_Decoder(nn.Module):
    """pytorch/VITs.py:13-58: raw reshape (B,144,dim)->(B,dim,12,12), 4 x [convT(k3,s2,p1,op1)+LeakyReLU],
    then a min/max normalisation taken over the whole batch tensor."""

    def __init__(self, num_output_channels, kernel_size, num_base_filters, projection_dim, precision: str = "bf16"):
        super().__init__()
        self.num_output_channels = num_output_channels
        self.kernel_size = kernel_size
        self.num_base_filters = num_base_filters
        self.projection_dim = projection_dim
        self.precision = precision
        for i in range(1, 5):
            setattr(self, f"deconv{i}", nn.ConvTranspose2d(
                in_channels=projection_dim, out_channels=projection_dim if i < 4 else num_output_channels,
                kernel_size=kernel_size, stride=2, padding=1, output_padding=1))
        self.leakyrelu = nn.LeakyReLU(0.1)

    def _engine(self):
        from .vit_engine import VitDecoderEngine
        eng = self.__dict__.get("_eng")
        if eng is None or eng.precision != self.precision:
            eng = VitDecoderEngine(self, self.precision)
            self.__dict__["_eng"] = eng
        return eng

    def set_precision(self, precision: str):
        self.precision = precision
        return self

    def invalidate_packed_weights(self):
        eng = self.__dict__.get("_eng")
        if eng is not None:
            eng.invalidate()

    def repack_weights(self):
        """weights changed in place (fused Adam): refresh every packed operand with one launch."""
        eng = self.__dict__.get("_eng")
        if eng is not None and not eng.repack_all():
            eng.invalidate()

    def get_conv2d_transpose(self, in_channels, out_channels, stride):
        conv = nn.ConvTranspose2d(in_channels=in_channels, out_channels=out_channels, kernel_size=self.kernel_size,
                                  stride=stride, padding=1, output_padding=1)
        nn.init.xavier_normal_(conv.weight)
        return conv

    @staticmethod
    def normalize_between_0_and_1(x):
        return vit_ops.minmax_normalize(x)

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError(f"CNN_Decoder: input is on {x.device}; the B200 hot path has no CPU fallback")
        params = []
        for i in range(1, 5):
            m = getattr(self, f"deconv{i}")
            params.extend((m.weight, m.bias))
        need = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params))
        return _VitDecoderFn.apply(self, need, x, *params)


class VIT_encoder_CNN_decoder(nn.Module):
    """pytorch/VITs.py:197-229."""

    def __init__(self, config, image_size, number_of_output_channels):
        super().__init__()
        self.config = config
        self.model_type = config['model type']
        self.image_size = image_size
        self.number_of_output_channels = number_of_output_channels
        self.num_base_filters = config["number of base filters"]
        self.kernel_size = config["convolution kernel size"]
        self.optimizer = config["optimizer"]
        self.dropout = config["dropout ratio"]
        self.patch_size = config["patch size"]
        self.projection_dim = config["projection dim"]
        self.num_attention_heads = config["num heads"]
        self.num_transformer_layers = config["transformer layers"]
        self.dim_head = self.projection_dim if config["dim head"] else 64  # -1 is truthy -> projection_dim (:212)
        self.precision = config.get("precision", "bf16")
        self.vit_encoder = CustomViT(image_size=int(image_size[1]), patch_size=self.patch_size, dim=self.projection_dim,
                                     depth=self.num_transformer_layers, heads=self.num_attention_heads,
                                     mlp_dim=self.projection_dim * 4, dim_head=self.dim_head,
                                     num_image_channels=int(image_size[-1]) if len(image_size) > 2 else 4,
                                     precision=self.precision)
        self.cnn_decoder = CNN_Decoder(num_output_channels=self.number_of_output_channels,
                                       kernel_size=self.kernel_size, num_base_filters=self.num_base_filters,
                                       projection_dim=self.projection_dim, precision=self.precision)

    def set_precision(self, precision: str):
        self.precision = precision
        self.vit_encoder.set_precision(precision)
        self.cnn_decoder.set_precision(precision)
        return self

    def invalidate_packed_weights(self):
        self.vit_encoder.invalidate_packed_weights()
        self.cnn_decoder.invalidate_packed_weights()

    def repack_weights(self):
        self.vit_encoder.repack_weights()
        self.cnn_decoder.repack_weights()

    def set_grad_ready_hook(self, hook) -> None:
        self.__dict__["_grad_ready_hook"] = hook

    def forward(self, x):
        x = self.vit_encoder(x)
        x = self.cnn_decoder(x)
        return x

    @torch.no_grad()
    def train_step(self, x: torch.Tensor, target: Optional[torch.Tensor] = None, *,
                   points: Optional[torch.Tensor] = None, sigma: float = 3.0, accumulation_steps: int = 1,
                   accumulate: bool = False, loss_scale: float = 1.0) -> torch.Tensor:
        """forward + MSE + backward with gradients written straight into ``param.grad``
        (pytorch/train_pytorch.py:132-137).  The min/max normalisation sits between the last layer and
        the loss, so the loss gradient is taken in NCHW and routed through its backward."""
        if not x.is_cuda:
            raise RuntimeError("VIT_encoder_CNN_decoder.train_step: CPU tensor (there is no CPU fallback)")
        enc, dec = self.vit_encoder._engine(), self.cnn_decoder._engine()
        hook = self.__dict__.get("_grad_ready_hook")
        b = x.shape[0]
        tokens, s_enc = enc.forward(x.contiguous().float(), save=True)
        c_out, cpad = self.number_of_output_channels, dec.out_cpad()
        hw = int(self.image_size[0]), int(self.image_size[1])
        fused_tail = (os.environ.get("POSEB200_VIT_TAIL_UNFUSED", "0") != "1"
                      and ops.minmax_mse_eligible(c_out, hw[0], hw[1], dec.act_dtype, cpad))
        g_nchw = dc_y = None
        if fused_tail:
            # normalisation + loss + their backward + LeakyReLU'(deconv4) in three passes over the heatmaps
            out, s_dec = dec.forward(tokens, b, save=True, normalize=False)
            loss_sum, dc_y = ops.minmax_mse_fwd_bwd(out, target, points=points, sigma=sigma,
                                                    accumulation_steps=accumulation_steps, loss_scale=loss_scale,
                                                    cpad=cpad)
        else:
            out, s_dec = dec.forward(tokens, b, save=True)
            loss_sum, g_nchw, _ = ops.mse_loss_fwd_bwd(out, target, points=points, sigma=sigma,
                                                       accumulation_steps=accumulation_steps, loss_scale=loss_scale,
                                                       want_grad_nchw=True)
        beta = 1.0 if accumulate else 0.0

        def _grad(p):
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            return p.grad

        def dec_sink(name: str):
            m = getattr(self.cnn_decoder, name)
            return _grad(m.weight), _grad(m.bias), beta

        def dec_done(name: str):
            if hook is not None:
                hook(f"cnn_decoder.{name}.bias")
                hook(f"cnn_decoder.{name}.weight")
        dec_sink.done = dec_done

        def enc_sink(name: str, p):
            return _grad(p), beta

        def enc_done(name: str):
            if hook is not None:
                hook(f"vit_encoder.{name}")
        enc_sink.done = enc_done

        g_tok = dec.backward(s_dec, g_nchw, dec_sink, need_input_grad=True, dc=dc_y)
        enc.backward(s_enc, g_tok, enc_sink)
        return loss_sum / float(out.numel() * accumulation_steps)

    @torch.no_grad()
    def predict_peaks(self, x: torch.Tensor, soft: bool = False) -> torch.Tensor:
        out = self.forward(x)
        return ops.peaks_softargmax(out) if soft else ops.peaks_argmax(out)


# ====================================================================================================================
# four cameras: shared ViT encoder per view, four rounds of cross-attention against the joint encoding, shared decoder
# ====================================================================================================================
class CrossAttention(nn.Module):
    """pytorch/VITs.py:235-250 (parameter container; executes as part of VIT4CamerasBaseLine on the B200 path)."""

    def __init__(self, input_dim, output_dim):
        super().__init__()
        self.layers = nn.Sequential(Transformer(dim=input_dim, depth=1, heads=4, dim_head=output_dim,
                                                mlp_dim=output_dim, dropout=0.),
                                    nn.LayerNorm(input_dim), nn.Linear(input_dim, output_dim), nn.GELU())

    def forward(self, x):  # pragma: no cover - guard
        raise NotImplementedError("CrossAttention executes as part of VIT4CamerasBaseLine.forward on the B200 hot path")


class _Vit4Fn(torch.autograd.Function):
    """the whole model as one autograd node: forward and backward are engine schedules of C-ABI launches."""

    @staticmethod
    def forward(ctx, module, need, x, *params):
        out, saved = module._run_forward(x.contiguous().float(), save=need)
        ctx.module, ctx.saved = module, saved
        return out

    @staticmethod
    def backward(ctx, g):
        module = ctx.module
        store: Dict[str, torch.Tensor] = {}

        def sink_for(prefix):
            def sink(name, p):
                store[prefix + name] = torch.empty_like(p)
                return store[prefix + name], 0.0
            return sink

        module._run_backward(ctx.saved, sink_for, g_out=g.contiguous().float())
        ctx.saved = None
        return (None, None, None, *[store.get(n) for n, _ in module._live_params()])


class VIT4CamerasBaseLine(nn.Module):
    """pytorch/VITs.py:253-306 (model type ALL_CAMS_18_POINTS_VIT).  The four views ride through the shared encoder,
    the cross-attention rounds (weights shared by the views of a round, every view against the same ORIGINAL joint
    encoding) and the shared decoder as ONE 4B batch, view-major; the decoder's min/max normalisation is taken per
    view, as the reference's four separate decoder calls do."""

    def __init__(self, config, image_size, number_of_output_channels):
        super().__init__()
        self.config = config
        self.model_type = config['model type']
        self.image_size = image_size
        self.number_of_output_channels = number_of_output_channels
        self.num_base_filters = config["number of base filters"]
        self.kernel_size = config["convolution kernel size"]
        self.optimizer = config["optimizer"]
        self.dropout = config["dropout ratio"]
        self.patch_size = config["patch size"]
        self.projection_dim = config["projection dim"]
        self.num_attention_heads = config["num heads"]
        self.num_transformer_layers = config["transformer layers"]
        self.dim_head = self.projection_dim if config["dim head"] else 64
        self.num_cross_attention_layers = 4
        self.precision = config.get("precision", "bf16")
        self.shared_vit_encoder = CustomViT(image_size=int(image_size[1]), patch_size=self.patch_size,
                                            dim=self.projection_dim, depth=self.num_transformer_layers,
                                            heads=self.num_attention_heads, mlp_dim=self.projection_dim * 4,
                                            dim_head=self.dim_head, precision=self.precision)
        self.cross_attentions = nn.ModuleList(CrossAttention(input_dim=self.projection_dim * 5,
                                                             output_dim=self.projection_dim)
                                              for _ in range(self.num_cross_attention_layers))
        self.shared_cnn_decoder = CNN_Decoder(num_output_channels=self.number_of_output_channels // 4,
                                              kernel_size=self.kernel_size, num_base_filters=self.num_base_filters,
                                              projection_dim=self.projection_dim, precision=self.precision)

    # ---- engine plumbing ----------------------------------------------------------------------------------------
    def _ca_engines(self):
        from .vit_engine import CrossAttentionEngine
        engs = self.__dict__.get("_ca")
        if engs is None or engs[0].precision != self.precision:
            engs = [CrossAttentionEngine(m, self.precision) for m in self.cross_attentions]
            self.__dict__["_ca"] = engs
        return engs

    def _live_params(self):
        return [(n, p) for n, p in self.named_parameters() if not n.endswith("cls_token")]

    def set_precision(self, precision: str):
        self.precision = precision
        self.shared_vit_encoder.set_precision(precision)
        self.shared_cnn_decoder.set_precision(precision)
        return self

    def invalidate_packed_weights(self):
        self.shared_vit_encoder.invalidate_packed_weights()
        self.shared_cnn_decoder.invalidate_packed_weights()
        for e in self.__dict__.get("_ca") or []:
            e.invalidate()

    def repack_weights(self):
        self.shared_vit_encoder.repack_weights()
        self.shared_cnn_decoder.repack_weights()
        for e in self.__dict__.get("_ca") or []:
            if not e.repack_all():
                e.invalidate()

    def set_grad_ready_hook(self, hook) -> None:
        self.__dict__["_grad_ready_hook"] = hook

    # ---- view <-> batch re-arrangements: strided column-block moves, no ATen copies -------------------------------
    @staticmethod
    def _views_to_batch(t: torch.Tensor) -> torch.Tensor:
        """[B, 4*c, ...] (torch.split(x, c, dim=1) order) -> [4B, c, ...] view-major."""
        b, c4 = t.shape[0], t.shape[1]
        inner = t[0, 0].numel() * (c4 // 4)
        out = torch.empty((4 * b, c4 // 4) + tuple(t.shape[2:]), device=t.device, dtype=t.dtype)
        for v in range(4):
            vit_ops.colblock(t, out, rows=b, ncols=inner, src_row_stride=4 * inner, dst_row_stride=inner,
                             src_col0=v * inner, dst_col0=v * b * inner)
        return out

    @staticmethod
    def _batch_to_views(t: torch.Tensor) -> torch.Tensor:
        """[4B, c, ...] view-major -> [B, 4*c, ...] (torch.cat of the four views along dim 1)."""
        b, c = t.shape[0] // 4, t.shape[1]
        inner = t[0].numel()
        out = torch.empty((b, 4 * c) + tuple(t.shape[2:]), device=t.device, dtype=t.dtype)
        for v in range(4):
            vit_ops.colblock(t, out, rows=b, ncols=inner, src_row_stride=inner, dst_row_stride=4 * inner,
                             src_col0=v * b * inner, dst_col0=v * inner)
        return out

    # ---- engine schedules ------------------------------------------------------------------------------------------
    def _run_encoder_and_rounds(self, x: torch.Tensor, save: bool):
        enc_eng, ca = self.shared_vit_encoder._engine(), self._ca_engines()
        if x.shape[1] != 16:
            raise ValueError("VIT4CamerasBaseLine: expected four 4-channel views (16 input channels)")
        b = x.shape[0]
        skip, s_enc = enc_eng.forward(self._views_to_batch(x), save)                 # [4B*S, dim], view-major
        dim = skip.shape[1]
        rows = skip.shape[0] // 4                                                   # B*S tokens per view
        s_tok = rows // b
        joint = torch.empty((rows, 4 * dim), device=x.device, dtype=skip.dtype)     # torch.cat(encodings, dim=-1) :295
        for v in range(4):
            vit_ops.colblock(skip, joint, rows=rows, ncols=dim, src_row_stride=dim, dst_row_stride=4 * dim,
                             src_col0=v * rows * dim, dst_col0=v * dim)
        enc, s_rounds = skip, []
        for eng in ca:
            xin = torch.empty((4 * rows, 5 * dim), device=x.device, dtype=skip.dtype)   # cat([enc_v, encodings]) :297
            vit_ops.colblock(enc, xin, rows=4 * rows, ncols=dim, src_row_stride=dim, dst_row_stride=5 * dim)
            vit_ops.colblock(joint, xin, rows=4 * rows, ncols=4 * dim, src_row_stride=4 * dim, dst_row_stride=5 * dim,
                             dst_col0=dim, src_rows_mod=rows)
            enc, s_i = eng.forward(xin, enc, 4 * b, s_tok, save)                        # ... + enc_v
            s_rounds.append(s_i)
        dec_in = ops.add(enc, skip)                                                     # enc_v + skip_v :301
        return dec_in, {"b": b, "rows": rows, "s": s_tok, "dim": dim, "enc": s_enc, "rounds": s_rounds}

    def _run_forward(self, x: torch.Tensor, save: bool):
        dec_in, saved = self._run_encoder_and_rounds(x, save)
        out4, s_dec = self.shared_cnn_decoder._engine().forward(dec_in, 4 * saved["b"], save, groups=4)
        saved["dec"] = s_dec
        return self._batch_to_views(out4), (saved if save else None)

    def _run_backward(self, saved: dict, sink_for, g_out: Optional[torch.Tensor] = None,
                      dc: Optional[torch.Tensor] = None) -> None:
        """g_out: gradient w.r.t. the [B, C, H, W] output, or dc: gradient w.r.t. deconv4's pre-activation (view-major,
        from the fused loss tail).  sink_for(prefix) -> ParamSink of that sub-module."""
        enc_eng, dec_eng, ca = self.shared_vit_encoder._engine(), self.shared_cnn_decoder._engine(), self._ca_engines()
        b, rows, s_tok, dim = saved["b"], saved["rows"], saved["s"], saved["dim"]
        dsink_raw = sink_for("shared_cnn_decoder.")

        def dec_sink(name: str):
            m = getattr(self.shared_cnn_decoder, name)
            (dw, beta), (db, _) = dsink_raw(name + ".weight", m.weight), dsink_raw(name + ".bias", m.bias)
            return dw, db, beta
        raw_done = getattr(dsink_raw, "done", None)
        if raw_done is not None:
            def dec_done(name: str):
                raw_done(name + ".bias")
                raw_done(name + ".weight")
            dec_sink.done = dec_done
        g4 = self._views_to_batch(g_out) if g_out is not None else None
        g_decin = dec_eng.backward(saved["dec"], g4, dec_sink, need_input_grad=True, dc=dc)     # [4*rows, dim]
        g_enc = g_decin.clone()
        g_joint = torch.empty((rows, 4 * dim), device=g_decin.device, dtype=g_decin.dtype)
        for i in range(len(ca) - 1, -1, -1):
            g_xin = ca[i].backward(saved["rounds"][i], g_enc, 4 * b, s_tok, sink_for(f"cross_attentions.{i}."))
            saved["rounds"][i] = None
            # enc_next = block(cat[enc, joint]) + enc: the block's share of d/d(enc) joins the residual's
            vit_ops.colblock(g_xin, g_enc, rows=4 * rows, ncols=dim, src_row_stride=5 * dim, dst_row_stride=dim,
                             accumulate=True)
            # the joint encoding fed all four views of every round
            vit_ops.colblock(g_xin, g_joint, rows=rows, ncols=4 * dim, src_row_stride=5 * dim, dst_row_stride=4 * dim,
                             src_col0=dim, nfold=4, fold_stride=rows * 5 * dim, accumulate=(i < len(ca) - 1))
        g_skip = ops.add(g_decin, g_enc)
        for v in range(4):
            vit_ops.colblock(g_joint, g_skip, rows=rows, ncols=dim, src_row_stride=4 * dim, dst_row_stride=dim,
                             src_col0=v * dim, dst_col0=v * rows * dim, accumulate=True)
        enc_eng.backward(saved["enc"], g_skip, sink_for("shared_vit_encoder."))

    # ---- nn.Module surface --------------------------------------------------------------------------------------------
    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError(f"VIT4CamerasBaseLine: input is on {x.device}; the B200 hot path has no CPU fallback")
        params = [p for _, p in self._live_params()]
        need = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        return _Vit4Fn.apply(self, need, x, *params)

    @torch.no_grad()
    def train_step(self, x: torch.Tensor, target: Optional[torch.Tensor] = None, *,
                   points: Optional[torch.Tensor] = None, sigma: float = 3.0, accumulation_steps: int = 1,
                   accumulate: bool = False, loss_scale: float = 1.0) -> torch.Tensor:
        """forward + MSE + backward with gradients written straight into ``param.grad`` (flat buckets)."""
        if not x.is_cuda:
            raise RuntimeError("VIT4CamerasBaseLine.train_step: CPU tensor (there is no CPU fallback)")
        hook = self.__dict__.get("_grad_ready_hook")
        dec = self.shared_cnn_decoder._engine()
        b = x.shape[0]
        dec_in, saved = self._run_encoder_and_rounds(x.contiguous().float(), save=True)
        c4, cpad = self.number_of_output_channels // 4, dec.out_cpad()
        hw = int(self.image_size[0]), int(self.image_size[1])
        numel = b * self.number_of_output_channels * hw[0] * hw[1]
        tgt4 = self._views_to_batch(target.contiguous().float()) if target is not None else None
        pts4 = self._views_to_batch(points.contiguous().float()) if points is not None and target is None else None
        loss_sum = torch.zeros(1, device=x.device, dtype=torch.float32)
        g_out = dc = None
        if ops.minmax_mse_eligible(c4, hw[0], hw[1], dec.act_dtype, cpad):
            out4, saved["dec"] = dec.forward(dec_in, 4 * b, True, normalize=False)
            dc = torch.empty((4 * b, hw[0], hw[1], cpad), device=x.device, dtype=torch.bfloat16)
            for v in range(4):      # per-view min/max (the reference's four decoder calls), one shared mean
                sl = slice(v * b, (v + 1) * b)
                ops.minmax_mse_fwd_bwd(out4[sl], tgt4[sl] if tgt4 is not None else None,
                                       points=pts4[sl] if pts4 is not None else None, sigma=sigma,
                                       accumulation_steps=accumulation_steps, loss_scale=loss_scale, cpad=cpad,
                                       numel=numel, loss_sum=loss_sum, grad_out=dc[sl])
        else:
            out4, saved["dec"] = dec.forward(dec_in, 4 * b, True, groups=4)
            ls, g4, _ = ops.mse_loss_fwd_bwd(out4, tgt4, points=pts4, sigma=sigma,
                                             accumulation_steps=accumulation_steps, loss_scale=loss_scale,
                                             want_grad_nchw=True)
            loss_sum, g_out = ls, self._batch_to_views(g4)
        beta = 1.0 if accumulate else 0.0

        def sink_for(prefix):
            def sink(name, p):
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
                return p.grad, beta

            def done(name):
                if hook is not None:
                    hook(prefix + name)
            sink.done = done
            return sink
        self._run_backward(saved, sink_for, g_out=g_out, dc=dc)
        return loss_sum / float(numel * accumulation_steps)

    @torch.no_grad()
    def predict_peaks(self, x: torch.Tensor, soft: bool = False) -> torch.Tensor:
        out, _ = self._run_forward(x.contiguous().float(), save=False)
        return ops.peaks_softargmax(out) if soft else ops.peaks_argmax(out)
