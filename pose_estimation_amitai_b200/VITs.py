"""Drop-in ViT-encoder / conv-transpose-decoder heatmap model (pytorch/VITs.py:13-58,197-229).

Same constructors, attribute names and state_dict keys as the reference (``vit_encoder.*``,
``cnn_decoder.deconv{1..4}.*``; 104 entries).  The dead classes of the reference file
(PositionalEncoding, TransformerBlock, ViTEncoder, TransformerDecoder, VIT_encoder_decoder -- not
reachable from Network.py) and the 4-camera model are outside the hot path.
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
