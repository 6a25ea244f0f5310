"""Drop-in replacements for the reference's conv heatmap modules (pytorch/CNNs.py).

Same constructors, attribute / sub-module names and state_dict keys as the reference
(`Encoder2DAtrous` CNNs.py:9-88, `Decoder2d` :92-157, `BasicNet` :160-186) so
``load_state_dict(strict=True)`` of a reference checkpoint succeeds and the same seed yields the
same random init; the forward / backward run on hand-written sm_100a kernels through the C ABI
(include/poseb200.h).  nn.Conv2d / nn.ConvTranspose2d / nn.BatchNorm2d sub-modules are parameter
containers only -- their ATen forwards are never called.  BatchNorm is constructed but inert and
Dropout is p=0 / never applied, exactly as in the reference (CNNs.py:56-71,143-149 commented out).

Optional config keys (absent => defaults): "precision": "bf16" (default) | "fp16" | "fp32".
"fp16" = IEEE-half forward operands (the dtype of the reference's own autocast region, train_pytorch.py:133)
against bf16 gradients: same tensor-core rate as bf16, 8x smaller forward rounding error, no loss scaling.
CPU tensors raise: there is no CPU fallback.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import nn

from . import ops
from .engine import PRECISIONS, DecoderEngine, DisentangleMidEngine, EncoderEngine, PointwiseEngine


def _require_cuda(x: torch.Tensor, who: str) -> None:
    if not x.is_cuda:
        raise RuntimeError(f"{who}: input is on {x.device}; the B200 hot path has no CPU fallback "
                           "(move the model and inputs to cuda)")


def _boundary(t: torch.Tensor) -> torch.Tensor:
    """A tensor handed to autograd at a module boundary.  autograd casts the incoming gradient to the OUTPUT's dtype,
    and heatmap-loss gradients (~1e-9) flush to zero in IEEE half, so fp16 activations cross the boundary widened to
    fp32 (exact; the consumer narrows them back without loss) and their gradients arrive unharmed."""
    return t.float() if t.dtype == torch.float16 else t


def _fresh_sink(module: nn.Module, store: Dict[str, Tuple[torch.Tensor, torch.Tensor]]):
    """gradient sink that allocates new tensors (autograd path: returned to autograd)."""
    def sink(name: str):
        m = getattr(module, name)
        dw, db = torch.empty_like(m.weight), torch.empty_like(m.bias)
        store[name] = (dw, db)
        return dw, db, 0.0
    return sink


def _param_sink(module: nn.Module, accumulate: bool, prefix: str = "", hook=None):
    """gradient sink that writes straight into ``param.grad`` (fused train-step path; with flat
    gradient buckets these are views into NCCL-reduced bucket memory).  `hook(full_name)` is called
    once the kernel producing that gradient has been enqueued (bucketed all-reduce trigger)."""
    def sink(name: str):
        m = getattr(module, name)
        beta = 1.0 if accumulate else 0.0
        for p in (m.weight, m.bias):
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        return m.weight.grad, m.bias.grad, beta

    def done(name: str):
        if hook is not None:
            hook(f"{prefix}{name}.bias")
            hook(f"{prefix}{name}.weight")
    sink.done = done
    return sink


class _EncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, need, x, *params):
        y, saved = module._engine().forward(x.contiguous().float(), save=need)
        ctx.module, ctx.saved = module, saved
        return _boundary(y.permute(0, 3, 1, 2))  # logical NCHW, physical NHWC (channels_last)

    @staticmethod
    def backward(ctx, g):
        module = ctx.module
        eng = module._engine()
        g_nhwc = g.permute(0, 2, 3, 1).contiguous().to(eng.grad_dtype)
        store: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}
        eng.backward(ctx.saved, g_nhwc, _fresh_sink(module, store))
        ctx.saved = None
        grads = []
        for name in module._layer_names:
            grads.extend(store[name])
        return (None, None, None, *grads)


class _DecoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, need, x, *params):
        eng = module._engine()
        x_nhwc = x.permute(0, 2, 3, 1).contiguous().to(eng.act_dtype)
        y, saved = eng.forward(x_nhwc, save=need)
        ctx.module, ctx.saved, ctx.need_x = module, saved, x.requires_grad
        return y

    @staticmethod
    def backward(ctx, g):
        module = ctx.module
        eng = module._engine()
        dc_y = ops.grad_ingest(g, ctx.saved["out"], eng.grad_dtype, cpad=eng.out_cpad())
        store: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}
        g_in = eng.backward(ctx.saved, dc_y, _fresh_sink(module, store), need_input_grad=ctx.need_x)
        ctx.saved = None
        grads = []
        for name in module._layer_names:
            grads.extend(store[name])
        gx = g_in.permute(0, 3, 1, 2) if g_in is not None else None
        return (None, None, gx, *grads)


class _EngineMixin:
    _engine_cls = None
    _layer_names: Tuple[str, ...] = ()

    def _engine(self):
        eng = self.__dict__.get("_eng")
        if eng is None or eng.precision != self.precision:
            eng = self._engine_cls(self, self.precision)
            self.__dict__["_eng"] = eng
        return eng

    def _params(self):
        out = []
        for name in self._layer_names:
            m = getattr(self, name)
            out.extend((m.weight, m.bias))
        return out

    def set_precision(self, precision: str):
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {PRECISIONS}")
        self.precision = precision
        return self

    def invalidate_packed_weights(self):
        eng = self.__dict__.get("_eng")
        if eng is not None:
            eng.invalidate()

    def repack_weights(self):
        """weights changed in place (fused Adam): refresh the packed operands with one launch."""
        eng = self.__dict__.get("_eng")
        if eng is not None and not (hasattr(eng, "repack_all") and eng.repack_all()):
            eng.invalidate()


class Encoder2DAtrous(_EngineMixin, nn.Module):
    """pytorch/CNNs.py:9-88."""
    _engine_cls = EncoderEngine
    _layer_names = tuple(f"conv{i}" for i in range(1, 10))

    def __init__(self, img_size, filters, kernel_size, dilation_rate, dropout, precision: str = "bf16"):
        super().__init__()
        self.image_size = img_size
        self.dilation_rate = int(dilation_rate)
        self.filters = int(filters)
        self.kernel_size = int(kernel_size)
        self.output_ratio = 4
        self.padding = 2
        self.precision = precision
        self.maxpool = nn.MaxPool2d(kernel_size=2, stride=2)
        self.leakyrelu = nn.LeakyReLU(0.1)
        # reference: self.dropout = int(dropout) -> nn.Dropout(float(int(0.5))) == Dropout(p=0) (CNNs.py:14,22)
        self.dropout = nn.Dropout(float(int(dropout)))
        f = self.filters
        widths = [(int(img_size[-1]), f), (f, f), (f, f), (f, 2 * f), (2 * f, 2 * f), (2 * f, 2 * f),
                  (2 * f, 4 * f), (4 * f, 4 * f), (4 * f, 4 * f)]
        for i, (ci, co) in enumerate(widths, 1):  # conv_i then bn_i: the reference's creation order
            setattr(self, f"conv{i}", self.get_conv2d(input_channels=ci, num_filters=co))
            setattr(self, f"bn{i}", nn.BatchNorm2d(co))

    def get_conv2d(self, num_filters, input_channels):
        return nn.Conv2d(in_channels=int(input_channels), out_channels=int(num_filters),
                         kernel_size=int(self.kernel_size), padding=int(self.padding),
                         dilation=int(self.dilation_rate))

    def get_output_size(self):
        return (self.image_size[0] // self.output_ratio, self.image_size[1] // self.output_ratio,
                self.filters * self.output_ratio)

    def forward(self, x):
        _require_cuda(x, "Encoder2DAtrous")
        params = self._params()
        need = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        return _EncoderFn.apply(self, need, x, *params)


class Decoder2d(_EngineMixin, nn.Module):
    """pytorch/CNNs.py:92-157."""
    _engine_cls = DecoderEngine
    _layer_names = tuple(f"conv2dTranspose{i}" for i in range(1, 5))

    def __init__(self, input_shape, num_output_channels, kernel_size, filters, dropout, precision: str = "bf16"):
        super().__init__()
        self.input_shape = input_shape
        self.num_output_channels = int(num_output_channels)
        self.filters = int(filters)
        self.kernel_size = int(kernel_size)
        self.precision = precision
        self.maxpool = nn.MaxPool2d(kernel_size=2, stride=2)
        self.leakyrelu = nn.LeakyReLU(0.1)
        self.dropout = nn.Dropout(float(dropout))  # built, never applied (CNNs.py:106,151-157)
        c = int(input_shape[-1])
        plan = [(c, c // 2, 2), (c // 2, c // 2, 1), (c // 2, c // 2, 1), (c // 2, self.num_output_channels, 2)]
        for i, (ci, co, stride) in enumerate(plan, 1):
            setattr(self, f"conv2dTranspose{i}",
                    nn.ConvTranspose2d(in_channels=ci, out_channels=co, kernel_size=self.kernel_size, stride=stride,
                                       padding=1, output_padding=1 if stride == 2 else 0))
            setattr(self, f"bn{i}", nn.BatchNorm2d(co))

    @staticmethod
    def normalize_between_0_and_1(x):
        from . import vit_ops
        return vit_ops.minmax_normalize(x)

    def get_conv2d_transpose(self, in_channels, out_channels, stride):
        return nn.ConvTranspose2d(in_channels=int(in_channels), out_channels=int(out_channels),
                                  kernel_size=int(self.kernel_size), stride=int(stride), padding=1,
                                  output_padding=1)

    def forward(self, x):
        _require_cuda(x, "Decoder2d")
        params = self._params()
        need = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params))
        return _DecoderFn.apply(self, need, x, *params)


class BasicNet(nn.Module):
    """pytorch/CNNs.py:160-186: decoder(encoder(x)); (B,Cin,H,W) fp32 -> (B,C,H,W) fp32."""

    def __init__(self, config, image_size, number_of_output_channels):
        super().__init__()
        self.config = config
        self.model_type = config['model type']
        self.image_size = image_size
        self.number_of_output_channels = number_of_output_channels
        self.num_base_filters = config["number of base filters"]
        self.kernel_size = config["convolution kernel size"]
        self.dilation_rate = config["dilation rate"]
        self.dropout = config["dropout ratio"]
        self.precision = config.get("precision", "bf16")
        self.encoder = Encoder2DAtrous(img_size=self.image_size, filters=self.num_base_filters,
                                       kernel_size=self.kernel_size, dilation_rate=self.dilation_rate,
                                       dropout=self.dropout, precision=self.precision)
        self.decoder = Decoder2d(input_shape=self.encoder.get_output_size(), filters=self.num_base_filters,
                                 kernel_size=self.kernel_size, dropout=self.dropout,
                                 num_output_channels=self.number_of_output_channels, precision=self.precision)

    def set_precision(self, precision: str):
        self.precision = precision
        self.encoder.set_precision(precision)
        self.decoder.set_precision(precision)
        return self

    def invalidate_packed_weights(self):
        self.encoder.invalidate_packed_weights()
        self.decoder.invalidate_packed_weights()

    def repack_weights(self):
        self.encoder.repack_weights()
        self.decoder.repack_weights()

    def forward(self, x):
        x = self.encoder(x)
        x = self.decoder(x)
        return x

    # ---- fused paths (what the Trainer / bench use; same math, fewer passes over HBM) ----------
    @torch.no_grad()
    def train_step(self, x: torch.Tensor, target: Optional[torch.Tensor] = None, *,
                   points: Optional[torch.Tensor] = None, sigma: float = 3.0, accumulation_steps: int = 1,
                   accumulate: bool = False, loss_scale: float = 1.0) -> torch.Tensor:
        """forward + MSE (+ fused Gaussian target when `points` is given) + backward, gradients
        written straight into ``param.grad`` (pytorch/train_pytorch.py:132-137 in one call).
        Returns the mean loss / accumulation_steps as a 1-element CUDA tensor (no host sync)."""
        _require_cuda(x, "BasicNet.train_step")
        enc, dec = self.encoder._engine(), self.decoder._engine()
        feat, s_enc = enc.forward(x.contiguous().float(), save=True)
        numel = x.shape[0] * self.number_of_output_channels * x.shape[2] * x.shape[3]
        if dec.head_fusable():
            # the head's epilogue computes the loss and dC itself: the fp32 heatmaps never reach HBM
            tgt = target.contiguous().float() if target is not None else None
            pts = points.contiguous().float() if points is not None and target is None else None
            loss_sum, dc_y, s_dec = dec.forward_loss(feat, s_enc["out_w"], target=tgt, points=pts, sigma=sigma,
                                                     accumulation_steps=accumulation_steps, loss_scale=loss_scale)
        else:
            out, s_dec = dec.forward(feat, save=True, x_w=s_enc["out_w"])
            loss_sum, _, dc_y = ops.mse_loss_fwd_bwd(out, target, points=points, sigma=sigma,
                                                     accumulation_steps=accumulation_steps, loss_scale=loss_scale,
                                                     grad_nhwc_dtype=dec.grad_dtype, cpad=dec.out_cpad())
        hook = self.__dict__.get("_grad_ready_hook")
        # the decoder's first-layer input gradient also applies LeakyReLU'(conv9) in its epilogue
        g_feat, dc_feat = dec.backward(s_dec, dc_y, _param_sink(self.decoder, accumulate, "decoder.", hook),
                                       need_input_grad=True, mask_below=s_enc["conv9"][1])
        enc.backward(s_enc, g_feat, _param_sink(self.encoder, accumulate, "encoder.", hook), dc_out=dc_feat)
        return loss_sum / float(numel * accumulation_steps)

    def set_grad_ready_hook(self, hook) -> None:
        """`hook(param_name)` fires as each gradient kernel is enqueued (parallel.FlatBuckets.grad_ready)."""
        self.__dict__["_grad_ready_hook"] = hook

    @torch.no_grad()
    def predict_peaks(self, x: torch.Tensor, soft: bool = False) -> torch.Tensor:
        """frame-sharded inference step: forward + per-joint peaks, (B,C,2) [x,y] float32 on device
        (pytorch/train_pytorch.py:155-170,199-213 without the heatmap D2H)."""
        _require_cuda(x, "BasicNet.predict_peaks")
        enc, dec = self.encoder._engine(), self.decoder._engine()
        chunk = int(self.config.get("inference chunk", 512)) if hasattr(self.config, "get") else 512
        if x.shape[0] > chunk:   # bound the activation working set (a 4096-frame batch would hold ~70 GB at once)
            return torch.cat([self.predict_peaks(x[i:i + chunk], soft) for i in range(0, x.shape[0], chunk)])
        feat, _ = enc.forward(x.contiguous().float(), save=False)
        if not soft and dec.head_fusable():
            return dec.forward_peaks(feat)       # arg-max inside the head's epilogue: no heatmap tensor at all
        out, _ = dec.forward(feat, save=False)
        return ops.peaks_softargmax(out) if soft else ops.peaks_argmax(out)



class _ViewsToBatchFn(torch.autograd.Function):
    """differentiable wrappers: the two re-arrangements are permutations and each other's inverse, so the backward
    of one is the other."""

    @staticmethod
    def forward(ctx, t):
        return _views_to_batch_raw(t)

    @staticmethod
    def backward(ctx, g):
        return _batch_to_views_raw(g)


class _BatchToViewsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        return _batch_to_views_raw(t)

    @staticmethod
    def backward(ctx, g):
        return _views_to_batch_raw(g)


def views_to_batch(t: torch.Tensor) -> torch.Tensor:
    return _ViewsToBatchFn.apply(t) if t.requires_grad and torch.is_grad_enabled() else _views_to_batch_raw(t)


def batch_to_views(t: torch.Tensor) -> torch.Tensor:
    return _BatchToViewsFn.apply(t) if t.requires_grad and torch.is_grad_enabled() else _batch_to_views_raw(t)


def _views_to_batch_raw(t: torch.Tensor) -> torch.Tensor:
    """[B, 4*c, ...] (views along channels, torch.split(x, c, dim=1) order) -> [4B, c, ...] view-major; four strided
    column-block moves (pb_colblock) instead of an ATen permute + copy."""
    from . import vit_ops
    if not t.is_cuda:      # host-side shape checks / tests
        b, c4 = t.shape[0], t.shape[1]
        return t.reshape(b, 4, c4 // 4, *t.shape[2:]).transpose(0, 1).reshape(4 * b, c4 // 4, *t.shape[2:])
    t = t.contiguous()
    b, c4 = t.shape[0], t.shape[1]
    inner = t[0, 0].numel() * (c4 // 4)
    out = torch.empty((4 * b, c4 // 4) + tuple(t.shape[2:]), device=t.device, dtype=t.dtype)
    for v in range(4):
        vit_ops.colblock(t, out, rows=b, ncols=inner, src_row_stride=4 * inner, dst_row_stride=inner,
                         src_col0=v * inner, dst_col0=v * b * inner)
    return out


def _batch_to_views_raw(t: torch.Tensor) -> torch.Tensor:
    """[4B, c, ...] view-major -> [B, 4*c, ...] (torch.cat(..., dim=1) of the four views)."""
    from . import vit_ops
    if not t.is_cuda:
        b = t.shape[0] // 4
        return t.reshape(4, b, *t.shape[1:]).transpose(0, 1).reshape(b, 4 * t.shape[1], *t.shape[2:])
    t = t.contiguous()
    b, c = t.shape[0] // 4, t.shape[1]
    inner = t[0].numel()
    out = torch.empty((b, 4 * c) + tuple(t.shape[2:]), device=t.device, dtype=t.dtype)
    for v in range(4):
        vit_ops.colblock(t, out, rows=b, ncols=inner, src_row_stride=inner, dst_row_stride=4 * inner,
                         src_col0=v * b * inner, dst_col0=v * inner)
    return out


class _PointwiseResidualFn(torch.autograd.Function):
    """y = conv1x1(x) + x on logical-NCHW / physical-NHWC tensors (autograd path of FourCamerasBaseLine)."""

    @staticmethod
    def forward(ctx, owner, x, weight, bias):
        eng = owner._pointwise_engine()
        x_nhwc = x.permute(0, 2, 3, 1).contiguous().to(eng.act_dtype)
        y = eng.forward(x_nhwc, residual=True)
        ctx.owner, ctx.need_x = owner, x.requires_grad
        ctx.save_for_backward(x_nhwc)
        return _boundary(y.permute(0, 3, 1, 2))

    @staticmethod
    def backward(ctx, g):
        eng = ctx.owner._pointwise_engine()
        (x_nhwc,) = ctx.saved_tensors
        g_nhwc = g.permute(0, 2, 3, 1).contiguous().to(eng.grad_dtype)
        conv = ctx.owner.shared_conv2d
        dw, db = torch.empty_like(conv.weight), torch.empty_like(conv.bias)
        gx = eng.backward(x_nhwc, g_nhwc, lambda name: (dw, db, 0.0), residual=True, need_input_grad=ctx.need_x)
        return None, (gx.permute(0, 3, 1, 2) if gx is not None else None), dw, db


class FourCamerasBaseLine(nn.Module):
    """pytorch/CNNs.py:189-237 (model type ALL_CAMS_18_POINTS): four 4-channel views through ONE shared encoder,
    their encodings mixed by a 1x1 conv (+ residual), ONE shared decoder on cat(view encoding, mixed encodings),
    outputs concatenated along channels.  The four views ride through the shared stacks as one 4B batch (same
    weights, per-sample-independent layers), so each stack runs once per step and its weight gradients are summed
    over the views by the kernels themselves."""

    def __init__(self, config, image_size, number_of_output_channels):
        super().__init__()
        self.config = config
        self.model_type = config['model type']
        self.image_size = image_size
        self.number_of_output_channels = number_of_output_channels
        self.num_base_filters = config["number of base filters"]
        self.kernel_size = config["convolution kernel size"]
        self.dilation_rate = config["dilation rate"]
        self.dropout = config["dropout ratio"]
        self.precision = config.get("precision", "bf16")
        self.shared_encoder = Encoder2DAtrous(img_size=(image_size[0], image_size[1], image_size[2] // 4),
                                              filters=self.num_base_filters, kernel_size=self.kernel_size,
                                              dilation_rate=self.dilation_rate, dropout=self.dropout,
                                              precision=self.precision)
        width = self.shared_encoder.get_output_size()[-1] * 4
        self.shared_conv2d = nn.Conv2d(width, width, kernel_size=1, padding=0, bias=True)
        input_size = list(self.shared_encoder.get_output_size())
        input_size[-1] *= 5
        self.shared_decoder = Decoder2d(input_shape=input_size,
                                        num_output_channels=self.number_of_output_channels // 4,
                                        kernel_size=self.kernel_size, filters=self.num_base_filters,
                                        dropout=self.dropout, precision=self.precision)

    # ---- engine plumbing ----------------------------------------------------------------------
    def _pointwise_engine(self) -> PointwiseEngine:
        eng = self.__dict__.get("_pw")
        if eng is None or eng.precision != self.precision:
            eng = PointwiseEngine(self, "shared_conv2d", self.precision)
            self.__dict__["_pw"] = eng
        return eng

    def set_precision(self, precision: str):
        self.precision = precision
        self.shared_encoder.set_precision(precision)
        self.shared_decoder.set_precision(precision)
        return self

    def invalidate_packed_weights(self):
        self.shared_encoder.invalidate_packed_weights()
        self.shared_decoder.invalidate_packed_weights()
        if "_pw" in self.__dict__:
            self.__dict__["_pw"].invalidate()

    def repack_weights(self):
        self.shared_encoder.repack_weights()
        self.shared_decoder.repack_weights()
        pw = self.__dict__.get("_pw")
        if pw is not None and not pw.repack_all():
            pw.invalidate()

    def set_grad_ready_hook(self, hook) -> None:
        self.__dict__["_grad_ready_hook"] = hook

    # ---- view <-> batch re-arrangements ---------------------------------------------------------
    _views_to_batch = staticmethod(views_to_batch)
    _batch_to_views = staticmethod(batch_to_views)

    def forward(self, x):
        _require_cuda(x, "FourCamerasBaseLine")
        if x.shape[1] != 16:
            raise ValueError("FourCamerasBaseLine: expected four 4-channel views (16 input channels)")
        enc = self.shared_encoder(self._views_to_batch(x))                    # [4B, 256, h, w]
        all_encoders = self._batch_to_views(enc)                              # [B, 1024, h, w]
        all_encoders = _PointwiseResidualFn.apply(self, all_encoders, self.shared_conv2d.weight,
                                                  self.shared_conv2d.bias)
        dec_in = torch.cat((enc, all_encoders.repeat(4, 1, 1, 1)), dim=1)     # [4B, 1280, h, w]
        return self._batch_to_views(self.shared_decoder(dec_in))              # [B, C, H, W]

    @torch.no_grad()
    def train_step(self, x: torch.Tensor, target: Optional[torch.Tensor] = None, *,
                   points: Optional[torch.Tensor] = None, sigma: float = 3.0, accumulation_steps: int = 1,
                   accumulate: bool = False, loss_scale: float = 1.0) -> torch.Tensor:
        """forward + MSE + backward with gradients written into ``param.grad`` (flat buckets), engine level --
        the fused-step counterpart of BasicNet.train_step for the four-camera model."""
        _require_cuda(x, "FourCamerasBaseLine.train_step")
        enc, dec, pw = self.shared_encoder._engine(), self.shared_decoder._engine(), self._pointwise_engine()
        b = x.shape[0]
        from . import vit_ops
        feat, s_enc = enc.forward(self._views_to_batch(x.float()), save=True)                  # [4B, h, w, 256]
        h, w, c = feat.shape[1], feat.shape[2], feat.shape[3]
        rows = b * h * w                                                                        # pixels of one view
        # the glue of CNNs.py:226-236 (cat of the four encodings, cat of a view's encoding with the mixed ones) as
        # strided column-block moves: no ATen permute / cat / expand copies on the step
        all_in = torch.empty((b, h, w, 4 * c), device=feat.device, dtype=feat.dtype)
        for v in range(4):
            vit_ops.colblock(feat, all_in, rows=rows, ncols=c, src_row_stride=c, dst_row_stride=4 * c,
                             src_col0=v * rows * c, dst_col0=v * c)
        all_enc = pw.forward(all_in, residual=True)
        dec_in = torch.empty((4 * b, h, w, 5 * c), device=feat.device, dtype=feat.dtype)
        vit_ops.colblock(feat, dec_in, rows=4 * rows, ncols=c, src_row_stride=c, dst_row_stride=5 * c)
        vit_ops.colblock(all_enc, dec_in, rows=4 * rows, ncols=4 * c, src_row_stride=4 * c, dst_row_stride=5 * c,
                         dst_col0=c, src_rows_mod=rows)
        tgt4 = self._views_to_batch(target).contiguous() if target is not None else None
        pts4 = self._views_to_batch(points).contiguous() if points is not None and target is None else None
        numel = x.shape[0] * self.number_of_output_channels * x.shape[2] * x.shape[3]
        if dec.head_fusable():
            loss_sum, dc_y, s_dec = dec.forward_loss(dec_in, None, target=tgt4.float() if tgt4 is not None else None,
                                                     points=pts4.float() if pts4 is not None else None, sigma=sigma,
                                                     accumulation_steps=accumulation_steps, loss_scale=loss_scale)
        else:
            out4, s_dec = dec.forward(dec_in, save=True)                                      # [4B, C/4, H, W]
            loss_sum, _, dc_y = ops.mse_loss_fwd_bwd(out4, tgt4, points=pts4, sigma=sigma,
                                                     accumulation_steps=accumulation_steps, loss_scale=loss_scale,
                                                     grad_nhwc_dtype=dec.grad_dtype, cpad=dec.out_cpad())
        hook = self.__dict__.get("_grad_ready_hook")
        g_dec_in = dec.backward(s_dec, dc_y, _param_sink(self.shared_decoder, accumulate, "shared_decoder.", hook),
                                need_input_grad=True)
        # the mixed encodings fed all four views: their gradient is the sum of the four views' shares (fp32 sums)
        g_all = torch.empty((b, h, w, 4 * c), device=g_dec_in.device, dtype=g_dec_in.dtype)
        vit_ops.colblock(g_dec_in, g_all, rows=rows, ncols=4 * c, src_row_stride=5 * c, dst_row_stride=4 * c, src_col0=c,
                         nfold=4, fold_stride=rows * 5 * c)

        def pw_sink(name: str):
            conv = self.shared_conv2d
            for p in (conv.weight, conv.bias):
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
            return conv.weight.grad, conv.bias.grad, 1.0 if accumulate else 0.0

        def pw_done(name: str):
            if hook is not None:
                hook("shared_conv2d.bias")
                hook("shared_conv2d.weight")
        pw_sink.done = pw_done
        g_all_in = pw.backward(all_in, g_all, pw_sink, residual=True)
        g_feat = torch.empty((4 * b, h, w, c), device=g_dec_in.device, dtype=g_dec_in.dtype)
        vit_ops.colblock(g_dec_in, g_feat, rows=4 * rows, ncols=c, src_row_stride=5 * c, dst_row_stride=c)
        for v in range(4):
            vit_ops.colblock(g_all_in, g_feat, rows=rows, ncols=c, src_row_stride=4 * c, dst_row_stride=c,
                             src_col0=v * c, dst_col0=v * rows * c, accumulate=True)
        enc.backward(s_enc, g_feat, _param_sink(self.shared_encoder, accumulate, "shared_encoder.", hook))
        return loss_sum / float(numel * accumulation_steps)

    @torch.no_grad()
    def predict_peaks(self, x: torch.Tensor, soft: bool = False) -> torch.Tensor:
        _require_cuda(x, "FourCamerasBaseLine.predict_peaks")
        out = self.forward(x)
        return ops.peaks_softargmax(out.contiguous()) if soft else ops.peaks_argmax(out.contiguous())


# ====================================================================================================================
# FourCamerasDisentanglement (pytorch/CNNs.py:240-352): shared encoder, re-projection of every view's features into a
# canonical frame (InvFTL), fusion with train-mode BatchNorm, re-projection back into each view (FTL), shared decoder
# ====================================================================================================================
class FTL(nn.Module):
    """pytorch/CNNs.py:322-331: (B,400,48,48) raw-reinterpreted as groups of 4, times the (3,4) camera matrix ->
    (B,300,48,48).  Standalone call on CUDA tensors (the model itself schedules pb_ftl through its engine)."""

    kin, kout = 4, 3

    def forward(self, x, P):
        _require_cuda(x, type(self).__name__)
        b = x.shape[0]
        xin = x.contiguous().float()
        groups = xin[0].numel() // self.kin
        out = torch.empty((b, groups * self.kout), device=x.device, dtype=torch.float32)
        ops.ftl(xin, P.reshape(b, self.kout, self.kin).contiguous().float(), out, kin=self.kin, kout=self.kout,
                groups=groups, in_batch_stride=xin[0].numel(), out_batch_stride=groups * self.kout)
        return out.view(b, (x.shape[1] // self.kin) * self.kout, x.shape[2], x.shape[3])


class InvFTL(FTL):
    """pytorch/CNNs.py:335-345: (B,300,48,48) in groups of 3, times the (4,3) inverse camera matrix -> (B,400,48,48)."""

    kin, kout = 3, 4


class _DisFn(torch.autograd.Function):
    """the whole model as one autograd node (forward / backward are engine schedules of C-ABI launches)."""

    @staticmethod
    def forward(ctx, module, need, x, cams, cams_inv, *params):
        out, saved = module._run_forward(x.contiguous().float(), cams, cams_inv, save=need)
        ctx.module, ctx.saved = module, saved
        return out

    @staticmethod
    def backward(ctx, g):
        module = ctx.module
        store: Dict[str, torch.Tensor] = {}

        def grad_of(name: str, p: torch.Tensor):
            store[name] = torch.empty_like(p)
            return store[name], 0.0

        module._run_backward(ctx.saved, grad_of, None, g_out=g.contiguous().float())
        ctx.saved = None
        return (None, None, None, None, None, *[store.get(n) for n, _ in module._live_params()])


class FourCamerasDisentanglement(nn.Module):
    """pytorch/CNNs.py:240-319 (model type ALL_CAMS_DISENTANGLED_PER_WING_CNN).  forward(x, camera_matrices,
    camera_matrices_inv): x [B,16,H,W] (four 4-channel views), camera_matrices [B,4,3,4], camera_matrices_inv
    [B,4,4,3].  The four views ride through the shared encoder / decoder and the per-view 1x1 convolutions as one 4B
    batch (view-major); batch_norm3 normalises each view's re-projection with its own batch statistics."""

    def __init__(self, config, image_size, number_of_output_channels):
        super().__init__()
        self.config = config
        self.model_type = config['model type']
        self.image_size = image_size
        self.number_of_output_channels = number_of_output_channels
        self.num_base_filters = config["number of base filters"]
        self.kernel_size = config["convolution kernel size"]
        self.dilation_rate = config["dilation rate"]
        self.dropout = config["dropout ratio"]
        self.precision = config.get("precision", "bf16")
        self.shared_encoder = Encoder2DAtrous(img_size=(image_size[0], image_size[1], image_size[2] // 4),
                                              filters=self.num_base_filters, kernel_size=self.kernel_size,
                                              dilation_rate=self.dilation_rate, dropout=self.dropout,
                                              precision=self.precision)
        width = int(self.shared_encoder.get_output_size()[-1])
        self.rearrange_layer_1 = nn.Conv2d(in_channels=width, out_channels=300, kernel_size=1, padding=0)
        self.FTL = FTL()
        self.fusion_layer_1 = nn.Conv2d(in_channels=1600, out_channels=400, kernel_size=1, padding=0)
        self.fusion_layer_2 = nn.Conv2d(in_channels=400, out_channels=400, kernel_size=1, padding=0)
        self.batch_norm1 = nn.BatchNorm2d(400)
        self.batch_norm2 = nn.BatchNorm2d(400)
        self.batch_norm3 = nn.BatchNorm2d(300)
        self.relu = nn.ReLU(inplace=True)
        self.invFTL = InvFTL()
        self.rearrange_layer_2 = nn.Conv2d(in_channels=300, out_channels=width, kernel_size=1, padding=0)
        self.shared_decoder = Decoder2d(input_shape=self.shared_encoder.get_output_size(),
                                        num_output_channels=int(self.number_of_output_channels // 4),
                                        kernel_size=int(self.kernel_size), filters=int(self.num_base_filters),
                                        dropout=float(self.dropout), precision=self.precision)

    # ---- plumbing -------------------------------------------------------------------------------------------------
    def _mid(self) -> DisentangleMidEngine:
        eng = self.__dict__.get("_mid_eng")
        if eng is None or eng.precision != self.precision:
            eng = DisentangleMidEngine(self, self.precision)
            self.__dict__["_mid_eng"] = eng
        return eng

    def _live_params(self):
        """parameters that receive gradients: the stacks' BatchNorms are inert (CNNs.py:56-71), the three fusion
        BatchNorms are live."""
        return [(n, p) for n, p in self.named_parameters() if not (".bn" in n and n.startswith("shared_"))]

    def set_precision(self, precision: str):
        self.precision = precision
        self.shared_encoder.set_precision(precision)
        self.shared_decoder.set_precision(precision)
        return self

    def invalidate_packed_weights(self):
        self.shared_encoder.invalidate_packed_weights()
        self.shared_decoder.invalidate_packed_weights()
        if "_mid_eng" in self.__dict__:
            self.__dict__["_mid_eng"].invalidate()

    def repack_weights(self):
        self.shared_encoder.repack_weights()
        self.shared_decoder.repack_weights()
        mid = self.__dict__.get("_mid_eng")
        if mid is not None and not mid.repack_all():
            mid.invalidate()

    def set_grad_ready_hook(self, hook) -> None:
        self.__dict__["_grad_ready_hook"] = hook

    _views_to_batch = staticmethod(views_to_batch)
    _batch_to_views = staticmethod(batch_to_views)

    # ---- engine schedules ------------------------------------------------------------------------------------------
    def _run_mid_forward(self, first: torch.Tensor, cams: torch.Tensor, cams_inv: torch.Tensor, b: int, save: bool):
        """first [4B,h,w,256] (view-major encoder features) -> decoder input [4B,h,w,256] (= rearrange_layer_2(...) +
        first), saved tensors for the backward."""
        from . import vit_ops
        mid = self._mid()
        n4, h, w, _ = first.shape
        npix = h * w
        c3, c4 = mid.stored(300), mid.stored(400)
        dt, dev = first.dtype, first.device
        training = self.training
        # camera matrices, view-major like the batch: index v*B + b
        p_vm = cams.float().permute(1, 0, 2, 3).reshape(n4, 3, 4).contiguous()
        pinv_vm = cams_inv.float().permute(1, 0, 2, 3).reshape(n4, 4, 3).contiguous()
        r1 = mid.fwd("rearrange_layer_1", first, n4, h, w)                                   # [4B,h,w,c3]
        # InvFTL on the NCHW reinterpretation (CNNs.py:335-345): NHWC -> NCHW, groups of 3 -> 4, back, views side by side
        r1_t = vit_ops.batched_transpose(r1.view(n4, npix, c3), n4, npix, c3)               # [4B,c3,npix]
        can_t = torch.empty((n4, 400, npix), device=dev, dtype=dt)
        ops.ftl(r1_t, pinv_vm, can_t, kin=3, kout=4, groups=300 * npix // 3, in_batch_stride=c3 * npix,
                out_batch_stride=400 * npix)
        can = vit_ops.batched_transpose(can_t, n4, 400, npix)                                # [4B,npix,400]
        fus_in = torch.empty((b, h, w, 1600), device=dev, dtype=dt)                          # torch.cat(canonic, 1) :297
        for v in range(4):
            vit_ops.colblock(can, fus_in, rows=b * npix, ncols=400, src_row_stride=400, dst_row_stride=1600,
                             src_col0=v * b * npix * 400, dst_col0=v * 400)
        f1 = mid.fwd("fusion_layer_1", fus_in, b, h, w)                                      # [B,h,w,c4]
        bn1, bn2, bn3 = self.batch_norm1, self.batch_norm2, self.batch_norm3
        h1, m1, s1 = ops.batchnorm_fwd(f1.view(-1, c4), bn1.weight, bn1.bias, bn1.running_mean, bn1.running_var,
                                       groups=1, channels=400, training=training, eps=bn1.eps, momentum=bn1.momentum)
        f2 = mid.fwd("fusion_layer_2", h1.view(b, h, w, c4), b, h, w)
        h2, m2, s2 = ops.batchnorm_fwd(f2.view(-1, c4), bn2.weight, bn2.bias, bn2.running_mean, bn2.running_var,
                                       groups=1, channels=400, training=training, eps=bn2.eps, momentum=bn2.momentum)
        # FTL into each view (CNNs.py:322-331): one NCHW copy of the fused features serves the four views
        h2_t = vit_ops.batched_transpose(h2.view(b, npix, c4), b, npix, c4)                  # [B,c4,npix]
        ent_t = torch.zeros((n4, c3, npix), device=dev, dtype=dt)                            # padding rows stay zero
        ops.ftl(h2_t, p_vm, ent_t, kin=4, kout=3, groups=400 * npix // 4, in_batch_stride=c4 * npix,
                out_batch_stride=c3 * npix, in_batch_mod=b)
        ent = vit_ops.batched_transpose(ent_t, n4, c3, npix)                                 # [4B,npix,c3]
        e3, m3, s3 = ops.batchnorm_fwd(ent.view(-1, c3), bn3.weight, bn3.bias, bn3.running_mean, bn3.running_var,
                                       groups=4, channels=300, training=training, eps=bn3.eps, momentum=bn3.momentum)
        if training:
            bn1.num_batches_tracked += 1
            bn2.num_batches_tracked += 1
            bn3.num_batches_tracked += 4
        dec_in = mid.fwd("rearrange_layer_2", e3.view(n4, h, w, c3), n4, h, w, add1=first)   # + first_encoder :311-314
        saved = None
        if save:
            saved = dict(first=first, p_vm=p_vm, pinv_vm=pinv_vm, fus_in=fus_in, f1=f1, h1=h1, m1=m1, s1=s1, f2=f2,
                         h2=h2, m2=m2, s2=s2, ent=ent, e3=e3, m3=m3, s3=s3, shape=(b, h, w, c3, c4))
        return dec_in, saved

    def _run_forward(self, x: torch.Tensor, cams: torch.Tensor, cams_inv: torch.Tensor, save: bool):
        if x.shape[1] != 16:
            raise ValueError("FourCamerasDisentanglement: expected four 4-channel views (16 input channels)")
        enc, dec = self.shared_encoder._engine(), self.shared_decoder._engine()
        b = x.shape[0]
        first, s_enc = enc.forward(self._views_to_batch(x), save=save)
        if first.shape[1] * first.shape[2] != 48 * 48:
            raise ValueError("FTL / InvFTL hard-code 48 x 48 feature maps (CNNs.py:327,340): 192 x 192 crops only")
        dec_in, s_mid = self._run_mid_forward(first, cams, cams_inv, b, save)
        out4, s_dec = dec.forward(dec_in, save=save)
        saved = dict(b=b, enc=s_enc, mid=s_mid, dec=s_dec) if save else None
        return self._batch_to_views(out4), saved

    def _run_backward(self, saved: dict, grad_of, hook, g_out: Optional[torch.Tensor] = None,
                      dc_y: Optional[torch.Tensor] = None) -> None:
        """grad_of(full parameter name, parameter) -> (gradient tensor to fill, beta).  g_out: gradient w.r.t. the
        [B,C,H,W] output; or dc_y: gradient w.r.t. the head's pre-activation (fused loss path)."""
        from . import vit_ops
        enc, dec, mid = self.shared_encoder._engine(), self.shared_decoder._engine(), self._mid()

        def stack_sink(module: nn.Module, prefix: str):
            def sink(name: str):
                m = getattr(module, name)
                (dw, beta), (db, _) = grad_of(f"{prefix}{name}.weight", m.weight), grad_of(f"{prefix}{name}.bias", m.bias)
                return dw, db, beta

            def done(name: str):
                if hook is not None:
                    hook(f"{prefix}{name}.bias")
                    hook(f"{prefix}{name}.weight")
            sink.done = done
            return sink

        if dc_y is None:
            dc_y = ops.grad_ingest(self._views_to_batch(g_out), saved["dec"]["out"], dec.grad_dtype, cpad=dec.out_cpad())
        g_decin = dec.backward(saved["dec"], dc_y, stack_sink(self.shared_decoder, "shared_decoder."), need_input_grad=True)
        sm = saved["mid"]
        b, h, w, c3, c4 = sm["shape"]
        n4, npix = 4 * b, h * w
        dt, dev = g_decin.dtype, g_decin.device
        msink = stack_sink(self, "")

        def bn_bwd(tag: str, bn: nn.BatchNorm2d, x2d, y2d, gy2d, mean, rstd, groups: int, channels: int):
            (dg, beta), (db, _) = grad_of(f"{tag}.weight", bn.weight), grad_of(f"{tag}.bias", bn.bias)
            gx = ops.batchnorm_bwd(x2d, y2d, gy2d, bn.weight, mean, rstd, dg, db, groups=groups, channels=channels,
                                   beta_acc=beta)
            if hook is not None:
                hook(f"{tag}.bias")
                hook(f"{tag}.weight")
            return gx

        # dec_in = rearrange_layer_2(e3) + first
        mid.wgrad("rearrange_layer_2", sm["e3"].view(n4, h, w, c3), g_decin, n4, h, w, msink)
        g_e3 = mid.dgrad("rearrange_layer_2", g_decin, n4, h, w, c3)
        g_ent = bn_bwd("batch_norm3", self.batch_norm3, sm["ent"].view(-1, c3), sm["e3"].view(-1, c3), g_e3.view(-1, c3),
                       sm["m3"], sm["s3"], 4, 300)
        # FTL backward: the transposed camera matrices, the four views' shares summed into the fused features
        g_ent_t = vit_ops.batched_transpose(g_ent.view(n4, npix, c3), n4, npix, c3)           # [4B,c3,npix]
        g_h2_t = torch.zeros((b, c4, npix), device=dev, dtype=dt)
        p_t = sm["p_vm"].transpose(1, 2).contiguous()                                         # [4B,4,3]
        for v in range(4):
            ops.ftl(g_ent_t[v * b:(v + 1) * b], p_t[v * b:(v + 1) * b].contiguous(), g_h2_t, kin=3, kout=4,
                    groups=300 * npix // 3, in_batch_stride=c3 * npix, out_batch_stride=c4 * npix, accumulate=v > 0)
        g_h2 = vit_ops.batched_transpose(g_h2_t, b, c4, npix)                                 # [B,npix,c4]
        g_f2 = bn_bwd("batch_norm2", self.batch_norm2, sm["f2"].view(-1, c4), sm["h2"], g_h2.view(-1, c4), sm["m2"],
                      sm["s2"], 1, 400)
        mid.wgrad("fusion_layer_2", sm["h1"].view(b, h, w, c4), g_f2.view(b, h, w, c4), b, h, w, msink)
        g_h1 = mid.dgrad("fusion_layer_2", g_f2.view(b, h, w, c4), b, h, w, c4)
        g_f1 = bn_bwd("batch_norm1", self.batch_norm1, sm["f1"].view(-1, c4), sm["h1"], g_h1.view(-1, c4), sm["m1"],
                      sm["s1"], 1, 400)
        mid.wgrad("fusion_layer_1", sm["fus_in"], g_f1.view(b, h, w, c4), b, h, w, msink)
        g_fus = mid.dgrad("fusion_layer_1", g_f1.view(b, h, w, c4), b, h, w, 1600)            # [B,h,w,1600]
        g_can = torch.empty((n4, npix, 400), device=dev, dtype=dt)                            # torch.cat backward: split
        for v in range(4):
            vit_ops.colblock(g_fus, g_can, rows=b * npix, ncols=400, src_row_stride=1600, dst_row_stride=400,
                             src_col0=v * 400, dst_col0=v * b * npix * 400)
        g_can_t = vit_ops.batched_transpose(g_can, n4, npix, 400)                             # [4B,400,npix]
        g_r1_t = torch.zeros((n4, c3, npix), device=dev, dtype=dt)
        ops.ftl(g_can_t, sm["pinv_vm"].transpose(1, 2).contiguous(), g_r1_t, kin=4, kout=3, groups=400 * npix // 4,
                in_batch_stride=400 * npix, out_batch_stride=c3 * npix)
        g_r1 = vit_ops.batched_transpose(g_r1_t, n4, c3, npix).view(n4, h, w, c3)
        mid.wgrad("rearrange_layer_1", sm["first"], g_r1, n4, h, w, msink)
        g_first = mid.dgrad("rearrange_layer_1", g_r1, n4, h, w, int(sm["first"].shape[-1]), add0=g_decin)
        enc.backward(saved["enc"], g_first, stack_sink(self.shared_encoder, "shared_encoder."))

    # ---- nn.Module surface --------------------------------------------------------------------------------------------
    def forward(self, x, camera_matrices, camera_matrices_inv):
        _require_cuda(x, "FourCamerasDisentanglement")
        params = [p for _, p in self._live_params()]
        need = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        return _DisFn.apply(self, need, x, camera_matrices, camera_matrices_inv, *params)

    @torch.no_grad()
    def train_step(self, x: torch.Tensor, target: Optional[torch.Tensor] = None, *, camera_matrices: torch.Tensor,
                   camera_matrices_inv: torch.Tensor, points: Optional[torch.Tensor] = None, sigma: float = 3.0,
                   accumulation_steps: int = 1, accumulate: bool = False, loss_scale: float = 1.0) -> torch.Tensor:
        """forward + MSE + backward, gradients written into ``param.grad`` (flat buckets); the head's epilogue
        computes the loss and its gradient (pb_convT_mse_fused) when the last layer tiles."""
        _require_cuda(x, "FourCamerasDisentanglement.train_step")
        enc, dec = self.shared_encoder._engine(), self.shared_decoder._engine()
        hook = self.__dict__.get("_grad_ready_hook")
        b = x.shape[0]
        first, s_enc = enc.forward(self._views_to_batch(x.contiguous().float()), save=True)
        dec_in, s_mid = self._run_mid_forward(first, camera_matrices, camera_matrices_inv, b, True)
        numel = b * self.number_of_output_channels * x.shape[2] * x.shape[3]
        tgt4 = self._views_to_batch(target.contiguous().float()) if target is not None else None
        pts4 = self._views_to_batch(points.contiguous().float()) if points is not None and target is None else None
        if dec.head_fusable():
            loss_sum, dc_y, s_dec = dec.forward_loss(dec_in, None, target=tgt4, points=pts4, sigma=sigma,
                                                     accumulation_steps=accumulation_steps, loss_scale=loss_scale)
        else:
            out4, s_dec = dec.forward(dec_in, save=True)
            loss_sum, _, dc_y = ops.mse_loss_fwd_bwd(out4, tgt4, points=pts4, sigma=sigma,
                                                     accumulation_steps=accumulation_steps, loss_scale=loss_scale,
                                                     grad_nhwc_dtype=dec.grad_dtype, cpad=dec.out_cpad())
        beta = 1.0 if accumulate else 0.0

        def grad_of(name: str, p: torch.Tensor):
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            return p.grad, beta

        self._run_backward(dict(b=b, enc=s_enc, mid=s_mid, dec=s_dec), grad_of, hook, dc_y=dc_y)
        return loss_sum / float(numel * accumulation_steps)

    @torch.no_grad()
    def predict_peaks(self, x: torch.Tensor, camera_matrices: torch.Tensor, camera_matrices_inv: torch.Tensor,
                      soft: bool = False) -> torch.Tensor:
        out, _ = self._run_forward(x.contiguous().float(), camera_matrices, camera_matrices_inv, save=False)
        return ops.peaks_softargmax(out) if soft else ops.peaks_argmax(out)
