"""Execution engine of the conv heatmap networks: schedules the C-ABI contractions and
elementwise kernels for forward and backward, owns packed-weight caches and saved activations.

Layout in HBM (DESIGN.md "data layout"): activations NHWC in ``act_dtype`` (bf16 in the fast
mode, fp32 in fp32 mode); one sign bit per element (``mask``) is stored by every LeakyReLU
epilogue so the backward never re-reads pre-activations; network input is the reference's NCHW
fp32 crop tensor, network output its NCHW fp32 heatmap tensor (pytorch/CNNs.py:183-186).

Backward fusion: the input-gradient contraction of layer L adds the skip gradient, stores the
plain gradient G (needed by the next residual add) and multiplies by LeakyReLU'(layer L-1) in
its epilogue, so ``dC`` (gradient w.r.t. the pre-activation, the operand of both wgrad and the
next dgrad) is produced without a separate elementwise pass.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, List, Optional, Tuple

import torch
from torch import nn

from . import ops
from .ops import Contraction, PB_ACT_LRELU, PB_ACT_MASKMUL, PB_ACT_NONE

GradSink = Callable[[str], Tuple[torch.Tensor, Optional[torch.Tensor], float]]

PRECISIONS = ("bf16", "fp16", "fp32")


def tc_globally_enabled() -> bool:
    return os.environ.get("POSEB200_DISABLE_TC", "0") != "1"


class Layer:
    """One conv-like layer bound to its nn.Module parameters."""

    def __init__(self, name: str, module: nn.Module, spec: Contraction):
        self.name, self.module, self.spec = name, module, spec
        self._packed: Dict[Tuple[str, torch.dtype, int, int], Tuple[tuple, torch.Tensor]] = {}

    def _tag(self) -> tuple:
        """identity of the weight values a packed operand was built from: torch's version counter and storage (moved
        by autograd-visible writes / rebinding) and the package's weight generation (moved by C-ABI writers such as
        the fused Adam, which torch's counter does not see)."""
        w = self.module.weight
        return (w._version, w.data_ptr(), ops.weights_generation())

    def packed(self, role: str, dtype: torch.dtype, ipad: int = 0, jpad: int = 0) -> torch.Tensor:
        key = (role, dtype, ipad, jpad)
        tag = self._tag()
        hit = self._packed.get(key)
        if hit is not None and hit[0] == tag:
            return hit[1]
        t = ops.pack_weights(self.module.weight, self.spec, role, dtype, ipad, jpad,
                             dst=hit[1] if hit is not None else None)
        self._packed[key] = (tag, t)
        return t

    def retag(self) -> None:
        """the cached operands were just refreshed in place (ConvStack.repack_all)."""
        tag = self._tag()
        for k, (_, t) in list(self._packed.items()):
            self._packed[k] = (tag, t)

    def invalidate(self) -> None:
        self._packed.clear()

    def pack_items(self) -> list:
        """pb_pack_weights_args refreshing every cached operand of this layer IN PLACE (same dst tensors)."""
        out = []
        w = self.module.weight
        for (role, dtype, ipad, jpad), (tag, t) in self._packed.items():
            if tag[:2] != (w._version, w.data_ptr()):
                return []          # rebound / modified through autograd: fall back to lazy re-packing
            a, _ = ops.pack_weights_args(w, self.spec, role, dtype, ipad, jpad, dst=t)
            out.append(a)
        return out


class ConvStack:
    """Shared machinery of Encoder2DAtrous / Decoder2d / CNN_Decoder engines."""

    def __init__(self, precision: str = "bf16"):
        self.set_precision(precision)
        self.layers: Dict[str, Layer] = {}
        self._ws: Optional[torch.Tensor] = None

    def set_precision(self, precision: str) -> None:
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {PRECISIONS}")
        self.precision = precision
        # forward activations / packed forward weights, and everything on the gradient side (dC, G, packed
        # input-gradient weights).  "fp16": IEEE-half forward operands (the reference's own autocast dtype,
        # pytorch/train_pytorch.py:133) against bf16 gradients, which need no loss scaling
        self.act_dtype, self.grad_dtype = {"bf16": (torch.bfloat16, torch.bfloat16),
                                           "fp16": (torch.float16, torch.bfloat16),
                                           "fp32": (torch.float32, torch.float32)}[precision]

    def invalidate(self) -> None:
        for l in self.layers.values():
            l.invalidate()
        first = getattr(self, "_first_lin", None)
        if first is not None:
            first.invalidate()
        self._pack_table = None

    def _all_layers(self) -> list:
        ls = list(self.layers.values())
        first = getattr(self, "_first_lin", None)
        return ls + ([first] if first is not None else [])

    def repack_all(self) -> bool:
        """After an in-place optimiser step (the fused Adam writes the flat parameter buffer through the C ABI, so
        tensor versions do not move): refresh every cached packed operand with ONE launch.  Returns False when
        the caches cannot be refreshed in place (caller invalidates instead)."""
        layers = self._all_layers()
        sig = tuple((id(l), k, v[1].data_ptr(), l.module.weight.data_ptr()) for l in layers for k, v in l._packed.items())
        if not sig:
            return True
        cached = getattr(self, "_pack_table", None)
        if cached is None or cached[0] != sig:
            items = []
            for l in layers:
                got = l.pack_items()
                if len(got) != len(l._packed):
                    return False
                items.extend(got)
            dev = layers[0].module.weight.device
            table, max_elems = ops.pack_table(items, dev)
            cached = (sig, table, len(items), max_elems)
            self._pack_table = cached
        ops.pack_weights_multi(cached[1], cached[2], cached[3])
        for l in layers:
            l.retag()
        return True

    def first_layer_tc(self) -> Optional[Layer]:
        """tensor-core form of the first (NCHW-input) layer, or None when it runs on CUDA cores."""
        return None

    # ---- implementation choice per contraction -------------------------------------------
    def impl_for(self, spec: Contraction, what: str) -> str:
        """'tc' (tcgen05) when the bf16 path tiles this shape, else 'simt'."""
        if self.precision == "fp32" or not tc_globally_enabled():
            return self._simt(spec, what)
        from . import tc_support
        return "tc" if tc_support.supported(spec, what) else self._simt(spec, what)

    def _simt(self, spec: Contraction, what: str) -> str:
        if self.precision == "fp16":
            raise RuntimeError(f"precision 'fp16' runs on the tcgen05 kernels only; {what} of {spec.kind} "
                               f"{spec.cin}->{spec.cout} is outside their tiling (use 'bf16' or 'fp32')")
        return "simt"

    def _workspace(self, spec: Contraction, pixels: int, device) -> torch.Tensor:
        # upper bound over the kernels that may take this contraction (148 = one wave of the halo kernel)
        need = max(ops.choose_ksplit(spec, pixels, impl="tc"), ops.choose_ksplit(spec, pixels), 148) * \
            ops.wgrad_workspace_len(spec, 64 * ((spec.cin + 63) // 64))
        if self._ws is None or self._ws.numel() < need or self._ws.device != device:
            self._ws = torch.empty(need, device=device, dtype=torch.float32)
        return self._ws

    # ---- single-layer helpers --------------------------------------------------------------
    def fwd_layer(self, layer: Layer, x: torch.Tensor, n: int, ih: int, iw: int, *, add1=None, save: bool,
                  in_nchw: bool = False, out_nchw: bool = False, pool_out: Optional[torch.Tensor] = None):
        """y = lrelu(conv(x) + b) (+ add1).  returns (y, mask|None, y_for_wgrad|None): the third is the tensor a later
        weight gradient reads as its activation operand -- y itself, or its bf16 twin in the "fp16" precision
        (tensor-core operands of one MMA share a format and the gradients are bf16)."""
        s = layer.spec
        oh, ow = s.out_hw(ih, iw)
        impl = getattr(layer, "force_impl", None) or (self.impl_for(s, "fwd") if not in_nchw else self._simt(s, "fwd"))
        mask = None
        if save and not out_nchw:
            mask = torch.empty((n * oh * ow, (s.cout + 31) // 32), device=x.device, dtype=torch.int32)
        cin_stored = s.cin if in_nchw else int(x.shape[-1])  # > cin when the operand is zero padded
        if impl == "tc":
            from . import tc_support
            w = layer.packed("oi", self.act_dtype, tc_support.pad_n(s.cout), jpad=cin_stored)
        else:
            if cin_stored != s.cin:
                raise RuntimeError("simt forward expects an unpadded activation tensor")
            w = layer.packed("io", torch.float32)
        twin = None
        if save and not out_nchw and self.act_dtype != self.grad_dtype:
            twin = torch.empty((n, oh, ow, s.cout), device=x.device, dtype=self.grad_dtype)
        # pool_out: the epilogue also emits lrelu(maxpool2x2(y)) -- and, when nothing is saved for a backward pass,
        # only that (y itself never reaches HBM)
        y = ops.conv(impl, x, w, s.fwd_taps(), n, ih, iw, cin_stored, oh, ow, s.cout, bias=layer.module.bias,
                     act=PB_ACT_LRELU, add1=add1, mask_out=mask, act_dtype=self.act_dtype, in_nchw=in_nchw,
                     out_nchw=out_nchw, out2=twin, pool_out=pool_out, pool_only=pool_out is not None and not save)
        return y, mask, (twin if twin is not None else y)

    def dgrad_layer(self, layer: Layer, dc: torch.Tensor, n: int, ih: int, iw: int, *, add0=None, want_g: bool,
                    mask_prev=None) -> Tuple[Optional[torch.Tensor], torch.Tensor]:
        """input gradient of `layer` (whose forward input is [n, ih, iw, cin]).
        v = dgrad(dc) + add0;  G = v (stored when want_g);  out = v * lrelu'(mask_prev) if mask_prev else v.
        returns (G|None, out)."""
        s = layer.spec
        oh, ow = s.out_hw(ih, iw)
        impl = self.impl_for(s, "dgrad")
        kdim = s.cout
        if impl == "tc":
            kdim = dc.shape[-1]  # channel-padded gradient of the last layer: K extent as stored
            w = layer.packed("io", self.grad_dtype, jpad=kdim)
        else:
            if dc.shape[-1] != s.cout:
                raise RuntimeError("simt dgrad expects an unpadded gradient tensor")
            w = layer.packed("oi", torch.float32)
        g = None
        if want_g and mask_prev is not None:
            g = torch.empty((n, ih, iw, s.cin), device=dc.device, dtype=self.grad_dtype)
        out = ops.conv(impl, dc, w, s.dgrad_taps(), n, oh, ow, kdim, ih, iw, s.cin, add0=add0, pre_out=g,
                       act=PB_ACT_MASKMUL if mask_prev is not None else PB_ACT_NONE, mask_in=mask_prev,
                       act_dtype=self.grad_dtype, prof_cin=s.cout)
        if want_g and mask_prev is None:
            g = out
        return g, out

    def wgrad_layer(self, layer: Layer, a_in: torch.Tensor, dc: torch.Tensor, n: int, ih: int, iw: int,
                    sink: GradSink, a_nchw: bool = False, dbias_sum: Optional[torch.Tensor] = None) -> None:
        """dbias_sum: the bias gradient is already known (the fused head's epilogue summed it): only the weight
        contraction runs here and db = beta * db + dbias_sum."""
        s = layer.spec
        impl = getattr(layer, "force_impl", None) or (self.impl_for(s, "wgrad") if not a_nchw else self._simt(s, "wgrad"))
        dw, db, beta = sink(layer.name)
        pixels = n * (ih * iw if s.kind == "convT2" else s.out_hw(ih, iw)[0] * s.out_hw(ih, iw)[1])
        if dbias_sum is not None and db is not None:
            from . import vit_ops
            vit_ops.colsum(dbias_sum, db, int(dbias_sum.shape[0]), s.cout, alpha=1.0, beta=beta)   # per-CTA rows -> db
            db = None
        ops.wgrad(impl, s, a_in, dc, n, ih, iw, dw, db, act_dtype=self.grad_dtype, a_nchw=a_nchw, beta=beta,
                  workspace=self._workspace(s, pixels, dc.device))
        done = getattr(sink, "done", None)
        if done is not None:
            done(layer.name)

    # ---- residual triple: a = f(in); b = f(a)+a; c = f(b)+b -------------------------------
    def pool_fused(self, layer: Layer, oh: int, ow: int, save: bool) -> bool:
        """the 2x2 max-pool + LeakyReLU behind `layer` runs inside its epilogue (tensor-core path; the "fp16" training
        step keeps the separate kernel, which also writes the pooled tensor's bf16 twin)."""
        from . import tc_support
        return (self.impl_for(layer.spec, "fwd") == "tc" and tc_support.pool_fusable(layer.spec, oh, ow)
                and not (save and self.act_dtype != self.grad_dtype))

    def fwd_triple(self, names: List[str], x, n, ih, iw, save: bool, saved: dict, in_nchw: bool = False, x_w=None,
                   pool: bool = False):
        """x_w: the tensor the first layer's weight gradient reads (x, or its bf16 twin in the "fp16" precision).
        pool: the caller pools the triple's output; when the last layer's epilogue can do it, the pooled tensor is
        returned as a fifth element (else None).  returns (c, oh, ow, c_w[, pooled])."""
        la, lb, lc = (self.layers[k] for k in names)
        first = self.first_layer_tc() if in_nchw else None
        x_w = x if x_w is None else x_w
        direct = None
        if first is not None:
            # Cin = 4 is far below the 64-channel K chunk: conv1 runs as a 1-tap tensor-core contraction over the
            # crop's im2col (k = ci*9+tap).  Forward (csrc/tc_conv1.cu) and weight gradient (csrc/tc_wgrad1.cu) build
            # that operand inside their kernels; the im2col TENSOR is only materialised for shapes they do not take.
            la = first
            kpad = 64 * ((la.spec.cin + 63) // 64)
            x_nchw = x
            cin_img = int(x_nchw.shape[1])
            twins = save and self.act_dtype != self.grad_dtype          # "fp16" training: bf16 twins for the weight gradients
            if ops.conv_first_supported(cin_img, self.first_ksize, la.spec.cout):
                from . import tc_support
                oh, ow = ih, iw
                ma = None
                if save:
                    ma = torch.empty((n * oh * ow, (la.spec.cout + 31) // 32), device=x.device, dtype=torch.int32)
                wp = la.packed("oi", self.act_dtype, tc_support.pad_n(la.spec.cout), jpad=kpad)
                a2 = torch.empty((n, oh, ow, la.spec.cout), device=x.device, dtype=self.grad_dtype) if twins else None
                a = ops.conv_first(x_nchw, wp, la.module.bias, la.spec.cout, self.first_dilation, self.act_dtype,
                                   mask_out=ma, ksize=self.first_ksize, out2=a2)
                direct = (a, ma, a2 if twins else a)
                if save:
                    if ops.wgrad_first_supported(cin_img, self.first_ksize, la.spec.cout, self.grad_dtype):
                        x_w = x_nchw      # the weight gradient builds its operand from the crop too (csrc/tc_wgrad1.cu)
                    else:
                        x_w = ops.im2col_first(x_nchw, self.first_ksize, self.first_dilation, kpad, self.grad_dtype)
            else:
                x = ops.im2col_first(x_nchw, self.first_ksize, self.first_dilation, kpad, self.act_dtype)
                x_w = x
                if save and self.act_dtype != self.grad_dtype:
                    x_w = ops.im2col_first(x_nchw, self.first_ksize, self.first_dilation, kpad, self.grad_dtype)
            in_nchw = False
        if direct is not None:
            a, ma, a_w = direct
        else:
            a, ma, a_w = self.fwd_layer(la, x, n, ih, iw, save=save, in_nchw=in_nchw)
        oh, ow = la.spec.out_hw(ih, iw)
        b, mb, b_w = self.fwd_layer(lb, a, n, oh, ow, add1=a, save=save)
        pooled = None
        if pool and self.pool_fused(lc, oh, ow, save):
            pooled = torch.empty((n, oh // 2, ow // 2, lc.spec.cout), device=b.device, dtype=self.act_dtype)
        c, mc, c_w = self.fwd_layer(lc, b, n, oh, ow, add1=b, save=save, pool_out=pooled)
        if save:
            saved[names[0]] = (x_w, ma, ih, iw)
            saved[names[1]] = (a_w, mb, oh, ow)
            saved[names[2]] = (b_w, mc, oh, ow)
        if pool:
            return c, oh, ow, c_w, pooled
        return c, oh, ow, c_w

    def bwd_triple(self, names: List[str], g_c, dc_c, n, saved: dict, sink: GradSink, need_input_grad: bool,
                   in_nchw: bool = False, mask_below=None):
        """g_c: plain gradient w.r.t. the triple's output c; dc_c = g_c * lrelu'(mask_c).
        returns the plain gradient w.r.t. the triple's input (or None); with `mask_below` (sign mask of the
        layer that produced the triple's input) returns (g_in, g_in * lrelu'(mask_below)) from one epilogue."""
        la, lb, lc = (self.layers[k] for k in names)
        x_in, ma, ih, iw = saved[names[0]]
        if in_nchw and x_in.dim() == 4 and x_in.dtype != torch.float32:
            la, in_nchw = self.first_layer_tc(), False  # saved operand is the im2col'd crop
        a, mb, oh, ow = saved[names[1]]
        b, _mc, _, _ = saved[names[2]]
        self.wgrad_layer(lc, b, dc_c, n, oh, ow, sink)
        g_b, dc_b = self.dgrad_layer(lc, dc_c, n, oh, ow, add0=g_c, want_g=True, mask_prev=mb)
        self.wgrad_layer(lb, a, dc_b, n, oh, ow, sink)
        _, dc_a = self.dgrad_layer(lb, dc_b, n, oh, ow, add0=g_b, want_g=False, mask_prev=ma)
        first = self.first_layer_tc() if in_nchw else None
        if (first is not None and x_in.dtype == torch.float32 and x_in.dim() == 4 and dc_a.dtype == torch.bfloat16
                and ops.wgrad_first_supported(int(x_in.shape[1]), self.first_ksize, la.spec.cout, dc_a.dtype)):
            # conv1's weight / bias gradient straight from the NCHW crop (no im2col tensor)
            dw, db, beta = sink(la.name)
            ops.wgrad_first(x_in, dc_a, dw, db, self.first_dilation, beta=beta,
                            workspace=self._workspace(first.spec, n * ih * iw, dc_a.device))
            done = getattr(sink, "done", None)
            if done is not None:
                done(la.name)
        else:
            self.wgrad_layer(la, x_in, dc_a, n, ih, iw, sink, a_nchw=in_nchw)
        if not need_input_grad:
            return None
        if mask_below is not None:
            return self.dgrad_layer(la, dc_a, n, ih, iw, want_g=True, mask_prev=mask_below)
        _, g_in = self.dgrad_layer(la, dc_a, n, ih, iw, want_g=False, mask_prev=None)
        return g_in


class EncoderEngine(ConvStack):
    """Encoder2DAtrous (pytorch/CNNs.py:9-88)."""

    def __init__(self, module: nn.Module, precision: str):
        super().__init__(precision)
        f, cin, d = module.filters, int(module.image_size[-1]), module.dilation_rate
        k = module.kernel_size
        chans = [(cin, f), (f, f), (f, f), (f, 2 * f), (2 * f, 2 * f), (2 * f, 2 * f), (2 * f, 4 * f),
                 (4 * f, 4 * f), (4 * f, 4 * f)]
        for i, (ci, co) in enumerate(chans, 1):
            name = f"conv{i}"
            self.layers[name] = Layer(name, getattr(module, name), Contraction("conv", ci, co, dilation=d, ksize=k))
        if 2 * d * ((k - 1) // 2) != 2 * module.padding:
            raise ValueError("Encoder2DAtrous: only 'same' geometry (padding == dilation*(k-1)/2) is supported")
        self.first_ksize, self.first_dilation = k, d
        # conv1 as a 1-tap contraction over the im2col'd crop: conv1.weight viewed as [Cout][Cin*k*k]
        self._first_lin = Layer("conv1", module.conv1, Contraction("linear", cin * k * k, f))
        self._first_lin.force_impl = "tc"

    def first_layer_tc(self) -> Optional[Layer]:
        from . import tc_support
        s = self._first_lin.spec
        ok = (self.precision in ("bf16", "fp16") and tc_globally_enabled() and tc_support.ENABLED["fwd"]
              and tc_support.ENABLED["wgrad"] and s.cout % 16 == 0 and s.cout <= 256 and s.cin <= 64)
        return self._first_lin if ok else None

    def forward(self, x_nchw: torch.Tensor, save: bool):
        n, _, h, w = x_nchw.shape
        saved: dict = {"n": n}
        cur, cur_w, ih, iw = x_nchw, None, h, w
        twins = save and self.act_dtype != self.grad_dtype
        for stage in range(3):
            names = [f"conv{3 * stage + j}" for j in (1, 2, 3)]
            if stage < 2:
                c, ih, iw, c_w, pooled = self.fwd_triple(names, cur, n, ih, iw, save, saved, in_nchw=(stage == 0),
                                                         x_w=cur_w, pool=True)
            else:
                c, ih, iw, c_w = self.fwd_triple(names, cur, n, ih, iw, save, saved, in_nchw=False, x_w=cur_w)
            if stage < 2:
                if pooled is not None:          # the conv epilogue pooled already (CNNs.py:77,82)
                    cur, cur_w = pooled, None
                elif twins:
                    cur, cur_w = ops.maxpool_lrelu_fwd(c, twin=True)
                else:
                    cur, cur_w = ops.maxpool_lrelu_fwd(c), None
                if save:
                    saved[f"pool{stage}"] = c
                ih, iw = ih // 2, iw // 2
            else:
                cur, cur_w = c, c_w
        saved["out_w"] = cur_w if cur_w is not None else cur      # what the next module's first weight gradient reads
        return cur, saved

    def backward(self, saved: dict, g_out: torch.Tensor, sink: GradSink, dc_out: Optional[torch.Tensor] = None) -> None:
        """g_out: plain gradient w.r.t. the encoder output (NHWC grad_dtype); dc_out: the same times
        LeakyReLU'(conv9) when the producer's epilogue already applied it (fused train step)."""
        n = saved["n"]
        mask9 = saved["conv9"][1]
        g_c = g_out
        dc_c = dc_out if dc_out is not None else ops.add(g_out, None, mask=mask9)
        for stage in (2, 1, 0):
            names = [f"conv{3 * stage + j}" for j in (1, 2, 3)]
            g_in = self.bwd_triple(names, g_c, dc_c, n, saved, sink, need_input_grad=stage > 0,
                                   in_nchw=(stage == 0))
            if stage > 0:
                x_pool = saved[f"pool{stage - 1}"]
                mask_prev = saved[f"conv{3 * stage}"][1]
                g_c, dc_c = ops.maxpool_lrelu_bwd(x_pool, g_in, mask_prev)


class DecoderEngine(ConvStack):
    """Decoder2d (pytorch/CNNs.py:92-157)."""

    def __init__(self, module: nn.Module, precision: str):
        super().__init__(precision)
        cin = int(module.input_shape[-1])
        mid, cout = cin // 2, module.num_output_channels
        kinds = [("convT2", cin, mid), ("convT1", mid, mid), ("convT1", mid, mid), ("convT2", mid, cout)]
        for i, (kind, ci, co) in enumerate(kinds, 1):
            name = f"conv2dTranspose{i}"
            self.layers[name] = Layer(name, getattr(module, name), Contraction(kind, ci, co, ksize=module.kernel_size))
        if module.kernel_size != 3:
            raise ValueError("Decoder2d: kernel size 3 is the only geometry the reference's padding=1 supports")

    names3 = ["conv2dTranspose1", "conv2dTranspose2", "conv2dTranspose3"]

    def out_cpad(self) -> int:
        last = self.layers["conv2dTranspose4"]
        if self.impl_for(last.spec, "dgrad") == "tc" or self.impl_for(last.spec, "wgrad") == "tc":
            from . import tc_support
            return tc_support.pad_n(last.spec.cout)
        return last.spec.cout

    def forward(self, x_nhwc: torch.Tensor, save: bool, x_w: Optional[torch.Tensor] = None):
        """x_w: the tensor conv2dTranspose1's weight gradient reads (default x_nhwc; the encoder's bf16 twin in the
        "fp16" precision, or a converted copy when the caller has none)."""
        n, ih, iw, _ = x_nhwc.shape
        saved: dict = {"n": n}
        if save and x_w is None and x_nhwc.dtype != self.grad_dtype:
            x_w = x_nhwc.to(self.grad_dtype)
        d3, oh, ow, d3_w = self.fwd_triple(self.names3, x_nhwc, n, ih, iw, save, saved, x_w=x_w)
        last = self.layers["conv2dTranspose4"]
        y, _, _ = self.fwd_layer(last, d3, n, oh, ow, save=False, out_nchw=True)
        if save:
            saved["conv2dTranspose4"] = (d3_w, None, oh, ow)
            saved["out"] = y
        return y, saved

    # ---- fused head (pb_convT_argmax_fused / pb_convT_mse_fused): the last layer consumes its heatmaps on chip ----
    def head_fusable(self) -> bool:
        """the last layer runs on the halo tcgen05 kernel with 16-channel padding (what the fused epilogues tile)."""
        import os
        last = self.layers["conv2dTranspose4"].spec
        return (self.precision in ("bf16", "fp16") and tc_globally_enabled() and self.impl_for(last, "fwd") == "tc"
                and self.impl_for(last, "dgrad") == "tc" and self.impl_for(last, "wgrad") == "tc"
                and last.cin % 64 == 0 and last.cout <= 256 and os.environ.get("POSEB200_NO_HEAD_FUSION", "0") != "1")

    def _head_operands(self, d3: torch.Tensor):
        from . import tc_support
        last = self.layers["conv2dTranspose4"]
        s = last.spec
        w = last.packed("oi", self.act_dtype, tc_support.pad_n(s.cout), jpad=int(d3.shape[-1]))
        return last, s, w

    def forward_peaks(self, x_nhwc: torch.Tensor, want_values: bool = False):
        """decoder forward whose last layer emits per-map arg-max peaks instead of heatmaps (inference)."""
        n, ih, iw, _ = x_nhwc.shape
        d3, oh, ow, _ = self.fwd_triple(self.names3, x_nhwc, n, ih, iw, False, {})
        last, s, w = self._head_operands(d3)
        return ops.head_argmax_fused(d3, w, s.fwd_taps(), n, oh, ow, int(d3.shape[-1]), s.cout,
                                     bias=last.module.bias, want_values=want_values)

    def forward_loss(self, x_nhwc: torch.Tensor, x_w: Optional[torch.Tensor], *, target=None, points=None,
                     sigma: float = 3.0, accumulation_steps: int = 1, loss_scale: float = 1.0):
        """training forward whose last layer emits (sum of squared errors, dC of the head) instead of heatmaps.
        returns (loss_sum, dc_y, saved) with `saved` ready for backward()."""
        n, ih, iw, _ = x_nhwc.shape
        saved: dict = {"n": n}
        if x_w is None and x_nhwc.dtype != self.grad_dtype:
            x_w = x_nhwc.to(self.grad_dtype)
        d3, oh, ow, d3_w = self.fwd_triple(self.names3, x_nhwc, n, ih, iw, True, saved, x_w=x_w)
        last, s, w = self._head_operands(d3)
        dbias = None
        if last.module.bias is not None and ops.head_folded_supported(int(d3.shape[-1]), s.cout, d3.dtype):
            dbias = ops.head_dbias_buffer(s.cout, d3.device)    # per-CTA partial sums from the head's own epilogue
        loss_sum, dc_y = ops.head_mse_fused(d3, w, s.fwd_taps(), n, oh, ow, int(d3.shape[-1]), s.cout,
                                            bias=last.module.bias, target=target, points=points, sigma=sigma,
                                            accumulation_steps=accumulation_steps, loss_scale=loss_scale,
                                            dbias_out=dbias)
        saved["conv2dTranspose4"] = (d3_w, None, oh, ow)
        saved["head_dbias"] = dbias
        return loss_sum, dc_y, saved

    def backward(self, saved: dict, dc_y: torch.Tensor, sink: GradSink, need_input_grad: bool, mask_below=None):
        """dc_y: gradient w.r.t. the last layer's pre-activation, NHWC grad_dtype [n, 4h, 4w, cpad]."""
        n = saved["n"]
        last = self.layers["conv2dTranspose4"]
        d3, _, oh, ow = saved["conv2dTranspose4"]
        mask3 = saved["conv2dTranspose3"][1]
        self.wgrad_layer(last, d3, dc_y, n, oh, ow, sink, dbias_sum=saved.get("head_dbias"))
        g_d3, dc_d3 = self.dgrad_layer(last, dc_y, n, oh, ow, want_g=True, mask_prev=mask3)
        return self.bwd_triple(self.names3, g_d3, dc_d3, n, saved, sink, need_input_grad, mask_below=mask_below)


class PointwiseEngine(ConvStack):
    """A 1x1 convolution with bias on NHWC pixels (FourCamerasBaseLine.shared_conv2d, pytorch/CNNs.py:205-208,229):
    a 1-tap contraction over [pixels, Cin] rows, the residual add of ``conv(x) + x`` fused into the epilogue of the
    forward kernel and of the input-gradient kernel."""

    def __init__(self, module: nn.Module, name: str, precision: str):
        super().__init__(precision)
        conv = getattr(module, name)
        if tuple(conv.kernel_size) != (1, 1) or tuple(conv.padding) != (0, 0) or tuple(conv.stride) != (1, 1):
            raise ValueError("PointwiseEngine: 1x1 / stride 1 / no padding only")
        self.name = name
        self.layers[name] = Layer(name, conv, Contraction("linear", conv.in_channels, conv.out_channels))

    def forward(self, x: torch.Tensor, residual: bool) -> torch.Tensor:
        layer = self.layers[self.name]
        s = layer.spec
        rows = x.numel() // s.cin
        impl = self.impl_for(s, "fwd")
        if impl == "tc":
            from . import tc_support
            w = layer.packed("oi", self.act_dtype, tc_support.pad_n(s.cout))
        else:
            w = layer.packed("io", torch.float32)
        y = ops.conv(impl, x, w, s.fwd_taps(), 1, 1, rows, s.cin, 1, rows, s.cout, bias=layer.module.bias,
                     act=PB_ACT_NONE, add1=x if residual else None, act_dtype=self.act_dtype)
        return y.view(*x.shape[:-1], s.cout)

    def backward(self, x: torch.Tensor, g: torch.Tensor, sink: GradSink, residual: bool,
                 need_input_grad: bool = True) -> Optional[torch.Tensor]:
        layer = self.layers[self.name]
        s = layer.spec
        rows = g.numel() // s.cout
        dw, db, beta = sink(self.name)
        ops.wgrad(self.impl_for(s, "wgrad"), s, x, g, 1, 1, rows, dw, db, act_dtype=self.grad_dtype, beta=beta,
                  workspace=self._workspace(s, rows, g.device))
        done = getattr(sink, "done", None)
        if done is not None:
            done(self.name)
        if not need_input_grad:
            return None
        impl = self.impl_for(s, "dgrad")
        w = layer.packed("io", self.grad_dtype) if impl == "tc" else layer.packed("oi", torch.float32)
        gx = ops.conv(impl, g, w, s.dgrad_taps(), 1, 1, rows, s.cout, 1, rows, s.cin, add0=g if residual else None,
                      act_dtype=self.grad_dtype)
        return gx.view(*g.shape[:-1], s.cin)


class DisentangleMidEngine(ConvStack):
    """The middle of FourCamerasDisentanglement (pytorch/CNNs.py:257-279,288-313): four 1x1 convolutions whose channel
    counts (300, 400, 1600) do not tile the tensor-core kernels -- activations are STORED with zero channel padding
    (300 -> 320, 400 -> 448; fp32 mode stores them unpadded for the CUDA-core kernels), the packed weights and the
    bias carry matching zero rows / entries, so the padding stays exactly zero through every layer."""

    NAMES = ("rearrange_layer_1", "fusion_layer_1", "fusion_layer_2", "rearrange_layer_2")

    def __init__(self, module: nn.Module, precision: str):
        if precision == "fp16":
            raise ValueError("FourCamerasDisentanglement runs in 'bf16' or 'fp32'")
        super().__init__(precision)
        self.m = module
        for name in self.NAMES:
            conv = getattr(module, name)
            self.layers[name] = Layer(name, conv, Contraction("linear", conv.in_channels, conv.out_channels))
        self._bias_pad: Dict[str, torch.Tensor] = {}

    def stored(self, channels: int) -> int:
        """stored width of a `channels`-wide activation."""
        if self.precision == "fp32":
            return channels
        return {300: 320, 400: 448}.get(channels, (channels + 63) // 64 * 64)

    def _impl(self, what: str) -> str:
        return "simt" if self.precision == "fp32" or not tc_globally_enabled() else "tc"

    def _bias(self, name: str, cout_s: int) -> torch.Tensor:
        from . import vit_ops
        bias = self.layers[name].module.bias
        if cout_s == bias.numel():
            return bias
        buf = self._bias_pad.get(name)
        if buf is None or buf.device != bias.device or buf.numel() != cout_s:
            buf = torch.zeros(cout_s, device=bias.device, dtype=torch.float32)
            self._bias_pad[name] = buf
        vit_ops.colblock(bias.detach(), buf, rows=1, ncols=bias.numel(), src_row_stride=bias.numel(), dst_row_stride=cout_s)
        return buf

    def fwd(self, name: str, x: torch.Tensor, n: int, h: int, w: int, *, add1=None) -> torch.Tensor:
        """x [n, h, w, cin_stored] -> conv1x1 + bias (+ add1) [n, h, w, stored(cout)]."""
        layer = self.layers[name]
        s = layer.spec
        cin_s, cout_s = int(x.shape[-1]), self.stored(s.cout)
        impl = self._impl("fwd")
        if impl == "tc":
            wp = layer.packed("oi", torch.bfloat16, ipad=cout_s, jpad=cin_s)
        else:
            wp = layer.packed("io", torch.float32)
        return ops.conv(impl, x, wp, s.fwd_taps(), n, h, w, cin_s, h, w, cout_s, bias=self._bias(name, cout_s),
                        act=PB_ACT_NONE, add1=add1, act_dtype=self.act_dtype)

    def dgrad(self, name: str, g: torch.Tensor, n: int, h: int, w: int, cin_s: int, *, add0=None) -> torch.Tensor:
        """g [n, h, w, stored(cout)] -> gradient w.r.t. the layer input [n, h, w, cin_s] (+ add0)."""
        layer = self.layers[name]
        s = layer.spec
        impl = self._impl("dgrad")
        if impl == "tc":
            wp = layer.packed("io", torch.bfloat16, ipad=cin_s, jpad=int(g.shape[-1]))
        else:
            wp = layer.packed("oi", torch.float32)
        return ops.conv(impl, g, wp, s.dgrad_taps(), n, h, w, int(g.shape[-1]), h, w, cin_s, add0=add0,
                        act_dtype=self.act_dtype)

    def wgrad(self, name: str, a_in: torch.Tensor, g: torch.Tensor, n: int, h: int, w: int, sink: GradSink) -> None:
        layer = self.layers[name]
        s = layer.spec
        dw, db, beta = sink(name)
        ops.wgrad(self._impl("wgrad"), s, a_in, g, n, h, w, dw, db, act_dtype=self.act_dtype, beta=beta,
                  workspace=self._workspace(s, n * h * w, g.device))
        done = getattr(sink, "done", None)
        if done is not None:
            done(name)
