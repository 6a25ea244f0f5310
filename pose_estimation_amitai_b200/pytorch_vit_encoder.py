"""Drop-in ViT encoder (pytorch/pytorch_vit_encoder.py): same classes, constructor signatures,
sub-module names and state_dict keys (``patch_to_embedding``, ``norm``, ``pos_embedding``,
``cls_token``, ``transformer.layers.{l}.0.{norm,to_qkv,to_out.0}``, ``transformer.layers.{l}.1.net.{0,1,4}``,
``transformer.norm``) and the same creation order, so the same seed draws the same init.

The sub-modules are parameter containers; ``CustomViT.forward`` runs the whole encoder through
the sm_100a kernels (vit_engine.VitEncoderEngine).  FeedForward / Attention / Transformer are
not individually callable -- the reference only ever calls them from CustomViT.forward.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
from torch import nn


def pair(t):
    return t if isinstance(t, tuple) else (t, t)


class _PartOfCustomViT(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - guard
        raise NotImplementedError(f"{type(self).__name__} executes as part of CustomViT.forward on the B200 hot path")


class FeedForward(_PartOfCustomViT):
    """pytorch_vit_encoder.py:12-28: LN -> Linear -> GELU -> Dropout(0) -> Linear -> Dropout(0)."""

    def __init__(self, dim, hidden_dim, dropout=0.):
        super().__init__()
        self.net = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, hidden_dim), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(hidden_dim, dim), nn.Dropout(dropout))


class Attention(_PartOfCustomViT):
    """pytorch_vit_encoder.py:31-78."""

    def __init__(self, dim, heads=8, dim_head=64, dropout=0., out_dim=None):
        super().__init__()
        if out_dim is None:
            out_dim = dim
        inner_dim = dim_head * heads
        project_out = not (heads == 1 and dim_head == dim)
        if not project_out:
            raise NotImplementedError("heads == 1 and dim_head == dim (no output projection) is not on the hot path")
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.norm = nn.LayerNorm(dim)
        self.attend = nn.Softmax(dim=-1)
        self.dropout = nn.Dropout(dropout)
        self.to_qkv = nn.Linear(dim, inner_dim * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, out_dim), nn.Dropout(dropout))


class Transformer(_PartOfCustomViT):
    """pytorch_vit_encoder.py:81-105 (final norm created BEFORE the layers, as in the reference)."""

    def __init__(self, dim, depth, heads, dim_head, mlp_dim, dropout=0.):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            self.layers.append(nn.ModuleList([Attention(dim, heads=heads, dim_head=dim_head, dropout=dropout),
                                              FeedForward(dim, mlp_dim, dropout=dropout)]))


def _named_live_params(module: nn.Module) -> List[Tuple[str, nn.Parameter]]:
    return [(n, p) for n, p in module.named_parameters() if n != "cls_token"]


class _VitEncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, need, img, *params):
        tokens, saved = module._engine().forward(img.contiguous().float(), save=need)
        ctx.module, ctx.saved = module, saved
        b = img.shape[0]
        return tokens.view(b, tokens.shape[0] // b, tokens.shape[1])

    @staticmethod
    def backward(ctx, g):
        module = ctx.module
        eng = module._engine()
        names = [n for n, _ in _named_live_params(module)]
        store: Dict[str, torch.Tensor] = {}

        def sink(name: str, p: torch.Tensor):
            store[name] = torch.empty_like(p)
            return store[name], 0.0

        eng.backward(ctx.saved, g.reshape(-1, g.shape[-1]).contiguous().to(eng.act_dtype), sink)
        ctx.saved = None
        return (None, None, None, *[store.get(n) for n in names])


class CustomViT(nn.Module):
    """pytorch_vit_encoder.py:107-149.  forward: (B,C,H,W) -> (B, num_patches, dim)."""

    def __init__(self, *, image_size, patch_size, dim, depth, heads, mlp_dim, num_image_channels=4, dim_head=64,
                 dropout=0., emb_dropout=0., precision: str = "bf16"):
        super().__init__()
        image_height, image_width = pair(image_size)
        patch_height, patch_width = pair(patch_size)
        assert image_height % patch_height == 0 and image_width % patch_width == 0, \
            'Image dimensions must be divisible by the patch size.'
        if dropout != 0. or emb_dropout != 0.:
            raise NotImplementedError("dropout > 0 is never used by the reference (VITs.py:213-219 passes none)")
        num_patches = (image_height // patch_height) * (image_width // patch_width)
        patch_dim = num_image_channels * patch_height * patch_width
        self.patch_size = patch_size
        self.dim = dim
        self.patch_dim = patch_dim
        self.precision = precision
        self.patch_to_embedding = nn.Linear(patch_dim, dim)
        self.norm = nn.LayerNorm(dim)
        self.pos_embedding = nn.Parameter(torch.randn(1, num_patches, dim))
        self.cls_token = nn.Parameter(torch.randn(1, 1, dim))  # never used by the forward (reference :126)
        self.dropout = nn.Dropout(emb_dropout)
        self.transformer = Transformer(dim, depth, heads, dim_head, mlp_dim, dropout)

    def _engine(self):
        from .vit_engine import VitEncoderEngine
        eng = self.__dict__.get("_eng")
        if eng is None or eng.precision != self.precision:
            eng = VitEncoderEngine(self, self.precision)
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

    def forward(self, img):
        if not img.is_cuda:
            raise RuntimeError(f"CustomViT: input is on {img.device}; the B200 hot path has no CPU fallback")
        params = [p for _, p in _named_live_params(self)]
        need = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        return _VitEncoderFn.apply(self, need, img, *params)
