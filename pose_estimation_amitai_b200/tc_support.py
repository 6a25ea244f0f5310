"""Which contractions the tcgen05 kernels (csrc/tc_conv.cu, csrc/tc_wgrad.cu) tile.  Everything
else runs on the CUDA-core gather-convolution (csrc/simt_conv.cu) -- still on the GPU."""
from __future__ import annotations

from .ops import Contraction

ENABLED = {"fwd": True, "dgrad": True, "wgrad": True}


def pad_n(n: int) -> int:
    """UMMA N must be a multiple of 16 for M=128."""
    return (n + 15) // 16 * 16


def _fwd_ok(spec: Contraction) -> bool:
    if spec.cin % 64 != 0:
        return False
    cp = pad_n(spec.cout)
    max_tile = 128 if spec.kind == "convT2" else 256
    return cp <= max_tile or cp % 32 == 0


def _dgrad_ok(spec: Contraction) -> bool:
    # a gather-conv with the channel roles swapped: K = cout (as stored), N = cin
    if spec.cin % 16 != 0:
        return False
    return spec.cin <= 256 or spec.cin % 32 == 0


def _wgrad_ok(spec: Contraction) -> bool:
    if spec.cout > 256 and spec.cout % 128 != 0:   # wide layers: column chunks of 256 (or 128)
        return False
    if spec.cin == 64:
        return spec.kind != "convT2"  # tap pairs share one gradient tile: needs tap-invariant g offsets
    return spec.cin % 128 == 0


def supported(spec: Contraction, what: str) -> bool:
    if not ENABLED.get(what, False):
        return False
    if spec.kind != "linear" and spec.ksize != 3:
        return False
    if what == "fwd":
        return _fwd_ok(spec)
    if spec.cout % 8 != 0:
        # the gradient tensor of this layer is channel-padded for TMA (16-byte rows); dgrad and wgrad
        # must then both take the padded layout
        return _dgrad_ok(spec) and _wgrad_ok(spec)
    return _dgrad_ok(spec) if what == "dgrad" else _wgrad_ok(spec)


def pool_fusable(spec: Contraction, oh: int, ow: int) -> bool:
    """the halo kernel's staged epilogue can emit lrelu(maxpool2x2(out)) next to (or instead of) `out`:
    stride-1 3x3 layers whose Cout tiles into 64-channel blocks, even output size (csrc/tc_conv2.cu, `pool`)."""
    import os
    return (spec.kind in ("conv", "convT1") and spec.cout % 64 == 0 and spec.cout <= 256 and spec.cin % 8 == 0
            and oh % 2 == 0 and ow % 2 == 0 and oh >= 16 and ow >= 8
            and os.environ.get("POSEB200_NO_POOL_FUSION", "0") != "1")
