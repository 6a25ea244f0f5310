"""Which contractions the tcgen05 kernels (csrc/tc_*.cu) tile.  Everything else runs on the
CUDA-core gather-convolution (csrc/simt_conv.cu)."""
from __future__ import annotations

from .ops import Contraction

ENABLED = {"fwd": False, "dgrad": False, "wgrad": False}


def pad_n(n: int) -> int:
    """UMMA N must be a multiple of 16 for M=128 (guide: Guideline 10)."""
    return (n + 15) // 16 * 16


def supported(spec: Contraction, what: str) -> bool:
    if not ENABLED.get(what, False):
        return False
    if spec.ksize != 3 and spec.kind != "linear":
        return False
    return spec.cin % 64 == 0 and (spec.cout % 64 == 0 or spec.cout <= 64)
