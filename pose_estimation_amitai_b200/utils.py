"""Peak extraction with the reference's call surface (pytorch/utils.py, pytorch/Augmentor.py:105-148).

``find_points(confmaps)`` in the reference trainer (pytorch/train_pytorch.py:327-331) takes a numpy
(N,H,W,C) array that was copied off the GPU and transposed on the host.  These functions accept
that, but also a CUDA tensor in either layout -- in which case nothing leaves the device except the
(N,C,2) result.  Both run the sm_100a kernels; there is no CPU implementation here.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _to_cuda_nhwc(x) -> torch.Tensor:
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if not torch.cuda.is_available():
        raise RuntimeError("peak extraction runs on the GPU only (no CPU fallback)")
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    return x.cuda(non_blocking=True)


def torch_find_peaks_argmax(x, return_numpy: bool = True):
    """the function pytorch/train_pytorch.py:22,330 imports (its reference definition is the torch
    code at Augmentor.py:105-148): (N,H,W,C) -> (N,C,2) [x=col, y=row] float32."""
    peaks = ops.peaks_argmax(_to_cuda_nhwc(x), layout="nhwc")
    return peaks.cpu().numpy() if return_numpy else peaks


tf_find_peaks_argmax = torch_find_peaks_argmax  # pytorch/utils.py:6-44 computes the same thing with TensorFlow


def tf_find_peaks(x, return_numpy: bool = False):
    """Augmentor.tf_find_peaks (pytorch/Augmentor.py:105-148): (N,H,W,C) -> (N,C,2) [x, y] torch tensor."""
    return torch_find_peaks_argmax(x, return_numpy=return_numpy)


def tf_find_peaks_with_values(x, return_numpy: bool = True):
    """Preprocessor.tf_find_peaks (pytorch/preprocessor.py:630-668): (N,H,W,C) -> (N,3,C) rows [x, y, value]."""
    peaks, vals = ops.peaks_argmax(_to_cuda_nhwc(x), layout="nhwc", want_values=True)
    out = torch.stack([peaks[..., 0], peaks[..., 1], vals.to(torch.float32)], dim=1)
    return out.cpu().numpy() if return_numpy else out


def find_peaks_soft_argmax(x, return_numpy: bool = True):
    """pytorch/utils.py:47-83: intensity centroid, (N,H,W,C) -> (N,C,2) [x, y]."""
    peaks = ops.peaks_softargmax(_to_cuda_nhwc(x), layout="nhwc")
    return peaks.cpu().numpy() if return_numpy else peaks


def peaks_from_heatmaps_nchw(heatmaps: torch.Tensor, soft: bool = False) -> torch.Tensor:
    """device-resident variant for network outputs (B,C,H,W): no transpose, no D2H of heatmaps
    (replaces Trainer.get_points_from_confmaps, pytorch/train_pytorch.py:207-213)."""
    return ops.peaks_softargmax(heatmaps) if soft else ops.peaks_argmax(heatmaps)


def render_gaussian_targets(points_xy: torch.Tensor, sigma: float = 3.0, size=(192, 192)) -> torch.Tensor:
    """SimpleDataGenerator.get_gaussian / ensure_sigma (tensorflow/simple_data_generator.py:119-136) on device."""
    return ops.gaussian_heatmaps(points_xy, sigma=sigma, size=size)
