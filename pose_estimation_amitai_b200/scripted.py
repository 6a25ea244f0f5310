"""torch.library registration of the hot path and the TorchScript form of a trained model.

The reference keeps its best model as ``torch.jit.script(self.model).save('best_model.pth')``
(pytorch/train_pytorch.py:177-181) and downstream code loads it with ``torch.jit.load``.  A module whose forward is a
chain of C-ABI launches has no TorchScript body of its own, so the body is ONE registered operator:

    poseb200::heatmaps(Tensor x, Tensor flat_weights, str spec) -> Tensor      [B,Cin,H,W] fp32 -> [B,C,H,W] fp32
    poseb200::peaks(Tensor x, Tensor flat_weights, str spec) -> Tensor         ... -> [B,C,2] (x, y) arg-max peaks

``spec`` is a JSON string (model config, image size, output channels, and where each state_dict tensor lives inside
``flat_weights``); the CUDA implementation rebuilds the drop-in module from it once per (weights, spec) and runs the
same sm_100a kernels as ``model(x)``.  ``ScriptedPoseNet`` is the scriptable nn.Module that holds ``flat_weights`` as
a buffer and calls the operator; ``save`` scripts it and writes the archive.  ``torch.jit.load(path)`` works in any
process that has imported this package (the import registers the operators) -- there is no CPU kernel behind them.
"""
from __future__ import annotations

import json
from typing import Dict, Tuple

import numpy as np
import torch
from torch import nn

_LIB = torch.library.Library("poseb200", "DEF")
_LIB.define("heatmaps(Tensor x, Tensor flat_weights, str spec) -> Tensor")
_LIB.define("peaks(Tensor x, Tensor flat_weights, str spec) -> Tensor")

_MODELS: Dict[Tuple[int, int, str], nn.Module] = {}


def _model_for(flat: torch.Tensor, spec: str) -> nn.Module:
    key = (flat.data_ptr(), flat._version, spec)
    model = _MODELS.get(key)
    if model is None:
        from . import Network
        meta = json.loads(spec)
        net = Network.Network(meta["config"], np.array(meta["image_size"]), int(meta["num_output_channels"]))
        model = net.model
        sd = {}
        for name, off, shape, dtype in meta["tensors"]:
            n = int(np.prod(shape)) if shape else 1
            t = flat[off:off + n].view(shape)
            sd[name] = t.to(getattr(torch, dtype)) if dtype != "float32" else t
        model.load_state_dict(sd, strict=True)
        model = model.to(flat.device).eval()
        if len(_MODELS) >= 8:
            _MODELS.clear()
        _MODELS[key] = model
    return model


def _heatmaps_cuda(x: torch.Tensor, flat_weights: torch.Tensor, spec: str) -> torch.Tensor:
    with torch.no_grad():
        return _model_for(flat_weights, spec)(x)


def _peaks_cuda(x: torch.Tensor, flat_weights: torch.Tensor, spec: str) -> torch.Tensor:
    model = _model_for(flat_weights, spec)
    with torch.no_grad():
        if hasattr(model, "predict_peaks"):
            return model.predict_peaks(x)
        from . import ops
        return ops.peaks_argmax(model(x).contiguous())


def _no_cpu(*_args):
    raise RuntimeError("poseb200 operators run on CUDA tensors only: the B200 hot path has no CPU fallback")


_LIB.impl("heatmaps", _heatmaps_cuda, "CUDA")
_LIB.impl("peaks", _peaks_cuda, "CUDA")
_LIB.impl("heatmaps", _no_cpu, "CPU")
_LIB.impl("peaks", _no_cpu, "CPU")


def _meta_out(x: torch.Tensor, spec: str, peaks: bool) -> torch.Tensor:
    c = int(json.loads(spec)["num_output_channels"])
    return x.new_empty((x.shape[0], c, 2) if peaks else (x.shape[0], c, x.shape[2], x.shape[3]), dtype=torch.float32)


_LIB.impl("heatmaps", lambda x, w, spec: _meta_out(x, spec, False), "Meta")
_LIB.impl("peaks", lambda x, w, spec: _meta_out(x, spec, True), "Meta")


class ScriptedPoseNet(nn.Module):
    """``forward(x)`` = the wrapped model's heatmaps, ``peaks(x)`` its arg-max keypoints; scriptable."""

    def __init__(self, model: nn.Module):
        super().__init__()
        tensors, chunks, off = [], [], 0
        for name, t in model.state_dict().items():
            flat = t.detach().reshape(-1).to(torch.float32)
            tensors.append((name, off, list(t.shape), str(t.dtype).replace("torch.", "")))
            chunks.append(flat)
            off += flat.numel()
        dev = next(model.parameters()).device
        self.register_buffer("flat_weights", torch.cat(chunks).to(dev))
        config = {k: v for k, v in dict(model.config).items() if isinstance(v, (int, float, str, list, bool))}
        config["precision"] = getattr(model, "precision", config.get("precision", "bf16"))
        self.spec: str = json.dumps({"config": config, "image_size": [int(v) for v in model.image_size],
                                     "num_output_channels": int(model.number_of_output_channels),
                                     "tensors": tensors})

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return torch.ops.poseb200.heatmaps(x, self.flat_weights, self.spec)

    @torch.jit.export
    def peaks(self, x: torch.Tensor) -> torch.Tensor:
        return torch.ops.poseb200.peaks(x, self.flat_weights, self.spec)


def script(model: nn.Module) -> torch.jit.ScriptModule:
    return torch.jit.script(ScriptedPoseNet(model))


def save(model: nn.Module, path: str) -> None:
    """what ``torch.jit.script(self.model).save(path)`` is in the reference (pytorch/train_pytorch.py:179-180)."""
    script(model).save(path)
