"""pose_estimation_amitai_b200 -- the B200-native hot path of lior-kotlar/pose-estimation-amitai.

(The directory is spelled with underscores so it is importable; DESIGN.md uses the same name.)
Sub-modules mirror the reference's pytorch/ files: CNNs, VITs, pytorch_vit_encoder, Network,
constants, utils (peaks), train_pytorch (Trainer entry point).
"""
from . import _lib  # noqa: F401
from . import scripted  # noqa: F401  (registers the poseb200::heatmaps / ::peaks operators torch.jit.load needs)

__all__ = ["_lib", "ops", "CNNs", "Network", "constants"]
