"""Model factory with the reference's surface (pytorch/Network.py:7-36):
``Network(config, image_size, num_output_channels).get_model()``."""
from __future__ import annotations

import torch

from .constants import *  # noqa: F401,F403  (np comes from here, as in the reference)
from . import CNNs


class Network:
    def __init__(self, config, image_size, num_output_channels):
        self.config = config
        self.image_size = np.array(image_size)
        self.num_output_channels = num_output_channels
        self.model_type = self.config['model type']
        self.model = self.config_model()

    def config_model(self):
        """pytorch/Network.py:15-26: every model type the reference dispatches on."""
        if self.model_type in (MODEL_18_POINTS_PER_WING, MODEL_18_POINTS_3_GOOD_CAMERAS, ALL_POINTS_MODEL):
            return CNNs.BasicNet(self.config, self.image_size, self.num_output_channels)
        if self.model_type == MODEL_18_POINTS_PER_WING_VIT:
            from . import VITs
            return VITs.VIT_encoder_CNN_decoder(self.config, self.image_size, self.num_output_channels)
        if self.model_type == ALL_CAMS_18_POINTS:
            return CNNs.FourCamerasBaseLine(self.config, self.image_size, self.num_output_channels)
        if self.model_type == ALL_CAMS_18_POINTS_VIT:
            from . import VITs
            return VITs.VIT4CamerasBaseLine(self.config, self.image_size, self.num_output_channels)
        if self.model_type == ALL_CAMS_DISENTANGLED_PER_WING_CNN:
            return CNNs.FourCamerasDisentanglement(self.config, self.image_size, self.num_output_channels)
        raise ValueError(f"model type {self.model_type!r} is not one Network.config_model dispatches on "
                         "(pytorch/Network.py:15-26)")

    def get_model(self):
        """pytorch/Network.py:28-36 moves the model to cuda-if-available and prints a torchsummary
        table; here a CUDA device is required (no CPU fallback) and the summary is a one-liner."""
        if not torch.cuda.is_available():
            raise RuntimeError("Network.get_model: no CUDA device -- the B200 hot path has no CPU fallback")
        self.model.to(torch.device("cuda"))
        n_params = sum(p.numel() for p in self.model.parameters())
        print(f"{type(self.model).__name__}: {n_params} parameters, precision {self.model.precision}", flush=True)
        return self.model
