"""Model-type / mode tags shared with the reference (pytorch/constants.py).

Every tag of the reference is a string equal to its own name, so they are generated from the name lists below,
grouped by what dispatches on them; `np` is re-exported because the reference's ``Network.py`` gets it through
``from constants import *`` (Network.py:1,10) and ours mirrors that import.
"""
import numpy as np

_SINGLE_VIEW_MODELS = (            # Network.config_model -> BasicNet / VIT_encoder_CNN_decoder (Network.py:16-20)
    "MODEL_18_POINTS_PER_WING", "MODEL_18_POINTS_3_GOOD_CAMERAS", "ALL_POINTS_MODEL", "MODEL_18_POINTS_PER_WING_VIT")
_MULTI_CAMERA_MODELS = (           # Network.config_model -> the four-camera models (Network.py:21-26)
    "ALL_CAMS_18_POINTS", "ALL_CAMS_DISENTANGLED_PER_WING_CNN", "ALL_CAMS_18_POINTS_VIT")
_OTHER_TAGS = (                    # read by the reference's preprocessor / TensorFlow twin only
    "PER_WING_MODEL", "TRAIN_ON_3_GOOD_CAMERAS_MODEL", "ALL_CAMS_DISENTANGLED_PER_WING_VIT", "PRETRAINED_LEAP",
    "GPTNET", "ALL_POINTS_MODEL_VIT", "ALL_CAMS", "MEAN_SQUARE_ERROR", "MOVIE_TRAIN_SET", "RANDOM_TRAIN_SET")

globals().update({tag: tag for tag in _SINGLE_VIEW_MODELS + _MULTI_CAMERA_MODELS + _OTHER_TAGS})

# wing-point index ranges of the 14-point layout (left wing 0..6, right wing 7..13)
LEFT_INDEXES, RIGHT_INDEXES = np.split(np.arange(14), 2)

__all__ = ["np", "LEFT_INDEXES", "RIGHT_INDEXES", *_SINGLE_VIEW_MODELS, *_MULTI_CAMERA_MODELS, *_OTHER_TAGS]
