"""Model-type / mode tags shared with the reference (pytorch/constants.py).

This is drop-in interface surface: every tag is a string equal to its own name, the values a reference
``train_config.json`` carries in "model type".  ``np`` is re-exported because the reference's ``Network.py`` gets
it through ``from constants import *`` (Network.py:1,10) and ours mirrors that import.
"""
import numpy as np

# single-view models: Network.config_model -> BasicNet / VIT_encoder_CNN_decoder (Network.py:16-20)
MODEL_18_POINTS_PER_WING = "MODEL_18_POINTS_PER_WING"
MODEL_18_POINTS_3_GOOD_CAMERAS = "MODEL_18_POINTS_3_GOOD_CAMERAS"
ALL_POINTS_MODEL = "ALL_POINTS_MODEL"
MODEL_18_POINTS_PER_WING_VIT = "MODEL_18_POINTS_PER_WING_VIT"

# four-camera models (Network.py:21-26)
ALL_CAMS_18_POINTS = "ALL_CAMS_18_POINTS"
ALL_CAMS_DISENTANGLED_PER_WING_CNN = "ALL_CAMS_DISENTANGLED_PER_WING_CNN"
ALL_CAMS_18_POINTS_VIT = "ALL_CAMS_18_POINTS_VIT"

# tags only the reference's preprocessor / TensorFlow twin dispatch on (kept so configs and imports resolve)
PER_WING_MODEL = "PER_WING_MODEL"
TRAIN_ON_3_GOOD_CAMERAS_MODEL = "TRAIN_ON_3_GOOD_CAMERAS_MODEL"
ALL_CAMS_DISENTANGLED_PER_WING_VIT = "ALL_CAMS_DISENTANGLED_PER_WING_VIT"
PRETRAINED_LEAP = "PRETRAINED_LEAP"
GPTNET = "GPTNET"
ALL_POINTS_MODEL_VIT = "ALL_POINTS_MODEL_VIT"
ALL_CAMS = "ALL_CAMS"
MEAN_SQUARE_ERROR = "MEAN_SQUARE_ERROR"
MOVIE_TRAIN_SET = "MOVIE_TRAIN_SET"
RANDOM_TRAIN_SET = "RANDOM_TRAIN_SET"

# wing-point index ranges of the 14-point layout (left wing 0..6, right wing 7..13)
LEFT_INDEXES = np.arange(0, 7)
RIGHT_INDEXES = np.arange(7, 14)
