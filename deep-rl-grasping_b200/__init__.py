"""b200grasp -- B200-native SAC learner behind the stable-baselines model API used by
BarisYazici/deep-rl-grasping (manipulation_main/training/sb_helper.py:104-128,175).

Import as ``b200grasp`` (``b200grasp.py`` at the repo root aliases this directory, whose name
``deep-rl-grasping_b200`` is not a Python identifier).
"""
from . import _lib, sb_io, synth  # noqa: F401
from . import callbacks, encoders, h5min, spaces, vec_env  # noqa: F401
from .learner import Learner  # noqa: F401
from .bdq import BDQ, BDQLearner  # noqa: F401
from .sac import SAC, CnnPolicy, MlpPolicy  # noqa: F401
from .vec_env import DummyVecEnv, VecNormalize  # noqa: F401

__all__ = ["BDQ", "BDQLearner", "SAC", "CnnPolicy", "MlpPolicy", "Learner", "DummyVecEnv", "VecNormalize", "callbacks", "encoders", "h5min", "spaces", "sb_io", "synth", "vec_env"]
