"""``stable_baselines.common`` namespace as the reference imports it (sb_helper.py:11-21, train_stable_baselines.py:13-15,
base_callbacks.py:11-14): thin re-exports of the host modules of this package."""
import random

import numpy as np

from . import callbacks, evaluation, noise, policies, vec_env  # noqa: F401


def set_global_seeds(seed):
    """[SB2] common/misc_util.py set_global_seeds without the TensorFlow part (sb_helper.py:13).  The learner's own
    streams (replay indices, policy noise) are seeded through the model's ``seed=`` argument."""
    if seed is None:
        return
    np.random.seed(seed)
    random.seed(seed)
