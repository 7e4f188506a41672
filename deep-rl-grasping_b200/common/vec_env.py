from ..vec_env import DummyVecEnv, RunningMeanStd, VecEnv, VecNormalize, sync_envs_normalization  # noqa: F401


class SubprocVecEnv:            # sb_helper.py:19 imports it, the shipped configs run DummyVecEnv (train_stable_baselines.py:54)
    def __init__(self, *a, **k):
        raise NotImplementedError("SubprocVecEnv is not built: the reference runs a single DummyVecEnv")


class VecFrameStack(SubprocVecEnv):
    def __init__(self, *a, **k):
        raise NotImplementedError("VecFrameStack is not built (imported but unused by the reference)")
