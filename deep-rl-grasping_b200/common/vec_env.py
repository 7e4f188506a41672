from ..vec_env import DummyVecEnv, RunningMeanStd, SubprocVecEnv, VecEnv, VecNormalize, sync_envs_normalization  # noqa: F401


class VecFrameStack:            # imported by sb_helper.py:19, never constructed by the reference
    def __init__(self, *a, **k):
        raise NotImplementedError("VecFrameStack is not built (imported but unused by the reference)")
