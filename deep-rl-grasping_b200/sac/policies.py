"""Policy sentinels of ``stable_baselines.sac.policies`` (sb_helper.py:15-17, train_stable_baselines.py:17).  The network
the CUDA path implements is selected by the sentinel + ``policy_kwargs`` exactly as sb_helper.py:104-128 passes them;
layer-norm variants are not built and fail when a model is constructed with them."""
from ..sac_model import CnnPolicy, MlpPolicy  # noqa: F401


class LnMlpPolicy:
    unsupported = "layer-normalised SAC policies are not built (sb_helper.py never selects them for the shipped configs)"


class LnCnnPolicy(LnMlpPolicy):
    pass
