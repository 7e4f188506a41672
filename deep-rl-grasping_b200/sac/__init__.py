"""``stable_baselines.sac`` namespace: ``from b200grasp.sac.policies import MlpPolicy, CnnPolicy`` (sb_helper.py:15-17)."""
from ..sac_model import SAC, CnnPolicy, MlpPolicy, unwrap_vec_normalize  # noqa: F401
from . import policies  # noqa: F401
