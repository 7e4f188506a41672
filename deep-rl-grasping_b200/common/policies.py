"""``stable_baselines.common.policies`` names imported by sb_helper.py:11,18 for the TRPO/PPO branches (out of scope)."""


class MlpPolicy:
    unsupported = "actor-critic policies (TRPO / PPO branches) are outside the hot-path scope (DESIGN.md section 7)"


class CnnPolicy(MlpPolicy):
    pass
