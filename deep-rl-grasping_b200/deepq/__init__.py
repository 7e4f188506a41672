"""Import-compatibility namespace for ``stable_baselines.deepq`` (sb_helper.py:12): the DQN branch of ``SBPolicy.learn`` is
outside the scope table; the names import, constructing a model with them raises."""
from . import policies  # noqa: F401
