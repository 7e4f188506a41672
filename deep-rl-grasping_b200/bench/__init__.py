"""``stable_baselines.bench`` namespace: ``Monitor`` (train_stable_baselines.py:18,54)."""
from .monitor import Monitor, load_results  # noqa: F401
