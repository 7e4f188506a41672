from ..evaluation import evaluate_policy  # noqa: F401
