from ..callbacks import (BaseCallback, CallbackList, CheckpointCallback, EvalCallback, EventCallback,  # noqa: F401
                         EveryNTimesteps, as_callback)
