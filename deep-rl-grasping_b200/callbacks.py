"""Callback protocol the learner drives ([SB2] common/callbacks.py as used by
/root/reference/manipulation_main/training/base_callbacks.py:16-245 and sb_helper.py:25-54):
init_callback(model) -> on_training_start(locals, globals) -> {on_rollout_start, on_step -> bool,
on_rollout_end}* -> on_training_end; attributes n_calls, num_timesteps, model, training_env,
locals, globals, parent."""
from __future__ import annotations

from typing import List, Optional


class BaseCallback:
    def __init__(self, verbose: int = 0):
        self.model = None
        self.training_env = None
        self.n_calls = 0
        self.num_timesteps = 0
        self.verbose = verbose
        self.locals = None
        self.globals = None
        self.logger = None
        self.parent = None

    def init_callback(self, model) -> None:
        self.model = model
        self.training_env = model.get_env()
        self._init_callback()

    def _init_callback(self) -> None:
        pass

    def on_training_start(self, locals_, globals_) -> None:
        self.locals, self.globals = locals_, globals_
        self._on_training_start()

    def _on_training_start(self) -> None:
        pass

    def on_rollout_start(self) -> None:
        self._on_rollout_start()

    def _on_rollout_start(self) -> None:
        pass

    def _on_step(self) -> bool:
        return True

    def on_step(self) -> bool:
        self.n_calls += 1
        self.num_timesteps = self.model.num_timesteps
        return self._on_step()

    def on_rollout_end(self) -> None:
        self._on_rollout_end()

    def _on_rollout_end(self) -> None:
        pass

    def on_training_end(self) -> None:
        self._on_training_end()

    def _on_training_end(self) -> None:
        pass


class EventCallback(BaseCallback):
    def __init__(self, callback: Optional[BaseCallback] = None, verbose: int = 0):
        super().__init__(verbose)
        self.callback = callback
        if callback is not None:
            callback.parent = self

    def init_callback(self, model) -> None:
        super().init_callback(model)
        if self.callback is not None:
            self.callback.init_callback(model)

    def _on_training_start(self) -> None:
        if self.callback is not None:
            self.callback.on_training_start(self.locals, self.globals)

    def _on_event(self) -> bool:
        return self.callback.on_step() if self.callback is not None else True


class CallbackList(BaseCallback):
    def __init__(self, callbacks: List[BaseCallback]):
        super().__init__()
        self.callbacks = callbacks

    def _init_callback(self) -> None:
        for c in self.callbacks:
            c.init_callback(self.model)

    def _on_training_start(self) -> None:
        for c in self.callbacks:
            c.on_training_start(self.locals, self.globals)

    def _on_rollout_start(self) -> None:
        for c in self.callbacks:
            c.on_rollout_start()

    def _on_step(self) -> bool:
        ok = True
        for c in self.callbacks:
            ok = c.on_step() and ok
        return ok

    def _on_rollout_end(self) -> None:
        for c in self.callbacks:
            c.on_rollout_end()

    def _on_training_end(self) -> None:
        for c in self.callbacks:
            c.on_training_end()


class _FnCallback(BaseCallback):
    """Legacy ``callback(locals, globals) -> bool`` functions."""

    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def _on_step(self) -> bool:
        r = self.fn(self.locals, self.globals)
        return True if r is None else bool(r)


def as_callback(cb) -> BaseCallback:
    if cb is None:
        return BaseCallback()
    if isinstance(cb, (list, tuple)):
        return CallbackList([as_callback(c) for c in cb])
    if callable(cb) and not hasattr(cb, "on_step"):
        return _FnCallback(cb)
    return cb
