"""Episode monitor: ``Monitor(gym.make('gripper-env-v0', config=config), os.path.join(model_dir, "log_file"))``
(train_stable_baselines.py:54).  Writes ``<filename>.monitor.csv`` in the layout of the files shipped under
/root/reference/trained_models/*/log_file.monitor.csv (they come from the author's stable-baselines fork, absent from the
tree): a ``#{"t_start": ..., "env_id": ...}`` JSON comment line, then the header ``r,s,l,c,timesteps,t`` and one row per
episode -- reward sum, success flag (``info['is_success']``, robot.py:181), length, curriculum lambda (0 .. 1 in steps of
1/8 in the shipped logs; read from ``env.curriculum`` / ``info['curriculum_lambda']`` when available, else 0), cumulative
timestep index (``total_steps - 1``: the first 150-step episode logs 149) and seconds since start.  Duck-typed wrapper:
no dependency on gym.
"""
from __future__ import annotations

import csv
import json
import os
import time
from typing import List

import numpy as np


class Monitor:
    EXT = "monitor.csv"
    FIELDS = ("r", "s", "l", "c", "timesteps", "t")

    def __init__(self, env, filename=None, allow_early_resets=True, reset_keywords=(), info_keywords=()):
        self.env = env
        self.observation_space = getattr(env, "observation_space", None)
        self.action_space = getattr(env, "action_space", None)
        self.metadata = getattr(env, "metadata", {})
        self.reward_range = getattr(env, "reward_range", (-float("inf"), float("inf")))
        self.t_start = time.time()
        self.file_handler = None
        self.logger = None
        if filename is not None:
            if not filename.endswith(Monitor.EXT):
                filename = filename + "." + Monitor.EXT if not os.path.isdir(filename) else os.path.join(filename, Monitor.EXT)
            self.file_handler = open(filename, "wt")
            env_id = getattr(getattr(env, "spec", None), "id", None)
            self.file_handler.write("#%s\n" % json.dumps({"t_start": self.t_start, "env_id": env_id}))
            self.logger = csv.DictWriter(self.file_handler, fieldnames=Monitor.FIELDS + tuple(reset_keywords) + tuple(info_keywords))
            self.logger.writeheader()
            self.file_handler.flush()
        self.reset_keywords, self.info_keywords = tuple(reset_keywords), tuple(info_keywords)
        self.allow_early_resets = allow_early_resets
        self.rewards: List[float] = []
        self.needs_reset = True
        self.episode_rewards: List[float] = []
        self.episode_lengths: List[int] = []
        self.episode_times: List[float] = []
        self.total_steps = 0
        self.current_reset_info = {}

    def __getattr__(self, name):          # everything else is the wrapped env's
        if name in ("env", "__setstate__"):
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return getattr(self.env, "unwrapped", self.env)

    def reset(self, **kwargs):
        if not self.allow_early_resets and not self.needs_reset:
            raise RuntimeError("Tried to reset an environment before done; wrap it with Monitor(env, path, allow_early_resets=True)")
        self.rewards = []
        self.needs_reset = False
        for k in self.reset_keywords:
            if k not in kwargs:
                raise ValueError("Expected you to pass kwarg %s into reset" % k)
            self.current_reset_info[k] = kwargs[k]
        return self.env.reset(**kwargs)

    def _curriculum(self, info):
        cur = getattr(self.unwrapped, "curriculum", None)
        for attr in ("_lambda", "lambda_", "current_lambda"):
            if cur is not None and hasattr(cur, attr):
                return float(getattr(cur, attr))
        return float(info.get("curriculum_lambda", 0.0))

    def step(self, action):
        if self.needs_reset:
            raise RuntimeError("Tried to step environment that needs reset")
        obs, rew, done, info = self.env.step(action)
        self.rewards.append(float(rew))
        self.total_steps += 1
        if done:
            self.needs_reset = True
            ep_rew, ep_len = float(sum(self.rewards)), len(self.rewards)
            ep = {"r": round(ep_rew, 6), "s": float(bool(info.get("is_success", False))), "l": ep_len, "c": self._curriculum(info),
                  "timesteps": self.total_steps - 1, "t": round(time.time() - self.t_start, 6)}
            for k in self.info_keywords:
                ep[k] = info[k]
            self.episode_rewards.append(ep_rew)
            self.episode_lengths.append(ep_len)
            self.episode_times.append(time.time() - self.t_start)
            ep.update(self.current_reset_info)
            if self.logger:
                self.logger.writerow(ep)
                self.file_handler.flush()
            info = dict(info)
            info["episode"] = ep
        return obs, rew, done, info

    def close(self):
        if self.file_handler is not None:
            self.file_handler.close()
            self.file_handler = None
        if hasattr(self.env, "close"):
            self.env.close()

    def get_total_steps(self):
        return self.total_steps

    def get_episode_rewards(self):
        return self.episode_rewards

    def get_episode_lengths(self):
        return self.episode_lengths

    def get_episode_times(self):
        return self.episode_times


def load_results(path):
    """Rows of every ``*monitor.csv`` under ``path`` as a list of dicts, ordered by absolute time (t_start + t)."""
    rows = []
    for fn in sorted(os.listdir(path)):
        if not fn.endswith(Monitor.EXT):
            continue
        with open(os.path.join(path, fn)) as f:
            head = f.readline()
            assert head[0] == "#"
            t0 = json.loads(head[1:])["t_start"]
            for r in csv.DictReader(f):
                r = {k: float(v) for k, v in r.items()}
                r["t"] += t0
                rows.append(r)
    rows.sort(key=lambda r: r["t"])
    return rows
