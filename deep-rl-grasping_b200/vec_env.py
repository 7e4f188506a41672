"""Minimal VecEnv / VecNormalize with the surface the reference harness touches
(sb_helper.py:75,101-103,118-119; train_stable_baselines.py:52-54,88-91; utils.py:71-76;
base_callbacks.py:140-149).  Restates [SB2] common/vec_env/{dummy_vec_env,vec_normalize}.py and
common/running_mean_std.py so that a reference user can switch imports without gym's VecEnv
machinery.  Statistics are float64 numpy, exactly what the learner consumes at sample time.
"""
from __future__ import annotations

import pickle
import sys
import types
from typing import Callable, List, Sequence

import numpy as np


class RunningMeanStd:
    """[SB2] common/running_mean_std.py (parallel-variance update, count starts at epsilon)."""

    def __init__(self, epsilon: float = 1e-4, shape=()):
        self.mean = np.zeros(shape, np.float64)
        self.var = np.ones(shape, np.float64)
        self.count = epsilon

    def update(self, arr: np.ndarray) -> None:
        arr = np.asarray(arr, np.float64)
        self.update_from_moments(arr.mean(axis=0), arr.var(axis=0), arr.shape[0])

    def update_from_moments(self, batch_mean, batch_var, batch_count: int) -> None:
        delta = batch_mean - self.mean
        tot = self.count + batch_count
        new_mean = self.mean + delta * batch_count / tot
        m_a = self.var * self.count
        m_b = batch_var * batch_count
        m_2 = m_a + m_b + np.square(delta) * self.count * batch_count / tot
        self.mean, self.var, self.count = new_mean, m_2 / tot, tot


class VecEnv:
    """Marker base class (``isinstance(env, VecEnv)`` in base_callbacks.py:50 and [SB2] evaluate_policy)."""
    num_envs = 1


class DummyVecEnv(VecEnv):
    """Sequential vectorised env: ``DummyVecEnv([lambda: env, ...])`` (train_stable_baselines.py:54)."""

    def __init__(self, env_fns: Sequence[Callable]):
        self.envs = [fn() for fn in env_fns]
        self.num_envs = len(self.envs)
        e = self.envs[0]
        self.observation_space, self.action_space = e.observation_space, e.action_space
        self.buf_infos: List[dict] = [{} for _ in self.envs]
        self._actions = None

    def reset(self):
        return np.stack([np.asarray(e.reset()) for e in self.envs])

    def step_async(self, actions):
        self._actions = actions

    def step_wait(self):
        obs, rews, dones = [], [], []
        for i, (e, a) in enumerate(zip(self.envs, self._actions)):
            o, r, d, info = e.step(a)
            if d:
                info = dict(info)
                info["terminal_observation"] = o
                o = e.reset()
            obs.append(np.asarray(o)); rews.append(r); dones.append(d)
            self.buf_infos[i] = info
        return np.stack(obs), np.asarray(rews, np.float32), np.asarray(dones), list(self.buf_infos)

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        for e in self.envs:
            if hasattr(e, "close"):
                e.close()

    def get_attr(self, name, indices=None):
        return [getattr(e, name) for e in self.envs]

    def env_method(self, name, *a, **k):
        return [getattr(e, name)(*a, **k) for e in self.envs]


class VecNormalize(VecEnv):
    """[SB2] VecNormalize(venv, training=True, norm_obs=True, norm_reward=True, clip_obs=10.,
    clip_reward=10., gamma=0.99, epsilon=1e-8)."""

    def __init__(self, venv, training=True, norm_obs=True, norm_reward=True, clip_obs=10.0, clip_reward=10.0, gamma=0.99,
                 epsilon=1e-8):
        self.venv = venv
        self.num_envs = venv.num_envs
        self.observation_space, self.action_space = venv.observation_space, venv.action_space
        self.obs_rms = RunningMeanStd(shape=self.observation_space.shape)
        self.ret_rms = RunningMeanStd(shape=())
        self.clip_obs, self.clip_reward = clip_obs, clip_reward
        self.ret = np.zeros(self.num_envs)
        self.gamma, self.epsilon = gamma, epsilon
        self.training, self.norm_obs, self.norm_reward = training, norm_obs, norm_reward
        self.old_obs, self.old_rews = np.array([]), np.array([])

    # ---- VecEnv surface
    @property
    def envs(self):
        return self.venv.envs

    @property
    def buf_infos(self):
        return self.venv.buf_infos

    def step_async(self, actions):
        self.venv.step_async(actions)

    def step_wait(self):
        obs, rews, news, infos = self.venv.step_wait()
        self.ret = self.ret * self.gamma + rews
        self.old_obs, self.old_rews = obs, rews
        if self.training:
            self.obs_rms.update(obs)
            self.ret_rms.update(self.ret)
        obs = self.normalize_obs(obs)
        rews = self.normalize_reward(rews)
        self.ret[news] = 0
        return obs, rews, news, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def reset(self):
        obs = self.venv.reset()
        self.old_obs = obs
        self.ret = np.zeros(self.num_envs)
        if self.training:
            self.obs_rms.update(obs)
        return self.normalize_obs(obs)

    def close(self):
        self.venv.close()

    def get_attr(self, name, indices=None):
        return self.venv.get_attr(name, indices)

    # ---- normalisation (float64 like numpy in SB2; the learner applies the same formula on the device)
    def normalize_obs(self, obs):
        if self.norm_obs:
            obs = np.clip((obs - self.obs_rms.mean) / np.sqrt(self.obs_rms.var + self.epsilon), -self.clip_obs, self.clip_obs)
        return obs

    def normalize_reward(self, reward):
        if self.norm_reward:
            reward = np.clip(reward / np.sqrt(self.ret_rms.var + self.epsilon), -self.clip_reward, self.clip_reward)
        return reward

    def get_original_obs(self):
        return self.old_obs.copy()

    def get_original_reward(self):
        return self.old_rews.copy()

    # ---- persistence (vecnormalize.pkl; sb_helper.py:246-247, train_stable_baselines.py:88-91)
    def __getstate__(self):
        st = self.__dict__.copy()
        for k in ("venv", "num_envs", "ret"):
            st.pop(k, None)
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)
        self.venv = None

    def set_venv(self, venv):
        self.venv = venv
        self.num_envs = venv.num_envs
        self.ret = np.zeros(self.num_envs)

    def save(self, path: str, sb_compatible: bool = True):
        """Pickle.  With ``sb_compatible`` the pickle names stable_baselines' own classes so that the
        reference's ``VecNormalize.load`` (train_stable_baselines.py:91) can read it back."""
        if not sb_compatible:
            with open(path, "wb") as f:
                pickle.dump(self, f)
            return
        mods = {}
        for name in ("stable_baselines", "stable_baselines.common", "stable_baselines.common.vec_env",
                     "stable_baselines.common.vec_env.vec_normalize", "stable_baselines.common.running_mean_std"):
            if name not in sys.modules:
                mods[name] = types.ModuleType(name)
        sys.modules.update(mods)
        try:
            vn_cls = type("VecNormalize", (), {"__module__": "stable_baselines.common.vec_env.vec_normalize"})
            rm_cls = type("RunningMeanStd", (), {"__module__": "stable_baselines.common.running_mean_std"})
            had = [getattr(sys.modules[vn_cls.__module__], "VecNormalize", None), getattr(sys.modules[rm_cls.__module__], "RunningMeanStd", None)]
            sys.modules[vn_cls.__module__].VecNormalize = vn_cls
            sys.modules[rm_cls.__module__].RunningMeanStd = rm_cls
            obj = vn_cls.__new__(vn_cls)
            st = self.__getstate__()
            for k in ("obs_rms", "ret_rms"):
                r = rm_cls.__new__(rm_cls)
                r.__dict__.update(st[k].__dict__)
                st[k] = r
            obj.__dict__.update(st)
            with open(path, "wb") as f:
                pickle.dump(obj, f)
            if had[0] is not None:
                sys.modules[vn_cls.__module__].VecNormalize = had[0]
            if had[1] is not None:
                sys.modules[rm_cls.__module__].RunningMeanStd = had[1]
        finally:
            for name in mods:
                sys.modules.pop(name, None)

    @staticmethod
    def load(path: str, venv):
        """Reads either our pickle or one written by stable-baselines (trained_models/*/vecnormalize.pkl)."""
        from .sb_io import load_vecnormalize
        d = load_vecnormalize(path)
        vn = VecNormalize(venv, training=True, norm_obs=d["norm_obs"], norm_reward=d["norm_reward"], clip_obs=d["clip_obs"],
                          clip_reward=d["clip_reward"], gamma=d["gamma"], epsilon=d["epsilon"])
        vn.obs_rms.mean, vn.obs_rms.var, vn.obs_rms.count = d["obs_mean"], d["obs_var"], d["obs_count"]
        vn.ret_rms.mean, vn.ret_rms.var, vn.ret_rms.count = np.float64(d["ret_mean"]), np.float64(d["ret_var"]), d["ret_count"]
        return vn


def sync_envs_normalization(env, eval_env) -> None:
    """[SB2] common/vec_env/__init__.py: copy the running statistics of every VecNormalize layer of ``env`` into the
    matching layer of ``eval_env`` (base_callbacks.py:81 before each evaluation)."""
    import copy
    e, ev = env, eval_env
    while e is not None and ev is not None:
        if isinstance(e, VecNormalize) and isinstance(ev, VecNormalize):
            ev.obs_rms = copy.deepcopy(e.obs_rms)
            ev.ret_rms = copy.deepcopy(e.ret_rms)
        e, ev = getattr(e, "venv", None), getattr(ev, "venv", None)
