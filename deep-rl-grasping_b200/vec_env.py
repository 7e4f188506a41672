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
from typing import Optional,  Callable, List, Sequence

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


def _subproc_worker(remote, parent_remote, fn_bytes):
    """One environment per process ([SB2] common/vec_env/subproc_vec_env.py protocol: step / reset / close / get_spaces /
    get_attr / env_method); auto-reset on done with the terminal observation in ``info``."""
    import pickle
    parent_remote.close()
    env = pickle.loads(fn_bytes)()
    try:
        while True:
            cmd, data = remote.recv()
            if cmd == "step":
                o, r, d, info = env.step(data)
                if d:
                    info = dict(info)
                    info["terminal_observation"] = np.asarray(o)
                    o = env.reset()
                remote.send((np.asarray(o), r, d, info))
            elif cmd == "reset":
                remote.send(np.asarray(env.reset()))
            elif cmd == "get_spaces":
                remote.send((env.observation_space, env.action_space))
            elif cmd == "get_attr":
                remote.send(getattr(env, data))
            elif cmd == "env_method":
                remote.send(getattr(env, data[0])(*data[1], **data[2]))
            elif cmd == "close":
                if hasattr(env, "close"):
                    env.close()
                remote.close()
                break
            else:
                raise NotImplementedError(cmd)
    except (EOFError, KeyboardInterrupt):
        pass


class SubprocVecEnv(VecEnv):
    """Vectorised environments in worker processes: the host-side actor loop that feeds the GPU-resident replay buffer
    (BASELINE config 5: 128 PyBullet envs on host cores; the reference imports it at sb_helper.py:19).  ``step_async``
    sends every action first, ``step_wait`` collects, so the environments advance in parallel on the host cores while
    the previous batch of transitions is already on its way to the device."""

    def __init__(self, env_fns: Sequence[Callable], start_method: Optional[str] = None):
        import multiprocessing as mp
        import cloudpickle
        ctx = mp.get_context(start_method or ("forkserver" if "forkserver" in mp.get_all_start_methods() else "spawn"))
        self.num_envs = len(env_fns)
        self.remotes, self.work_remotes = zip(*[ctx.Pipe() for _ in env_fns])
        self.procs = []
        for wr, r, fn in zip(self.work_remotes, self.remotes, env_fns):
            p = ctx.Process(target=_subproc_worker, args=(wr, r, cloudpickle.dumps(fn)), daemon=True)
            p.start()
            self.procs.append(p)
            wr.close()
        self.remotes[0].send(("get_spaces", None))
        self.observation_space, self.action_space = self.remotes[0].recv()
        self.buf_infos: List[dict] = [{} for _ in env_fns]
        self.waiting = self.closed = False

    def step_async(self, actions):
        for r, a in zip(self.remotes, actions):
            r.send(("step", a))
        self.waiting = True

    def step_wait(self):
        res = [r.recv() for r in self.remotes]
        self.waiting = False
        obs, rews, dones, infos = zip(*res)
        self.buf_infos = list(infos)
        return np.stack(obs), np.asarray(rews, np.float32), np.asarray(dones), list(infos)

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def reset(self):
        for r in self.remotes:
            r.send(("reset", None))
        return np.stack([r.recv() for r in self.remotes])

    def get_attr(self, name, indices=None):
        idx = range(self.num_envs) if indices is None else indices
        for i in idx:
            self.remotes[i].send(("get_attr", name))
        return [self.remotes[i].recv() for i in idx]

    def env_method(self, name, *a, **k):
        for r in self.remotes:
            r.send(("env_method", (name, a, k)))
        return [r.recv() for r in self.remotes]

    @property
    def envs(self):
        raise AttributeError("SubprocVecEnv has no in-process envs; use get_attr / env_method")

    def close(self):
        if self.closed:
            return
        if self.waiting:
            for r in self.remotes:
                r.recv()
        for r in self.remotes:
            r.send(("close", None))
        for p in self.procs:
            p.join(timeout=5)
        self.closed = True


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
        # The pickle must NAME stable_baselines' classes.  Stand-in classes carrying those module / class names are
        # registered in sys.modules only for the duration of the dump; whatever was there before (a real stable_baselines
        # install included) is put back in `finally`, also when the dump raises.
        vn_mod, rm_mod = "stable_baselines.common.vec_env.vec_normalize", "stable_baselines.common.running_mean_std"
        vn_cls = type("VecNormalize", (), {"__module__": vn_mod})
        rm_cls = type("RunningMeanStd", (), {"__module__": rm_mod})
        created, missing = [], object()
        saved = {}
        try:
            for name in ("stable_baselines", "stable_baselines.common", "stable_baselines.common.vec_env", vn_mod, rm_mod):
                if name not in sys.modules:
                    sys.modules[name] = types.ModuleType(name)
                    created.append(name)
            for mod, attr, cls in ((vn_mod, "VecNormalize", vn_cls), (rm_mod, "RunningMeanStd", rm_cls)):
                saved[(mod, attr)] = getattr(sys.modules[mod], attr, missing)
                setattr(sys.modules[mod], attr, cls)
            obj = vn_cls.__new__(vn_cls)
            st = self.__getstate__()
            for k in ("obs_rms", "ret_rms"):
                r = rm_cls.__new__(rm_cls)
                r.__dict__.update(st[k].__dict__)
                st[k] = r
            for k in ("observation_space", "action_space"):      # our Box is not importable on the reference side
                st.pop(k, None)
            obj.__dict__.update(st)
            with open(path, "wb") as f:
                pickle.dump(obj, f)
        finally:
            for (mod, attr), old in saved.items():
                if mod in sys.modules and mod not in created:
                    if old is missing:
                        try:
                            delattr(sys.modules[mod], attr)
                        except AttributeError:
                            pass
                    else:
                        setattr(sys.modules[mod], attr, old)
            for name in created:
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
