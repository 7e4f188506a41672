"""``SAC`` -- the stable-baselines model object the reference constructs and drives
(/root/reference/manipulation_main/training/sb_helper.py:104-128,175,186-198,241-247;
train_stable_baselines.py:95-106; utils.py:71; base_callbacks.py:84-111,145-146), re-hosted on the
B200 learner.  Same constructor keywords, ``learn / predict / save / load / get_parameters /
load_parameters / get_env / get_vec_normalize_env``, same callback protocol, same zip format.

What runs where: env stepping (PyBullet) and this loop stay on host cores; replay storage,
minibatch sampling, VecNormalize-at-sample-time, the whole gradient step and the target update
run on the GPU behind the C ABI (include/b200grasp.h).  There is no CPU fallback.
"""
from __future__ import annotations

import os
import time
from collections import OrderedDict, deque
from typing import Optional

import numpy as np

from . import _lib, sb_io
from .callbacks import as_callback
from .learner import Learner
from .vec_env import DummyVecEnv, VecNormalize


class CnnPolicy:      # sentinels standing in for stable_baselines.sac.policies.{CnnPolicy,MlpPolicy}
    pass


class MlpPolicy:
    pass


_PRECISIONS = {"fp32": _lib.B2G_PREC_FP32_SIMT, "bf16x3": _lib.B2G_PREC_BF16X3, "bf16": _lib.B2G_PREC_BF16}


def _constfn(v):
    return v if callable(v) else (lambda _frac: float(v))


def _is_vec(env):
    return hasattr(env, "num_envs")


def unwrap_vec_normalize(env) -> Optional[VecNormalize]:
    e = env
    while e is not None:
        if isinstance(e, VecNormalize) or type(e).__name__ == "VecNormalize":
            return e
        e = getattr(e, "venv", None)
    return None


class SAC:
    def __init__(self, policy, env, gamma=0.99, learning_rate=3e-4, buffer_size=50000, learning_starts=100, train_freq=1,
                 batch_size=64, tau=0.005, ent_coef="auto", target_update_interval=1, gradient_steps=1,
                 target_entropy="auto", action_noise=None, random_exploration=0.0, verbose=0, tensorboard_log=None,
                 _init_setup_model=True, policy_kwargs=None, full_tensorboard_log=False, seed=None, n_cpu_tf_sess=None,
                 precision="bf16x3", device=0, rank=0, nranks=1, nccl_id=None):
        if ent_coef != "auto":
            raise NotImplementedError("only ent_coef='auto' (every shipped zip; SURVEY.md section 8c) is built")
        if target_update_interval != 1:
            raise NotImplementedError("target_update_interval must be 1 (every shipped zip)")
        if action_noise is not None:
            raise NotImplementedError("action_noise is not used by the reference (zip: action_noise None)")
        if getattr(policy, "unsupported", None):
            raise NotImplementedError(policy.unsupported)
        self.policy = policy
        self.policy_kwargs = dict(policy_kwargs or {})
        layers = list(self.policy_kwargs.get("layers", [64, 64]))
        if layers != [64, 64]:
            raise NotImplementedError("SAC.layers must be [64, 64] (config/gripper_grasp.yaml:81)")
        if self.policy_kwargs.get("layer_norm", False):
            raise NotImplementedError("layer_norm=True is not used by the reference (sb_helper.py:95)")
        self.gamma, self.tau = float(gamma), float(tau)
        self.learning_rate = learning_rate
        self.buffer_size, self.batch_size = int(buffer_size), int(batch_size)
        self.learning_starts, self.train_freq, self.gradient_steps = int(learning_starts), int(train_freq), int(gradient_steps)
        self.random_exploration = float(random_exploration)
        self.verbose, self.tensorboard_log, self.seed = verbose, tensorboard_log, seed
        self.ent_coef, self.target_entropy = ent_coef, target_entropy
        self.precision = precision
        self._dev = dict(device=device, rank=rank, nranks=nranks, nccl_id=nccl_id)
        self._layout_from_zip = False
        self.num_timesteps, self.n_updates = 0, 0
        self.episode_rewards = [0.0]
        self.ep_info_buf = deque(maxlen=100)
        self.learner: Optional[Learner] = None
        self.env = None
        self._vec_normalize_env = None
        self._rng = np.random.default_rng(seed)
        if env is not None:
            self.set_env(env)
            if _init_setup_model:
                self.setup_model()

    # ------------------------------------------------------------------ env plumbing
    def set_env(self, env):
        if not _is_vec(env):
            env = DummyVecEnv([lambda: env])
        self.env = env
        self.n_envs = env.num_envs
        self.observation_space, self.action_space = env.observation_space, env.action_space
        self._vec_normalize_env = unwrap_vec_normalize(env)

    def get_env(self):
        return self.env

    def get_vec_normalize_env(self):
        return self._vec_normalize_env

    def setup_model(self):
        obs_shape = tuple(self.observation_space.shape)
        n_act = int(np.prod(self.action_space.shape))
        if len(obs_shape) == 3 and "cnn_extractor" not in self.policy_kwargs and not self._layout_from_zip:
            # sb_helper.py:93-95: CnnPolicy with policy_kwargs={} means stable-baselines' plain nature_cnn over ALL planes and
            # no direct feature -- a different network (and different zip variables) from augmented_nature_cnn.  Refuse
            # instead of silently building the augmented extractor.  (SAC.load knows the layout from the zip itself.)
            raise NotImplementedError("CnnPolicy without policy_kwargs['cnn_extractor'] selects stable-baselines' plain nature_cnn "
                                      "(simplified + depth branch, sb_helper.py:93-95), which is not built; pass "
                                      "cnn_extractor=create_augmented_nature_cnn(1) as sb_helper.py:88-91 does")
        tgt = -float(n_act) if self.target_entropy == "auto" else float(self.target_entropy)
        self.learner = Learner(obs_shape, n_act=n_act, hidden=64, batch_size=self.batch_size, buffer_size=self.buffer_size,
                               gamma=self.gamma, tau=self.tau, target_entropy=tgt, seed=int(self.seed or 0),
                               precision=_PRECISIONS[self.precision], **self._dev)
        self._init_parameters()
        self._sync_norm_stats()

    def _init_parameters(self):
        """[SB2] ortho_init(sqrt 2) for conv/linear, Glorot-uniform for tf.layers.dense, zero biases,
        log_ent_coef = 0, target = copy of values_fn."""
        rng = np.random.default_rng(self.seed)
        p = OrderedDict()
        for name, shape in self.learner.param_shapes.items():
            if name.startswith("target/"):
                p[name] = p["model/" + name[len("target/"):]].copy()
            elif name.endswith("/w"):
                flat = (int(np.prod(shape[:-1])), shape[-1])
                u, _, v = np.linalg.svd(rng.standard_normal(flat), full_matrices=False)
                q = u if u.shape == flat else v
                p[name] = (np.sqrt(2.0) * q.reshape(shape)).astype(np.float32)
            elif name.endswith("/kernel"):
                lim = np.sqrt(6.0 / (shape[0] + shape[1]))
                p[name] = rng.uniform(-lim, lim, size=shape).astype(np.float32)
            else:
                p[name] = np.zeros(shape, np.float32)
        self.learner.load_parameters(p)

    def close(self):
        """Releases the device learner (replay ring included: 2 * buffer_size * obs_elems * 4 bytes of HBM)."""
        if self.learner is not None:
            self.learner.close()
            self.learner = None

    def _sync_norm_stats(self):
        vn = self._vec_normalize_env
        if vn is None:
            self.learner.set_norm_stats(norm_obs=False, norm_reward=False)
        else:
            self.learner.set_norm_stats(vn.obs_rms.mean, vn.obs_rms.var, float(vn.ret_rms.var), vn.clip_obs, vn.clip_reward,
                                        vn.epsilon, norm_obs=vn.norm_obs, norm_reward=vn.norm_reward)

    # ------------------------------------------------------------------ action scaling ([SB2] common/math_util.py)
    def _scale_action(self, a):
        low, high = self.action_space.low, self.action_space.high
        return 2.0 * ((a - low) / (high - low)) - 1.0

    def _unscale_action(self, a):
        low, high = self.action_space.low, self.action_space.high
        return low + 0.5 * (a + 1.0) * (high - low)

    # ------------------------------------------------------------------ learn
    def learn(self, total_timesteps, callback=None, log_interval=4, tb_log_name="SAC", reset_num_timesteps=True,
              replay_wrapper=None):
        """[SB2] SAC.learn: one env step then (every train_freq steps) gradient_steps minibatch updates."""
        if reset_num_timesteps:
            self.num_timesteps = 0
        callback = as_callback(callback)
        callback.init_callback(self)
        lr_fn = _constfn(self.learning_rate)
        vn = self._vec_normalize_env
        n_env = self.n_envs
        obs = self.env.reset()
        obs_ = vn.get_original_obs() if vn is not None else obs          # un-normalised copy stored in the replay
        ep_rew = np.zeros(n_env)
        infos_values = {}
        t_start = time.time()
        self._locals = {"self": self, "writer": None, "total_timesteps": total_timesteps}
        callback.on_training_start(self._locals, globals())
        callback.on_rollout_start()
        step = 0
        while step < total_timesteps:
            if self.num_timesteps < self.learning_starts or self._rng.random() < self.random_exploration:
                unscaled = np.stack([np.asarray(self.action_space.sample()) for _ in range(n_env)])
                action = self._scale_action(unscaled)
            else:
                src = obs_ if vn is not None else obs        # the device normalises raw obs with the same statistics
                if vn is not None:
                    self._sync_norm_stats()                  # act on the wrapper's CURRENT statistics, like SB does
                action = self.learner.act(np.asarray(src, np.float32), deterministic=False)
                unscaled = self._unscale_action(action)
            new_obs, reward, done, infos = self.env.step(unscaled)
            self.num_timesteps += n_env
            step += n_env
            if callback.on_step() is False:
                break
            new_obs_ = vn.get_original_obs() if vn is not None else new_obs
            reward_ = vn.get_original_reward() if vn is not None else reward
            # DummyVecEnv auto-resets: the transition's next_obs is the terminal observation
            nxt = np.array(new_obs_, np.float32, copy=True)
            for i, info in enumerate(infos):
                if done[i] and isinstance(info, dict) and "terminal_observation" in info:
                    nxt[i] = info["terminal_observation"]
            self.learner.replay_add(np.asarray(obs_, np.float32), np.asarray(action, np.float32), np.asarray(reward_, np.float32),
                                    nxt, np.asarray(done, np.float32))
            obs, obs_ = new_obs, new_obs_
            ep_rew += np.asarray(reward_, np.float64).reshape(-1)
            for i in range(n_env):
                if done[i]:
                    self.episode_rewards.append(float(ep_rew[i]))
                    self.ep_info_buf.append({"r": float(ep_rew[i])})
                    ep_rew[i] = 0.0
            if (self.num_timesteps // n_env) % self.train_freq == 0:
                callback.on_rollout_end()
                if self.learner.replay_size() >= self.batch_size and self.num_timesteps >= self.learning_starts:
                    if vn is not None:
                        self._sync_norm_stats()              # statistics current at sample time ([SB2] ReplayBuffer.sample(env=))
                    frac = 1.0 - step / total_timesteps
                    lr = float(lr_fn(frac))
                    # enqueue only: the gradient step runs on the device while the host goes on to the next env.step(); the losses
                    # (SB's infos_values, used for logging alone) are read back when something is actually logged
                    self.learner.step_async(self.gradient_steps, lr)
                    self.n_updates += self.gradient_steps
                    self._last_lr = lr
                callback.on_rollout_start()
            if self.verbose >= 1 and done.any() and log_interval and len(self.episode_rewards) % log_interval == 0:
                fps = int(step / max(1e-9, time.time() - t_start))
                if self.n_updates:
                    infos_values = self.learner.step(0, getattr(self, "_last_lr", 3e-4))       # 0 steps: just fetch the latest losses
                print({"episodes": len(self.episode_rewards), "mean 100 episode reward": round(float(np.mean(self.episode_rewards[-101:-1] or [0])), 1),
                       "n_updates": self.n_updates, "fps": fps, "total timesteps": self.num_timesteps,
                       **{k: infos_values.get(k) for k in ("policy_loss", "qf1_loss", "qf2_loss", "value_loss", "entropy", "ent_coef")}})
        callback.on_training_end()
        return self

    # ------------------------------------------------------------------ predict ([SB2] SAC.predict)
    def predict(self, observation, state=None, mask=None, deterministic=True):
        observation = np.asarray(observation, np.float32)
        single = observation.shape == tuple(self.observation_space.shape)
        obs = observation.reshape((-1,) + tuple(self.observation_space.shape))
        vn = self._vec_normalize_env
        if vn is not None and vn.norm_obs:
            # ``predict`` receives observations ALREADY normalised by the VecNormalize wrapper (utils.py:71 feeds
            # task.reset()/step() outputs); the device normalises raw ones, so undo the wrapper's transform.
            obs = obs * np.sqrt(vn.obs_rms.var + vn.epsilon) + vn.obs_rms.mean
            self._sync_norm_stats()
        act = self.learner.act(obs.astype(np.float32), deterministic=deterministic)
        act = self._unscale_action(act.reshape((-1,) + tuple(self.action_space.shape)))
        return (act[0] if single else act), None

    # ------------------------------------------------------------------ parameters / persistence
    def get_parameters(self):
        return OrderedDict((n + ":0", a) for n, a in self.learner.get_parameters().items())

    def load_parameters(self, load_path_or_dict, exact_match=True):
        params = load_path_or_dict
        if isinstance(params, str):
            _, params = sb_io.load_sb_zip(params)
        self.learner.load_parameters(params, exact_match=exact_match)

    def _data(self):
        return {
            "gamma": self.gamma, "learning_rate": self.learning_rate if not callable(self.learning_rate) else float(self.learning_rate(1.0)),
            "buffer_size": self.buffer_size, "learning_starts": self.learning_starts, "train_freq": self.train_freq,
            "batch_size": self.batch_size, "tau": self.tau, "ent_coef": self.ent_coef,
            "target_entropy": self.target_entropy if isinstance(self.target_entropy, str) else float(self.target_entropy),
            "verbose": self.verbose, "n_envs": getattr(self, "n_envs", 1), "seed": self.seed, "action_noise": None,
            "random_exploration": self.random_exploration, "_vectorize_action": True, "n_cpu_tf_sess": None,
            "policy": "CnnPolicy" if len(self.observation_space.shape) == 3 else "MlpPolicy",
            "policy_kwargs": {k: v for k, v in self.policy_kwargs.items() if k != "cnn_extractor"},
            "observation_space": {"shape": list(self.observation_space.shape), "low": float(np.min(self.observation_space.low)),
                                  "high": float(np.max(self.observation_space.high))},
            "action_space": {"shape": list(self.action_space.shape), "low": [float(x) for x in np.ravel(self.action_space.low)],
                             "high": [float(x) for x in np.ravel(self.action_space.high)]},
            "b200grasp": {"precision": self.precision, "n_updates": self.n_updates},
        }

    def save(self, save_path, cloudpickle=False):
        d = os.path.dirname(save_path)
        if d:
            os.makedirs(d, exist_ok=True)
        sb_io.save_sb_zip(save_path, self._data(), self.learner.get_parameters())

    @classmethod
    def load(cls, load_path, env=None, custom_objects=None, **kwargs):
        """Reads zips written by this class or by stable-baselines 2.10 (e.g.
        trained_models/SAC_depth_1mbuffer/best_model/best_model.zip)."""
        from .spaces import Box
        if not os.path.exists(load_path) and os.path.exists(load_path + ".zip"):
            load_path += ".zip"
        data, params = sb_io.load_sb_zip(load_path)
        if env is None:
            if "model/pi/cnn1/w" in params:
                c = params["model/pi/cnn1/w"].shape[2] + 1
                obs_space = Box(0.0, 255.0, (64, 64, c))
            else:
                obs_space = Box(-np.inf, np.inf, (params["model/pi/fc0/kernel"].shape[0],))
            n_act = params["model/pi/dense/kernel"].shape[1]

            class _Spaces:
                num_envs = 1
                observation_space = obs_space
                action_space = Box(-1.0, 1.0, (n_act,))
            env_like = _Spaces()
        else:
            env_like = env if _is_vec(env) else DummyVecEnv([lambda: env])
        kw = {}
        for k in ("gamma", "buffer_size", "learning_starts", "train_freq", "batch_size", "tau"):
            if isinstance(data.get(k), (int, float)):
                kw[k] = data[k]
        if isinstance(data.get("learning_rate"), (int, float)):
            kw["learning_rate"] = data["learning_rate"]
        # The zip's buffer_size (1e6 in every shipped model) is the TRAINING ring: 2 * 1e6 * obs_elems * 4 B = 65.6 GB for
        # depth, 164 GB for RGB-D.  A loaded model is used for inference or as a parameter donor (sb_helper.py:113-115 builds
        # a second model just to call get_parameters), so it gets a small ring unless the caller asks for one explicitly.
        kw["buffer_size"] = min(int(kw.get("buffer_size", 1000)), 1000)
        kw.update(kwargs)
        model = cls(policy=data.get("policy", "CnnPolicy"), env=None, _init_setup_model=False,
                    policy_kwargs={"layers": [64, 64]}, **kw)
        model._layout_from_zip = True
        model.env = env_like if env is not None else None
        model.n_envs = env_like.num_envs
        model.observation_space, model.action_space = env_like.observation_space, env_like.action_space
        model._vec_normalize_env = unwrap_vec_normalize(env_like) if env is not None else None
        model.setup_model()
        model.learner.load_parameters(params, exact_match=True)
        return model
