"""Shared helpers for the parity tests: seeded batches from the committed fixtures."""
import os

import numpy as np

import b200grasp
from b200grasp import synth
from oracle import sac_ref as R

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(key):
    """key in {'sac_depth', 'sac_encoder'} -> (cfg, trained params, vecnormalize stats)."""
    vn = dict(np.load(os.path.join(GOLD, f"vecnorm_{key}.npz")))
    cfg = R.SACConfig(obs_shape=tuple(vn["obs_mean"].shape))
    raw = dict(np.load(os.path.join(GOLD, f"{key}_params.npz")))
    params = {n: raw[n] for n, _ in R.param_specs(cfg)}
    return cfg, params, vn


def make_batch(vn, B, seed=synth.DATA_SEED):
    raw = synth.make_transitions(B, vn["obs_mean"], vn["obs_var"], seed=seed)
    norm = dict(obs=R.normalize_obs(raw["obs"], vn["obs_mean"], vn["obs_var"]),
                next_obs=R.normalize_obs(raw["next_obs"], vn["obs_mean"], vn["obs_var"]),
                act=raw["act"], rew=R.normalize_reward(raw["rew"], float(vn["ret_var"])), done=raw["done"])
    return raw, norm, synth.make_eps(B, seed=seed + 1)


def make_learner(cfg, vn, B, params=None, buffer_size=1024, precision=0, **kw):
    L = b200grasp.Learner(cfg.obs_shape, n_act=cfg.n_act, batch_size=B, buffer_size=buffer_size, gamma=cfg.gamma,
                          tau=cfg.tau, target_entropy=cfg.target_entropy, precision=precision, **kw)
    L.set_norm_stats(vn["obs_mean"], vn["obs_var"], float(vn["ret_var"]), float(vn["clip_obs"]), float(vn["clip_reward"]),
                     float(vn["epsilon"]))
    if params is not None:
        L.load_parameters(params)
    return L


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))
