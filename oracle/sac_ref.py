"""PyTorch-CPU restatement of the stable-baselines 2.10.1 SAC minibatch step.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  PARITY UNPINNED (TF1/SB2 cannot run
here); every function cites the reference call site or the SB2 function it restates.

What is restated:
  * ``augmented_nature_cnn``      /root/reference/manipulation_main/training/custom_obs_policy.py:15-43
  * ``SACPolicy.make_actor``      [SB2] stable_baselines/sac/policies.py (vars model/pi/*)
  * ``SACPolicy.make_critics``    [SB2] stable_baselines/sac/policies.py (vars model/values_fn/*)
  * losses + 3 Adam + Polyak      [SB2] stable_baselines/sac/sac.py ``setup_model`` / ``_train_step``
    constructed at /root/reference/manipulation_main/training/sb_helper.py:104-128
  * VecNormalize at sample time   [SB2] common/vec_env/vec_normalize.py, configured at
    /root/reference/manipulation_main/training/sb_helper.py:118-119
Parameter names and layouts are the ones in the shipped zips
(``trained_models/SAC_depth_1mbuffer/best_model/best_model.zip``): conv filters HWIO, conv bias
(1,n,1,1), dense kernels [in,out].
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

EPS = 1e-6            # [SB2] sac/policies.py EPS
LOG_STD_MAX = 2.0     # [SB2] sac/policies.py
LOG_STD_MIN = -20.0
ADAM_B1, ADAM_B2, ADAM_EPS = 0.9, 0.999, 1e-8   # tf.train.AdamOptimizer defaults


@dataclass
class SACConfig:
    """Shapes + hyper-parameters.  Defaults = BASELINE config 2 (config/gripper_grasp.yaml:69-86)."""
    obs_shape: Tuple[int, ...] = (64, 64, 2)   # (H, W, C_img + n_direct) or (D,) for the MLP policy
    n_act: int = 5                             # actuator.py:72-73
    layers: Tuple[int, ...] = (64, 64)         # config/gripper_grasp.yaml:81
    n_direct: int = 1                          # sb_helper.py:89 create_augmented_nature_cnn(1)
    gamma: float = 0.99                        # config/gripper_grasp.yaml:73
    tau: float = 0.005                         # zip 'tau'
    target_entropy: float = -5.0               # zip 'target_entropy' (= -n_act)

    @property
    def cnn(self) -> bool:
        return len(self.obs_shape) == 3

    @property
    def c_img(self) -> int:
        return self.obs_shape[2] - 1       # custom_obs_policy.py:32 drops exactly the last plane

    @property
    def feat_dim(self) -> int:
        return (512 + self.n_direct) if self.cnn else int(self.obs_shape[0])


def _cnn_specs(prefix: str, c_img: int) -> List[Tuple[str, Tuple[int, ...]]]:
    return [
        (f"{prefix}/cnn1/w", (8, 8, c_img, 32)), (f"{prefix}/cnn1/b", (1, 32, 1, 1)),
        (f"{prefix}/cnn2/w", (4, 4, 32, 64)), (f"{prefix}/cnn2/b", (1, 64, 1, 1)),
        (f"{prefix}/cnn3/w", (3, 3, 64, 64)), (f"{prefix}/cnn3/b", (1, 64, 1, 1)),
        (f"{prefix}/cnn_fc1/w", (1024, 512)), (f"{prefix}/cnn_fc1/b", (512,)),
    ]


def _mlp_specs(prefix: str, in_dim: int, layers: Sequence[int]) -> List[Tuple[str, Tuple[int, ...]]]:
    out, d = [], in_dim
    for i, h in enumerate(layers):
        out += [(f"{prefix}/fc{i}/kernel", (d, h)), (f"{prefix}/fc{i}/bias", (h,))]
        d = h
    return out


def param_specs(cfg: SACConfig) -> List[Tuple[str, Tuple[int, ...]]]:
    """Ordered (name, shape) list; same order as ``parameter_list`` in the SB zips."""
    h = cfg.layers[-1]
    specs: List[Tuple[str, Tuple[int, ...]]] = []
    if cfg.cnn:
        specs += _cnn_specs("model/pi", cfg.c_img)
    specs += _mlp_specs("model/pi", cfg.feat_dim, cfg.layers)
    specs += [("model/pi/dense/kernel", (h, cfg.n_act)), ("model/pi/dense/bias", (cfg.n_act,)),
              ("model/pi/dense_1/kernel", (h, cfg.n_act)), ("model/pi/dense_1/bias", (cfg.n_act,))]
    for scope in ("model/values_fn", "target/values_fn"):
        sub: List[Tuple[str, Tuple[int, ...]]] = []
        if cfg.cnn:
            sub += _cnn_specs(scope, cfg.c_img)
        sub += _mlp_specs(f"{scope}/vf", cfg.feat_dim, cfg.layers)
        sub += [(f"{scope}/vf/vf/kernel", (h, 1)), (f"{scope}/vf/vf/bias", (1,))]
        if scope == "model/values_fn":
            for q in ("qf1", "qf2"):
                sub += _mlp_specs(f"{scope}/{q}", cfg.feat_dim + cfg.n_act, cfg.layers)
                sub += [(f"{scope}/{q}/{q}/kernel", (h, 1)), (f"{scope}/{q}/{q}/bias", (1,))]
            specs += sub + [("model/log_ent_coef", ())]
        else:
            specs += sub
    return specs


def group_of(name: str) -> str:
    """Which optimiser owns a variable ([SB2] get_vars('model/pi') etc.)."""
    if name.startswith("model/pi/"):
        return "pi"
    if name.startswith("model/values_fn/"):
        return "values"
    if name == "model/log_ent_coef":
        return "ent"
    return "target"


def _ortho(rng: np.random.Generator, shape, scale) -> np.ndarray:
    """[SB2] tf_layers.ortho_init: SVD-orthogonal flat matrix reshaped; init_scale sqrt(2)."""
    if len(shape) == 2:
        flat = shape
    else:
        flat = (int(np.prod(shape[:-1])), shape[-1])
    a = rng.standard_normal(flat)
    u, _, v = np.linalg.svd(a, full_matrices=False)
    q = u if u.shape == flat else v
    return (scale * q.reshape(shape)).astype(np.float32)


def init_params(cfg: SACConfig, seed: int = 0) -> "OrderedDict[str, np.ndarray]":
    """Fresh init: ortho(sqrt2) conv/linear, Glorot-uniform tf.layers.dense, zero bias, log_alpha=0;
    target = copy of model/values_fn ([SB2] SAC.setup_model target_init_op)."""
    rng = np.random.default_rng(seed)
    p: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for name, shape in param_specs(cfg):
        if name.startswith("target/"):
            p[name] = p["model/" + name[len("target/"):]].copy()
        elif name.endswith("/w"):
            p[name] = _ortho(rng, shape, math.sqrt(2.0))
        elif name.endswith("/kernel"):
            lim = math.sqrt(6.0 / (shape[0] + shape[1]))
            p[name] = rng.uniform(-lim, lim, size=shape).astype(np.float32)
        else:
            p[name] = np.zeros(shape, np.float32)
    return p


# --------------------------------------------------------------------------------------------
# VecNormalize at sample time ([SB2] ReplayBuffer.sample(env=vec_normalize) ->
# VecNormalize.normalize_obs / normalize_reward; float64 numpy then cast to fp32 by feed_dict)
# --------------------------------------------------------------------------------------------
def normalize_obs(obs: np.ndarray, mean: np.ndarray, var: np.ndarray, clip: float = 10.0,
                  eps: float = 1e-8) -> np.ndarray:
    o = (obs.astype(np.float64) - mean.astype(np.float64)) / np.sqrt(var.astype(np.float64) + eps)
    return np.clip(o, -clip, clip).astype(np.float32)


def normalize_reward(rew: np.ndarray, ret_var: float, clip: float = 10.0, eps: float = 1e-8) -> np.ndarray:
    r = rew.astype(np.float64) / np.sqrt(np.float64(ret_var) + eps)
    return np.clip(r, -clip, clip).astype(np.float32)


# --------------------------------------------------------------------------------------------
# networks
# --------------------------------------------------------------------------------------------
def cnn_features(x: torch.Tensor, p: Dict[str, torch.Tensor], prefix: str, cfg: SACConfig,
                 keep: Optional[dict] = None) -> torch.Tensor:
    """augmented_nature_cnn (custom_obs_policy.py:15-43). ``x`` is NHWC, already /255."""
    c = cfg.c_img
    img = x[..., :c]
    feat = x[..., -1].reshape(x.shape[0], -1)[:, :cfg.n_direct]   # custom_obs_policy.py:28-30
    h = img.permute(0, 3, 1, 2)
    acts = []
    for name, stride in (("cnn1", 4), ("cnn2", 2), ("cnn3", 1)):
        w = p[f"{prefix}/{name}/w"].permute(3, 2, 0, 1)           # HWIO -> OIHW (cross-correlation)
        b = p[f"{prefix}/{name}/b"].reshape(1, -1, 1, 1)
        h = F.relu(F.conv2d(h, w, stride=stride) + b)             # VALID
        acts.append(h)
    flat = h.permute(0, 2, 3, 1).reshape(h.shape[0], -1)          # conv_to_fc on NHWC: (y*4+x)*64+c
    h4 = F.relu(flat @ p[f"{prefix}/cnn_fc1/w"] + p[f"{prefix}/cnn_fc1/b"])
    if keep is not None:
        keep[prefix] = acts + [h4]
    return torch.cat([h4, feat], dim=1)


def features(x: torch.Tensor, p, prefix: str, cfg: SACConfig, keep=None) -> torch.Tensor:
    return cnn_features(x, p, prefix, cfg, keep) if cfg.cnn else x.reshape(x.shape[0], -1)


def mlp(z: torch.Tensor, p, prefix: str, n_layers: int) -> torch.Tensor:
    for i in range(n_layers):
        z = F.relu(z @ p[f"{prefix}/fc{i}/kernel"] + p[f"{prefix}/fc{i}/bias"])
    return z


def forward_losses(p: Dict[str, torch.Tensor], batch: Dict[str, torch.Tensor], eps_noise: torch.Tensor,
                   cfg: SACConfig, keep: Optional[dict] = None) -> Dict[str, torch.Tensor]:
    """SURVEY Appendix A.  ``batch`` holds the *normalised* (obs, act, rew, next_obs, done)."""
    nl = len(cfg.layers)
    scale = 255.0 if cfg.cnn else 1.0                 # observation_input(scale=cnn), Box(0,255) robot.py:224-228
    x = batch["obs"] / scale
    xn = batch["next_obs"] / scale
    act = batch["act"]
    rew = batch["rew"].reshape(-1, 1)
    done = batch["done"].reshape(-1, 1)

    # actor
    h_pi = features(x, p, "model/pi", cfg, keep)
    g = mlp(h_pi, p, "model/pi", nl)
    mu = g @ p["model/pi/dense/kernel"] + p["model/pi/dense/bias"]
    log_std = torch.clamp(g @ p["model/pi/dense_1/kernel"] + p["model/pi/dense_1/bias"], LOG_STD_MIN, LOG_STD_MAX)
    std = torch.exp(log_std)
    u = mu + eps_noise * std
    logp = (-0.5 * (((u - mu) / (std + EPS)) ** 2 + 2 * log_std + math.log(2 * math.pi))).sum(1)
    entropy = (log_std + 0.5 * math.log(2 * math.pi * math.e)).sum(1)
    pi = torch.tanh(u)
    logp = logp - torch.log(1 - pi ** 2 + EPS).sum(1)
    logp = logp.reshape(-1, 1)

    # critics (one shared feature; second make_critics(reuse=True) re-evaluates the same function)
    h_v = features(x, p, "model/values_fn", cfg, keep)

    def head(prefix, inp, out_name):
        z = mlp(inp, p, prefix, nl)
        return z @ p[f"{prefix}/{out_name}/kernel"] + p[f"{prefix}/{out_name}/bias"]

    v = head("model/values_fn/vf", h_v, "vf")
    q1 = head("model/values_fn/qf1", torch.cat([h_v, act], 1), "qf1")
    q2 = head("model/values_fn/qf2", torch.cat([h_v, act], 1), "qf2")
    q1_pi = head("model/values_fn/qf1", torch.cat([h_v, pi], 1), "qf1")
    q2_pi = head("model/values_fn/qf2", torch.cat([h_v, pi], 1), "qf2")
    h_t = features(xn, p, "target/values_fn", cfg, keep)
    v_targ = head("target/values_fn/vf", h_t, "vf")

    log_alpha = p["model/log_ent_coef"]
    alpha = torch.exp(log_alpha)
    q_backup = (rew + (1 - done) * cfg.gamma * v_targ).detach()
    qf1_loss = 0.5 * ((q_backup - q1) ** 2).mean()
    qf2_loss = 0.5 * ((q_backup - q2) ** 2).mean()
    v_backup = (torch.minimum(q1_pi, q2_pi) - alpha * logp).detach()
    value_loss = 0.5 * ((v - v_backup) ** 2).mean()
    policy_loss = (alpha * logp - q1_pi).mean()
    ent_coef_loss = -(log_alpha * (logp + cfg.target_entropy).detach()).mean()
    return dict(q1=q1, q2=q2, v=v, logp=logp, pi=pi, mu=mu, log_std=log_std, q1_pi=q1_pi, q2_pi=q2_pi,
                v_targ=v_targ, h_pi=h_pi, h_v=h_v,
                policy_loss=policy_loss, qf1_loss=qf1_loss, qf2_loss=qf2_loss, value_loss=value_loss,
                ent_coef_loss=ent_coef_loss, entropy=entropy.mean(), ent_coef=alpha)


# --------------------------------------------------------------------------------------------
# one gradient step
# --------------------------------------------------------------------------------------------
@dataclass
class OptState:
    """Three tf.train.AdamOptimizer instances (policy / values / entropy), each with its own t."""
    m: Dict[str, np.ndarray] = field(default_factory=dict)
    v: Dict[str, np.ndarray] = field(default_factory=dict)
    t: Dict[str, int] = field(default_factory=lambda: {"pi": 0, "values": 0, "ent": 0})

    @staticmethod
    def zeros(params) -> "OptState":
        st = OptState()
        for n, a in params.items():
            if group_of(n) != "target":
                st.m[n] = np.zeros_like(a, dtype=np.float64)
                st.v[n] = np.zeros_like(a, dtype=np.float64)
        return st


def sac_step(params: Dict[str, np.ndarray], opt: OptState, batch: Dict[str, np.ndarray], eps_noise: np.ndarray,
             lr: float, cfg: SACConfig, dtype=torch.float32):
    """One ``SAC._train_step`` + ``target_update_op``.  Returns (outputs, grads, new_params, new_opt).

    Gradient ownership follows ``minimize(..., var_list=...)``: policy loss -> model/pi only;
    qf1+qf2+value loss -> model/values_fn only; ent_coef_loss -> log_ent_coef.  All gradients are
    taken from the same pre-update forward; Adam is the TF1 form
    (lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps)).
    """
    np_dt = np.float64 if dtype == torch.float64 else np.float32
    tp = {n: torch.tensor(np.asarray(a, dtype=np_dt), dtype=dtype, requires_grad=(group_of(n) != "target"))
          for n, a in params.items()}
    tb = {k: torch.tensor(np.asarray(v, dtype=np_dt), dtype=dtype) for k, v in batch.items()}
    te = torch.tensor(np.asarray(eps_noise, dtype=np_dt), dtype=dtype)
    out = forward_losses(tp, tb, te, cfg)

    names = {g: [n for n in tp if group_of(n) == g] for g in ("pi", "values", "ent")}
    g_pi = torch.autograd.grad(out["policy_loss"], [tp[n] for n in names["pi"]], retain_graph=True)
    g_v = torch.autograd.grad(out["qf1_loss"] + out["qf2_loss"] + out["value_loss"],
                              [tp[n] for n in names["values"]], retain_graph=True)
    g_e = torch.autograd.grad(out["ent_coef_loss"], [tp[n] for n in names["ent"]])
    grads: Dict[str, np.ndarray] = {}
    for g, gl in (("pi", g_pi), ("values", g_v), ("ent", g_e)):
        for n, t in zip(names[g], gl):
            grads[n] = t.detach().numpy().astype(np_dt)

    new_p = OrderedDict((n, np.asarray(a, dtype=np_dt).copy()) for n, a in params.items())
    new_opt = OptState(m=dict(opt.m), v=dict(opt.v), t=dict(opt.t))
    for g in ("pi", "values", "ent"):
        new_opt.t[g] = opt.t[g] + 1
        t = new_opt.t[g]
        lr_t = np_dt(lr) * np.sqrt(np_dt(1) - np_dt(ADAM_B2) ** t) / (np_dt(1) - np_dt(ADAM_B1) ** t)
        for n in names[g]:
            gr = grads[n]
            m = (ADAM_B1 * opt.m[n] + (1 - ADAM_B1) * gr).astype(np_dt)
            v = (ADAM_B2 * opt.v[n] + (1 - ADAM_B2) * gr * gr).astype(np_dt)
            new_opt.m[n], new_opt.v[n] = m, v
            new_p[n] = (new_p[n] - lr_t * m / (np.sqrt(v) + np_dt(ADAM_EPS))).astype(np_dt)
    # target_update_op: runs after the step ops, on the UPDATED model/values_fn
    for n in params:
        if n.startswith("target/"):
            src = "model/" + n[len("target/"):]
            new_p[n] = ((1 - cfg.tau) * new_p[n] + cfg.tau * new_p[src]).astype(np_dt)

    res = {k: (v.detach().numpy() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}
    res["grad_norm_pi"] = float(np.sqrt(sum(float((grads[n].astype(np.float64) ** 2).sum()) for n in names["pi"])))
    res["grad_norm_values"] = float(np.sqrt(sum(float((grads[n].astype(np.float64) ** 2).sum()) for n in names["values"])))
    res["grad_ent"] = float(grads["model/log_ent_coef"])
    return res, grads, new_p, new_opt


def policy_act(params: Dict[str, np.ndarray], obs_norm: np.ndarray, cfg: SACConfig, deterministic: bool = True,
               eps_noise: Optional[np.ndarray] = None) -> np.ndarray:
    """[SB2] SAC.predict -> policy_tf.step: tanh(mu) if deterministic else tanh(mu + eps*std)."""
    tp = {n: torch.tensor(a, dtype=torch.float32) for n, a in params.items() if n.startswith("model/pi/")}
    x = torch.tensor(obs_norm, dtype=torch.float32) / (255.0 if cfg.cnn else 1.0)
    h = features(x, tp, "model/pi", cfg)
    g = mlp(h, tp, "model/pi", len(cfg.layers))
    mu = g @ tp["model/pi/dense/kernel"] + tp["model/pi/dense/bias"]
    if deterministic:
        return torch.tanh(mu).numpy()
    ls = torch.clamp(g @ tp["model/pi/dense_1/kernel"] + tp["model/pi/dense_1/bias"], LOG_STD_MIN, LOG_STD_MAX)
    return torch.tanh(mu + torch.tensor(eps_noise, dtype=torch.float32) * torch.exp(ls)).numpy()
