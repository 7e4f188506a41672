"""Independent numpy-float64 restatement of the SAC step with a HAND-DERIVED backward.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``; PARITY UNPINNED).  It exists to cross-check
``oracle/sac_ref.py`` (PyTorch autograd): two restatements written differently must agree to
float64 round-off.  It is also the written-out derivation the CUDA backward follows
(deep-rl-grasping_b200/csrc/heads.cu, conv_*.cu).

Same citations as sac_ref.py: custom_obs_policy.py:15-43 (CNN), [SB2] sac/policies.py
(make_actor / make_critics), [SB2] sac/sac.py setup_model (losses), SURVEY.md Appendix A.
"""
from __future__ import annotations

import math
import numpy as np

EPS = 1e-6
LOG_STD_MAX, LOG_STD_MIN = 2.0, -20.0
CONVS = (("cnn1", 8, 4), ("cnn2", 4, 2), ("cnn3", 3, 1))


def im2col(x, k, s):
    """x [B,H,W,C] -> cols [B*OH*OW, k*k*C], K index = (ky*k+kx)*C+c (HWIO filter flattening)."""
    b, h, w, c = x.shape
    oh, ow = (h - k) // s + 1, (w - k) // s + 1
    st = x.strides
    v = np.lib.stride_tricks.as_strided(x, (b, oh, ow, k, k, c), (st[0], st[1] * s, st[2] * s, st[1], st[2], st[3]))
    return np.ascontiguousarray(v).reshape(b * oh * ow, k * k * c), (b, oh, ow)


def col2im(dcols, xshape, k, s):
    b, h, w, c = xshape
    oh, ow = (h - k) // s + 1, (w - k) // s + 1
    d = dcols.reshape(b, oh, ow, k, k, c)
    dx = np.zeros(xshape, dcols.dtype)
    for ky in range(k):
        for kx in range(k):
            dx[:, ky:ky + s * oh:s, kx:kx + s * ow:s, :] += d[:, :, :, ky, kx, :]
    return dx


def cnn_fwd(x, p, prefix, c_img, n_direct=1):
    cache = {}
    h = np.ascontiguousarray(x[..., :c_img])
    feat = x[..., -1].reshape(x.shape[0], -1)[:, :n_direct]
    for name, k, s in CONVS:
        w = p[f"{prefix}/{name}/w"]
        cols, (b, oh, ow) = im2col(h, k, s)
        y = np.maximum(cols @ w.reshape(-1, w.shape[-1]) + p[f"{prefix}/{name}/b"].reshape(1, -1), 0.0)
        cache[name] = (h.shape, cols, y)
        h = y.reshape(b, oh, ow, -1)
    flat = h.reshape(h.shape[0], -1)
    h4 = np.maximum(flat @ p[f"{prefix}/cnn_fc1/w"] + p[f"{prefix}/cnn_fc1/b"], 0.0)
    cache["fc"] = (flat, h4)
    return np.concatenate([h4, feat], 1), cache


def cnn_bwd(dfeat, p, prefix, cache, grads):
    """dfeat [B,512+n_direct]; the direct-feature columns are inputs -> their gradient is dropped."""
    flat, h4 = cache["fc"]
    dz = dfeat[:, :512] * (h4 > 0)
    grads[f"{prefix}/cnn_fc1/w"] = flat.T @ dz
    grads[f"{prefix}/cnn_fc1/b"] = dz.sum(0)
    dh = dz @ p[f"{prefix}/cnn_fc1/w"].T
    for name, k, s in reversed(CONVS):
        xshape, cols, y = cache[name]
        w = p[f"{prefix}/{name}/w"]
        dz = dh.reshape(y.shape) * (y > 0)
        grads[f"{prefix}/{name}/w"] = (cols.T @ dz).reshape(w.shape)
        grads[f"{prefix}/{name}/b"] = dz.sum(0).reshape(p[f"{prefix}/{name}/b"].shape)
        if name != "cnn1":
            dh = col2im(dz @ w.reshape(-1, w.shape[-1]).T, xshape, k, s)


def mlp_fwd(z, p, prefix, out_name, nl):
    acts = [z]
    for i in range(nl):
        z = np.maximum(z @ p[f"{prefix}/fc{i}/kernel"] + p[f"{prefix}/fc{i}/bias"], 0.0)
        acts.append(z)
    outs = [z @ p[f"{prefix}/{o}/kernel"] + p[f"{prefix}/{o}/bias"] for o in out_name]
    return outs, acts


def mlp_bwd(douts, p, prefix, out_name, nl, acts, grads=None):
    """Returns d(input).  If ``grads`` is None only the input-gradient is formed (qf1 at pi)."""
    dz = 0.0
    for o, d in zip(out_name, douts):
        if grads is not None:
            grads[f"{prefix}/{o}/kernel"] = grads.get(f"{prefix}/{o}/kernel", 0.0) + acts[-1].T @ d
            grads[f"{prefix}/{o}/bias"] = grads.get(f"{prefix}/{o}/bias", 0.0) + d.sum(0)
        dz = dz + d @ p[f"{prefix}/{o}/kernel"].T
    for i in reversed(range(nl)):
        dz = dz * (acts[i + 1] > 0)
        if grads is not None:
            grads[f"{prefix}/fc{i}/kernel"] = grads.get(f"{prefix}/fc{i}/kernel", 0.0) + acts[i].T @ dz
            grads[f"{prefix}/fc{i}/bias"] = grads.get(f"{prefix}/fc{i}/bias", 0.0) + dz.sum(0)
        dz = dz @ p[f"{prefix}/fc{i}/kernel"].T
    return dz


def sac_grads(params, batch, eps_noise, cfg):
    """Forward + hand-derived backward in float64.  Returns (outputs, grads) like sac_ref.sac_step."""
    p = {n: np.asarray(a, np.float64) for n, a in params.items()}
    nl, A = len(cfg.layers), cfg.n_act
    scale = 255.0 if cfg.cnn else 1.0
    x = np.asarray(batch["obs"], np.float64) / scale
    xn = np.asarray(batch["next_obs"], np.float64) / scale
    act = np.asarray(batch["act"], np.float64)
    rew = np.asarray(batch["rew"], np.float64).reshape(-1, 1)
    done = np.asarray(batch["done"], np.float64).reshape(-1, 1)
    eps_n = np.asarray(eps_noise, np.float64)
    B = x.shape[0]

    def feats(xx, prefix):
        if cfg.cnn:
            return cnn_fwd(xx, p, prefix, cfg.c_img, cfg.n_direct)
        return xx.reshape(B, -1), None

    # ---- forward
    h_pi, c_pi = feats(x, "model/pi")
    (mu, ls_raw), a_pi = mlp_fwd(h_pi, p, "model/pi", ("dense", "dense_1"), nl)
    ls = np.clip(ls_raw, LOG_STD_MIN, LOG_STD_MAX)
    std = np.exp(ls)
    u = mu + eps_n * std
    t = (u - mu) / (std + EPS)
    pi = np.tanh(u)
    logp = (-0.5 * (t ** 2 + 2 * ls + math.log(2 * math.pi))).sum(1) - np.log(1 - pi ** 2 + EPS).sum(1)
    logp = logp.reshape(-1, 1)
    entropy = (ls + 0.5 * math.log(2 * math.pi * math.e)).sum(1).mean()
    h_v, c_v = feats(x, "model/values_fn")
    (v,), a_vf = mlp_fwd(h_v, p, "model/values_fn/vf", ("vf",), nl)
    (q1,), a_q1 = mlp_fwd(np.concatenate([h_v, act], 1), p, "model/values_fn/qf1", ("qf1",), nl)
    (q2,), a_q2 = mlp_fwd(np.concatenate([h_v, act], 1), p, "model/values_fn/qf2", ("qf2",), nl)
    (q1_pi,), a_q1p = mlp_fwd(np.concatenate([h_v, pi], 1), p, "model/values_fn/qf1", ("qf1",), nl)
    (q2_pi,), _ = mlp_fwd(np.concatenate([h_v, pi], 1), p, "model/values_fn/qf2", ("qf2",), nl)
    h_t, _ = feats(xn, "target/values_fn")
    (v_targ,), _ = mlp_fwd(h_t, p, "target/values_fn/vf", ("vf",), nl)
    log_alpha = p["model/log_ent_coef"]
    alpha = math.exp(float(log_alpha))
    q_backup = rew + (1 - done) * cfg.gamma * v_targ
    v_backup = np.minimum(q1_pi, q2_pi) - alpha * logp
    out = dict(q1=q1, q2=q2, v=v, logp=logp, pi=pi, q1_pi=q1_pi, q2_pi=q2_pi, v_targ=v_targ, h_pi=h_pi, h_v=h_v,
               qf1_loss=0.5 * ((q_backup - q1) ** 2).mean(), qf2_loss=0.5 * ((q_backup - q2) ** 2).mean(),
               value_loss=0.5 * ((v - v_backup) ** 2).mean(), policy_loss=(alpha * logp - q1_pi).mean(),
               ent_coef_loss=-(float(log_alpha) * (logp + cfg.target_entropy)).mean(), entropy=entropy, ent_coef=alpha)

    # ---- backward
    g = {}
    # values: d(L_q1 + L_q2 + L_v)
    dF = mlp_bwd([(v - v_backup) / B], p, "model/values_fn/vf", ("vf",), nl, a_vf, g)
    dF = dF + mlp_bwd([(q1 - q_backup) / B], p, "model/values_fn/qf1", ("qf1",), nl, a_q1, g)[:, :cfg.feat_dim]
    dF = dF + mlp_bwd([(q2 - q_backup) / B], p, "model/values_fn/qf2", ("qf2",), nl, a_q2, g)[:, :cfg.feat_dim]
    if cfg.cnn:
        cnn_bwd(dF, p, "model/values_fn", c_v, g)
    # policy: d mean(alpha*logp - Q1(s, pi)); qf1 weights are constants here, only d/d(pi) is used
    dpi_q = mlp_bwd([np.full((B, 1), -1.0 / B)], p, "model/values_fn/qf1", ("qf1",), nl, a_q1p, None)[:, cfg.feat_dim:]
    one_m = 1 - pi ** 2
    du = (alpha / B) * 2 * pi * one_m / (one_m + EPS) + dpi_q * one_m
    dls = du * eps_n * std + (alpha / B) * (-t * eps_n * std * EPS / (std + EPS) ** 2 - 1.0)
    dls = dls * ((ls_raw >= LOG_STD_MIN) & (ls_raw <= LOG_STD_MAX))
    dFp = mlp_bwd([du, dls], p, "model/pi", ("dense", "dense_1"), nl, a_pi, g)
    if cfg.cnn:
        cnn_bwd(dFp, p, "model/pi", c_pi, g)
    g["model/log_ent_coef"] = np.float64(-(logp + cfg.target_entropy).mean())
    grads = {n: np.asarray(g[n]).reshape(np.shape(params[n])) for n in params if not n.startswith("target/")}
    return out, grads
