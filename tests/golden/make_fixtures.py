"""Regenerates tests/golden/* from /root/reference (run in the dev container only).

The GPU box never sees /root/reference, so everything parity tests need is extracted here:
  * trained parameter sets (names/shapes/values) from the shipped SB zips,
  * VecNormalize statistics + one real frame from vecnormalize.pkl,
  * the first rows of the reference's training logs (known-answer scalars),
  * oracle outputs on the seeded synthetic batch (regression vectors for the CUDA path).
Usage:  python tests/golden/make_fixtures.py
"""
import importlib.util
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference/trained_models"
OUT = os.path.join(ROOT, "tests", "golden")


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "deep-rl-grasping_b200", name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


sb_io, synth = _load("sb_io"), _load("synth")
from oracle import sac_ref as R  # noqa: E402

ZIPS = {
    "sac_depth": "SAC_depth_1mbuffer/best_model/best_model.zip",
    "sac_rgbd": "SAC_full_rgbd/SAC_full_rgbd.zip",
    "sac_encoder": "SAC_encoder_1mbuffer/SAC_best_1mbuffer_5m.zip",
    "bdq_8pads": "BDQ_8pads/BDQ_simple_8pads.zip",
    "bdq_33big": "BDQ_33pads_big/BDQ_33_big.zip",
}
VECN = {
    "sac_depth": "SAC_depth_1mbuffer/best_model/vecnormalize.pkl",
    "sac_rgbd": "SAC_full_rgbd/vecnormalize.pkl",
    "sac_encoder": "SAC_encoder_1mbuffer/vecnormalize.pkl",
}

manifest = {}
for key, rel in ZIPS.items():
    data, params = sb_io.load_sb_zip(os.path.join(REF, rel))
    manifest[key] = {
        "source": rel,
        "shapes": {n: list(a.shape) for n, a in params.items()},
        "n_floats": int(sum(a.size for a in params.values())),
        "data": {k: v for k, v in data.items() if isinstance(v, (int, float, str, bool, type(None)))},
    }
    if key in ("sac_depth", "sac_encoder"):       # values kept: realistic magnitudes for parity tests
        np.savez_compressed(os.path.join(OUT, key + "_params.npz"), **{n: a.astype(np.float32) for n, a in params.items()})
json.dump(manifest, open(os.path.join(OUT, "zip_manifest.json"), "w"), indent=1, sort_keys=True)

for key, rel in VECN.items():
    v = sb_io.load_vecnormalize(os.path.join(REF, rel))
    np.savez_compressed(os.path.join(OUT, "vecnorm_" + key + ".npz"), **{k: np.asarray(x) for k, x in v.items()})

import pandas as pd  # noqa: E402
logs = {}
for key, rel in {"sac_rgbd": "SAC_full_rgbd/logs.csv", "sac_depth": "SAC_depth_1mbuffer/logs.csv",
                 "sac_encoder": "SAC_encoder_1mbuffer/logs.full.csv"}.items():
    logs[key] = pd.read_csv(os.path.join(REF, rel)).head(4).to_dict(orient="list")
json.dump(logs, open(os.path.join(OUT, "logs_head.json"), "w"), indent=1)

# ---- oracle regression vectors on the seeded batch (trained depth weights, B=32)
for key, B in (("sac_depth", 32), ("sac_encoder", 64)):
    P = dict(np.load(os.path.join(OUT, key + "_params.npz")))
    vn = dict(np.load(os.path.join(OUT, "vecnorm_" + key + ".npz")))
    cfg = R.SACConfig(obs_shape=tuple(vn["obs_mean"].shape))
    P = {n: P[n] for n, _ in R.param_specs(cfg)}
    raw = synth.make_transitions(B, vn["obs_mean"], vn["obs_var"])
    batch = dict(obs=R.normalize_obs(raw["obs"], vn["obs_mean"], vn["obs_var"]),
                 next_obs=R.normalize_obs(raw["next_obs"], vn["obs_mean"], vn["obs_var"]),
                 act=raw["act"], rew=R.normalize_reward(raw["rew"], float(vn["ret_var"])), done=raw["done"])
    eps = synth.make_eps(B)
    res, grads, newp, _ = R.sac_step(P, R.OptState.zeros(P), batch, eps, 3e-4, cfg, torch.float64)
    keep = {k: np.asarray(res[k], np.float64) for k in
            ("q1", "q2", "v", "logp", "pi", "q1_pi", "q2_pi", "v_targ", "policy_loss", "qf1_loss", "qf2_loss",
             "value_loss", "ent_coef_loss", "entropy", "ent_coef", "grad_norm_pi", "grad_norm_values", "grad_ent")}
    keep["grad_norms"] = np.array([np.sqrt((grads[n].astype(np.float64) ** 2).sum()) for n in grads])
    keep["grad_names"] = np.array(list(grads.keys()))
    keep["param_delta_norms"] = np.array([np.sqrt(((newp[n].astype(np.float64) - P[n]) ** 2).sum()) for n in P])
    np.savez_compressed(os.path.join(OUT, f"golden_step_{key}_b{B}.npz"), **keep)
    print(key, {k: float(v) for k, v in keep.items() if np.ndim(v) == 0})
print("fixtures written to", OUT)
