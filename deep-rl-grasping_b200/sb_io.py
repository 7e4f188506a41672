"""Readers/writers for the reference's on-disk artefacts (no stable-baselines / gym needed).

* SB zip (``model.save``; /root/reference/manipulation_main/training/sb_helper.py:228-244):
  entries ``data`` (JSON; non-JSON values base64-cloudpickled under ":serialized:"),
  ``parameter_list`` (JSON list of TF variable names, ":0"-suffixed), ``parameters`` (np.savez).
* ``vecnormalize.pkl`` (``VecNormalize.save``; sb_helper.py:246-247, train_stable_baselines.py:88-91):
  a pickled ``stable_baselines.common.vec_env.VecNormalize``; read through a stub unpickler.
"""
from __future__ import annotations

import io
import json
import pickle
import zipfile
from collections import OrderedDict
from typing import Any, Dict, Tuple

import numpy as np


def load_sb_zip(path: str) -> Tuple[Dict[str, Any], "OrderedDict[str, np.ndarray]"]:
    """Returns (data, params).  Param names have the ':0' suffix stripped; order = parameter_list."""
    with zipfile.ZipFile(path) as z:
        data_raw = json.loads(z.read("data").decode())
        names = json.loads(z.read("parameter_list").decode())
        arrs = np.load(io.BytesIO(z.read("parameters")))
        params = OrderedDict((n[:-2] if n.endswith(":0") else n, np.asarray(arrs[n])) for n in names)
    data = {}
    for k, v in data_raw.items():
        if isinstance(v, dict) and ":serialized:" in v:
            data[k] = {kk: vv for kk, vv in v.items() if kk != ":serialized:"}
            data[k]["__serialized__"] = True
        else:
            data[k] = v
    return data, params


def save_sb_zip(path: str, data: Dict[str, Any], params: "OrderedDict[str, np.ndarray]") -> None:
    """Writes the three-entry SB zip.  Only JSON-able ``data`` values are stored (SB's loader
    tolerates missing cloudpickled fields when ``custom_objects`` / constructor kwargs supply them)."""
    names = [n + ":0" for n in params]
    buf = io.BytesIO()
    np.savez(buf, **{n + ":0": np.asarray(a, np.float32) for n, a in params.items()})
    clean = {}
    for k, v in data.items():
        try:
            json.dumps(v)
            clean[k] = v
        except TypeError:
            clean[k] = {":type:": str(type(v)), "repr": repr(v)}
    if not path.endswith(".zip"):
        path += ".zip"
    with zipfile.ZipFile(path, "w") as z:
        z.writestr("data", json.dumps(clean, indent=4))
        z.writestr("parameter_list", json.dumps(names))
        z.writestr("parameters", buf.getvalue())


class _Stub:
    """Placeholder for any stable_baselines.* / gym.* class met while unpickling."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self.__dict__["_state"] = state


class _StubUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.split(".")[0] in ("stable_baselines", "gym", "manipulation_main", "pybullet_envs"):
            return type(name, (_Stub,), {"__module__": module})
        if module.startswith("numpy.core"):
            module = module.replace("numpy.core", "numpy._core", 1)
        return super().find_class(module, name)


def load_vecnormalize(path: str) -> Dict[str, Any]:
    """Extracts obs_rms / ret_rms statistics + clipping constants from a vecnormalize.pkl."""
    with open(path, "rb") as f:
        o = _StubUnpickler(f).load()
    d = o.__dict__
    out = dict(
        obs_mean=np.asarray(d["obs_rms"].mean, np.float64), obs_var=np.asarray(d["obs_rms"].var, np.float64),
        obs_count=float(d["obs_rms"].count),
        ret_mean=float(np.asarray(d["ret_rms"].mean)), ret_var=float(np.asarray(d["ret_rms"].var)),
        ret_count=float(d["ret_rms"].count),
        clip_obs=float(d["clip_obs"]), clip_reward=float(d["clip_reward"]), epsilon=float(d["epsilon"]),
        gamma=float(d["gamma"]), norm_obs=bool(d.get("norm_obs", True)), norm_reward=bool(d.get("norm_reward", True)),
    )
    for k in ("old_obs", "old_rews"):
        if d.get(k) is not None:
            out[k] = np.asarray(d[k])
    return out
