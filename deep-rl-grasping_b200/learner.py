"""Thin numpy-facing wrapper of one ``b2g_sac`` handle (the device-resident learner).

``Learner`` is what ``SAC`` (sac.py, the stable-baselines-shaped front end) drives; tests and
bench.py also use it directly because it maps 1:1 onto the C ABI entry points.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Dict, Optional, Sequence

import numpy as np

from . import _lib


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _f32(a) -> np.ndarray:
    a = np.asarray(a, dtype=np.float32)          # (ascontiguousarray would promote 0-d to 1-d)
    return a if a.flags.c_contiguous else a.copy()


class Learner:
    def __init__(self, obs_shape: Sequence[int], n_act: int = 5, hidden: int = 64, batch_size: int = 64,
                 buffer_size: int = 100000, gamma: float = 0.99, tau: float = 0.005,
                 target_entropy: Optional[float] = None, seed: int = 0, precision: int = _lib.B2G_PREC_FP32_SIMT,
                 device: int = 0, rank: int = 0, nranks: int = 1, nccl_id: Optional[bytes] = None):
        self.lib = _lib.load()
        self.obs_shape = tuple(int(s) for s in obs_shape)
        self.n_act, self.batch_size = int(n_act), int(batch_size)
        cfg = _lib.SacCfg()
        if len(self.obs_shape) == 3:
            cfg.obs_h, cfg.obs_w, cfg.obs_c = self.obs_shape
            cfg.obs_dim = 0
        elif len(self.obs_shape) == 1:
            cfg.obs_h = cfg.obs_w = cfg.obs_c = 0
            cfg.obs_dim = self.obs_shape[0]
        else:
            raise ValueError(f"unsupported observation shape {obs_shape}")
        cfg.n_act, cfg.hidden, cfg.batch, cfg.buffer_capacity = n_act, hidden, batch_size, buffer_size
        cfg.gamma, cfg.tau = gamma, tau
        cfg.target_entropy = float(-n_act if target_entropy is None else target_entropy)
        cfg.seed, cfg.precision, cfg.device, cfg.rank, cfg.nranks = seed, precision, device, rank, nranks
        self._id_buf = None
        if nranks > 1:
            if nccl_id is None or len(nccl_id) != 128:
                raise ValueError("nranks > 1 needs the 128-byte nccl_id shared by all ranks")
            self._id_buf = C.create_string_buffer(bytes(nccl_id), 128)
            cfg.nccl_id = C.cast(self._id_buf, C.c_void_p)
            lib_path = _lib.default_nccl_lib()
            cfg.nccl_lib = lib_path.encode() if lib_path else None
        self.h = C.c_void_p()
        _lib.check(self.lib.b2g_sac_create(C.byref(cfg), C.byref(self.h)))
        self.obs_elems = int(np.prod(self.obs_shape))
        self._info = OrderedDict()
        name, numel, ndim = C.c_char_p(), C.c_int64(), C.c_int32()
        shape = (C.c_int64 * 4)()
        for i in range(self.lib.b2g_param_count(self.h)):
            _lib.check(self.lib.b2g_param_info(self.h, i, C.byref(name), C.byref(numel), C.byref(ndim), shape))
            self._info[name.value.decode()] = tuple(int(shape[k]) for k in range(ndim.value))

    # ---- lifetime
    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.b2g_sac_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- peer-memory data parallelism (include/b200grasp.h: b2g_sac_dp_export / b2g_sac_dp_connect)
    DP_EXPORT_BYTES = 192

    def dp_export(self) -> bytes:
        buf = C.create_string_buffer(self.DP_EXPORT_BYTES)
        _lib.check(self.lib.b2g_sac_dp_export(self.h, C.cast(buf, C.c_void_p)))
        return buf.raw

    def dp_connect(self, exports: "list[bytes]"):
        """exports: the dp_export() blob of every rank, in rank order.  From here on the optimiser launch of each step reduces
        the gradients, updates this rank's slice and writes the new parameters into every replica through NVLink peer memory."""
        blob = b"".join(exports)
        if len(blob) != self.DP_EXPORT_BYTES * len(exports):
            raise ValueError("every export blob must be DP_EXPORT_BYTES long")
        buf = C.create_string_buffer(blob, len(blob))
        _lib.check(self.lib.b2g_sac_dp_connect(self.h, C.cast(buf, C.c_void_p), len(exports)))

    def dp_connect_torch(self):
        """dp_export + all_gather over the initialised torch.distributed process group + dp_connect."""
        import torch.distributed as dist
        blobs = [None] * dist.get_world_size()
        dist.all_gather_object(blobs, self.dp_export())
        self.dp_connect(blobs)
        dist.barrier()

    @staticmethod
    def nccl_unique_id() -> bytes:
        lib = _lib.load()
        buf = C.create_string_buffer(128)
        p = _lib.default_nccl_lib()
        _lib.check(lib.b2g_nccl_unique_id(C.cast(buf, C.c_void_p), p.encode() if p else None))
        return buf.raw

    # ---- parameters (SB zip names / layouts)
    @property
    def param_shapes(self) -> "OrderedDict[str, tuple]":
        return self._info

    def get_parameters(self) -> "OrderedDict[str, np.ndarray]":
        out = OrderedDict()
        for n, shp in self._info.items():
            a = np.empty(shp, np.float32)
            _lib.check(self.lib.b2g_get_param(self.h, n.encode(), _fp(a.reshape(-1)), a.size))
            out[n] = a
        return out

    def load_parameters(self, params: Dict[str, np.ndarray], exact_match: bool = True):
        seen = set()
        for n, a in params.items():
            key = n[:-2] if n.endswith(":0") else n
            if key not in self._info:
                if exact_match:
                    raise ValueError(f"unknown variable {n}")
                continue
            a = _f32(a)
            if tuple(a.shape) != self._info[key]:
                raise ValueError(f"shape mismatch for {n}: {a.shape} vs {self._info[key]}")
            _lib.check(self.lib.b2g_set_param(self.h, key.encode(), _fp(a.reshape(-1)), a.size))
            seen.add(key)
        if exact_match and seen != set(self._info):
            raise ValueError(f"missing variables: {sorted(set(self._info) - seen)[:4]}...")

    def get_gradients(self) -> "OrderedDict[str, np.ndarray]":
        out = OrderedDict()
        for n, shp in self._info.items():
            if n.startswith("target/"):
                continue
            a = np.empty(shp, np.float32)
            _lib.check(self.lib.b2g_get_grad(self.h, n.encode(), _fp(a.reshape(-1)), a.size))
            out[n] = a
        return out

    def get_adam(self, name: str):
        shp = self._info[name]
        m, v = np.empty(shp, np.float32), np.empty(shp, np.float32)
        _lib.check(self.lib.b2g_get_adam(self.h, name.encode(), _fp(m.reshape(-1)), _fp(v.reshape(-1)), m.size))
        return m, v

    def reset_optimizer(self):
        _lib.check(self.lib.b2g_reset_optimizer(self.h))

    # ---- replay + normalisation
    def replay_add(self, obs, act, rew, next_obs, done):
        obs, next_obs, act = _f32(obs), _f32(next_obs), _f32(act)
        rew, done = _f32(np.reshape(rew, -1)), _f32(np.reshape(done, -1))
        n = rew.shape[0]
        assert obs.size == n * self.obs_elems and next_obs.size == obs.size and act.size == n * self.n_act
        _lib.check(self.lib.b2g_replay_add(self.h, _fp(obs), _fp(act), _fp(rew), _fp(next_obs), _fp(done), n))

    def replay_size(self) -> int:
        return int(self.lib.b2g_replay_size(self.h))

    def replay_get(self, slot: int) -> dict:
        """One stored (raw) transition, like ``ReplayBuffer.storage[slot]``."""
        obs, nxt = np.empty(self.obs_shape, np.float32), np.empty(self.obs_shape, np.float32)
        act, rew, done = np.empty(self.n_act, np.float32), np.empty(1, np.float32), np.empty(1, np.float32)
        _lib.check(self.lib.b2g_replay_get(self.h, int(slot), _fp(obs.reshape(-1)), _fp(act), _fp(rew), _fp(nxt.reshape(-1)), _fp(done)))
        return dict(obs=obs, act=act, rew=float(rew[0]), next_obs=nxt, done=float(done[0]))

    def last_batch(self) -> dict:
        """Replay slots, policy noise and per-sample outputs of the LAST gradient step (graph path included)."""
        B = self.batch_size
        idx = np.empty(B, np.int32)
        eps, ps, pi = np.empty((B, self.n_act), np.float32), np.empty((7, B), np.float32), np.empty((B, self.n_act), np.float32)
        _lib.check(self.lib.b2g_get_last_batch(self.h, idx.ctypes.data_as(C.POINTER(C.c_int32)), _fp(eps.reshape(-1)),
                                               _fp(ps.reshape(-1)), _fp(pi.reshape(-1))))
        out = dict(indices=idx, eps=eps, pi=pi)
        for i, k in enumerate(("q1", "q2", "v", "logp", "v_targ", "q1_pi", "q2_pi")):
            out[k] = ps[i].copy()
        return out

    def set_norm_stats(self, obs_mean=None, obs_var=None, ret_var=1.0, clip_obs=10.0, clip_reward=10.0, epsilon=1e-8,
                       norm_obs=True, norm_reward=True):
        dp = C.POINTER(C.c_double)
        if norm_obs:
            m = np.ascontiguousarray(obs_mean, np.float64).reshape(-1)
            v = np.ascontiguousarray(obs_var, np.float64).reshape(-1)
            assert m.size == self.obs_elems and v.size == self.obs_elems
            mp, vp = m.ctypes.data_as(dp), v.ctypes.data_as(dp)
        else:
            mp = vp = None
        _lib.check(self.lib.b2g_set_norm_stats(self.h, mp, vp, float(ret_var), float(clip_obs), float(clip_reward),
                                                float(epsilon), int(bool(norm_obs)), int(bool(norm_reward))))

    # ---- hot path
    def step(self, n_steps: int = 1, lr: float = 3e-4) -> dict:
        m = _lib.SacMetrics()
        _lib.check(self.lib.b2g_sac_step(self.h, n_steps, lr, C.byref(m)))
        return m.as_dict()

    def step_async(self, n_steps: int = 1, lr: float = 3e-4):
        _lib.check(self.lib.b2g_sac_step_async(self.h, n_steps, lr))

    def sync(self):
        _lib.check(self.lib.b2g_sync(self.h))

    def step_explicit(self, obs, act, rew, next_obs, done, eps, lr: float = 3e-4, apply_update: bool = True):
        B = self.batch_size
        obs, next_obs, act, eps = _f32(obs), _f32(next_obs), _f32(act), _f32(eps)
        rew, done = _f32(np.reshape(rew, -1)), _f32(np.reshape(done, -1))
        assert obs.size == B * self.obs_elems and act.size == B * self.n_act and eps.size == B * self.n_act and rew.size == B
        ps = np.empty((7, B), np.float32)
        pi = np.empty((B, self.n_act), np.float32)
        m = _lib.SacMetrics()
        _lib.check(self.lib.b2g_sac_step_explicit(self.h, _fp(obs), _fp(act), _fp(rew), _fp(next_obs), _fp(done), _fp(eps),
                                                   lr, int(apply_update), C.byref(m), _fp(ps), _fp(pi)))
        out = m.as_dict()
        for i, k in enumerate(("q1", "q2", "v", "logp", "v_targ", "q1_pi", "q2_pi")):
            out[k] = ps[i].copy()
        out["pi"] = pi
        return out

    def step_host_pipelined(self, obs, act, rew, next_obs, done, eps, lr: float = 3e-4):
        """Enqueue one step on a HOST batch (arrays must stay alive until the next call / flush; pinned float32
        arrays overlap the copy with the previous step).  Returns the previous step's losses, or None."""
        B = self.batch_size
        arrs = [_f32(obs), _f32(act), _f32(np.reshape(rew, -1)), _f32(next_obs), _f32(np.reshape(done, -1)), _f32(eps)]
        assert arrs[0].size == B * self.obs_elems and arrs[1].size == B * self.n_act and arrs[2].size == B
        self._pipe_keep = (getattr(self, "_pipe_keep", ()) + (arrs,))[-2:]   # host buffers stay alive while copies fly
        m, have = _lib.SacMetrics(), C.c_int(0)
        _lib.check(self.lib.b2g_sac_step_host_pipelined(self.h, *[_fp(a) for a in arrs], lr, C.byref(m), C.byref(have)))
        return m.as_dict() if have.value else None

    def pipeline_flush(self) -> dict:
        m = _lib.SacMetrics()
        _lib.check(self.lib.b2g_sac_pipeline_flush(self.h, C.byref(m)))
        return m.as_dict()

    def act(self, obs, deterministic: bool = True) -> np.ndarray:
        obs = _f32(obs).reshape(-1, self.obs_elems)
        out = np.empty((obs.shape[0], self.n_act), np.float32)
        _lib.check(self.lib.b2g_sac_act(self.h, _fp(obs), obs.shape[0], int(deterministic), _fp(out)))
        return out

    def launches_per_step(self) -> int:
        return int(self.lib.b2g_launches_per_step(self.h))

    def last_step_ms(self) -> float:
        return float(self.lib.b2g_last_step_ms(self.h))

    def profile_step(self, lr: float = 3e-4) -> "OrderedDict[str, float]":
        cap = 64
        names = (C.c_char_p * cap)()
        ms = (C.c_float * cap)()
        n = _lib.check(self.lib.b2g_profile_step(self.h, lr, names, ms, cap))
        return OrderedDict((names[i].decode(), float(ms[i])) for i in range(n))
