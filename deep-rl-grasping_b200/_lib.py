"""ctypes binding of libb200grasp.so (the C ABI in include/b200grasp.h).

The product path has NO CPU fallback: if the shared library is missing this raises, and if it
loads on a box without an sm_100 GPU ``b2g_sac_create`` fails with B2G_ECUDA.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200grasp.so")

B2G_PREC_FP32_SIMT, B2G_PREC_BF16X3, B2G_PREC_BF16 = 0, 1, 2

#: every symbol include/b200grasp.h declares (tests check the .so exports each of them)
SYMBOLS = [
    "b2g_last_error", "b2g_version", "b2g_nccl_unique_id", "b2g_sac_create", "b2g_sac_destroy", "b2g_sac_dp_export", "b2g_sac_dp_connect", "b2g_debug_dp_stamps", "b2g_debug_compact_host", "b2g_sync",
    "b2g_param_count", "b2g_param_info", "b2g_get_param", "b2g_set_param", "b2g_get_grad", "b2g_get_adam",
    "b2g_reset_optimizer", "b2g_replay_add", "b2g_replay_size", "b2g_replay_get", "b2g_get_last_batch", "b2g_set_norm_stats", "b2g_sac_step",
    "b2g_sac_step_async", "b2g_sac_step_explicit", "b2g_sac_step_host_pipelined", "b2g_sac_pipeline_flush", "b2g_sac_act", "b2g_launches_per_step", "b2g_last_step_ms",
    "b2g_profile_step",
    "b2g_bdq_create", "b2g_bdq_destroy", "b2g_bdq_param_count", "b2g_bdq_param_info", "b2g_bdq_get_param", "b2g_bdq_set_param",
    "b2g_bdq_get_grad", "b2g_bdq_replay_add", "b2g_bdq_replay_size", "b2g_bdq_set_norm_stats", "b2g_bdq_step",
    "b2g_bdq_step_explicit", "b2g_bdq_act", "b2g_bdq_set_per_beta", "b2g_bdq_get_last_per",
    "b2g_encoder_create", "b2g_encoder_destroy", "b2g_encoder_n_layers", "b2g_encoder_layer_shape", "b2g_encoder_set_weights",
    "b2g_encoder_encode", "b2g_debug_gemm",
]

ENC_MAX_LAYERS = 8


class EncoderCfg(C.Structure):
    _fields_ = [
        ("height", C.c_int32), ("width", C.c_int32), ("channels", C.c_int32), ("n_layers", C.c_int32),
        ("filters", C.c_int32 * ENC_MAX_LAYERS), ("kernel", C.c_int32 * ENC_MAX_LAYERS), ("strides", C.c_int32 * ENC_MAX_LAYERS),
        ("encoding_dim", C.c_int32), ("alpha", C.c_float), ("max_batch", C.c_int32), ("device", C.c_int32),
    ]


class BdqCfg(C.Structure):
    _fields_ = [
        ("obs_dim", C.c_int32), ("n_branches", C.c_int32), ("n_bins", C.c_int32), ("trunk0", C.c_int32), ("trunk1", C.c_int32),
        ("branch_hidden", C.c_int32), ("batch", C.c_int32), ("buffer_capacity", C.c_int64), ("gamma", C.c_float),
        ("target_update_freq", C.c_int32), ("trunk_grad_rescale", C.c_int32), ("seed", C.c_uint64), ("device", C.c_int32),
        ("rank", C.c_int32), ("nranks", C.c_int32), ("nccl_id", C.c_void_p), ("nccl_lib", C.c_char_p),
        ("prioritized_replay", C.c_int32), ("per_alpha", C.c_float), ("per_eps", C.c_float),
    ]


class BdqMetrics(C.Structure):
    _fields_ = [("loss", C.c_float), ("mean_q", C.c_float), ("grad_norm", C.c_float), ("n_updates", C.c_int64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class SacCfg(C.Structure):
    _fields_ = [
        ("obs_h", C.c_int32), ("obs_w", C.c_int32), ("obs_c", C.c_int32), ("obs_dim", C.c_int32),
        ("n_act", C.c_int32), ("hidden", C.c_int32), ("batch", C.c_int32), ("buffer_capacity", C.c_int64),
        ("gamma", C.c_float), ("tau", C.c_float), ("target_entropy", C.c_float), ("seed", C.c_uint64),
        ("precision", C.c_int32), ("device", C.c_int32), ("rank", C.c_int32), ("nranks", C.c_int32),
        ("nccl_id", C.c_void_p), ("nccl_lib", C.c_char_p),
    ]


class SacMetrics(C.Structure):
    _fields_ = [(n, C.c_float) for n in (
        "policy_loss", "qf1_loss", "qf2_loss", "value_loss", "ent_coef_loss", "entropy", "ent_coef",
        "grad_norm_pi", "grad_norm_values", "grad_ent", "mean_q1", "mean_q2", "mean_v", "mean_logp")] + \
        [("n_updates", C.c_int64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class B2GError(RuntimeError):
    pass


_lib = None


def load():
    """Loads libb200grasp.so (building it is ``__graft_entry__.build()`` / ``build.sh``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B2GError(f"{LIB_PATH} is missing: run ./build.sh (there is no CPU fallback for the learner path)")
    lib = C.CDLL(LIB_PATH)
    fp, dp, vp = C.POINTER(C.c_float), C.POINTER(C.c_double), C.c_void_p
    lib.b2g_last_error.restype = C.c_char_p
    lib.b2g_nccl_unique_id.argtypes = [vp, C.c_char_p]
    lib.b2g_sac_create.argtypes = [C.POINTER(SacCfg), C.POINTER(vp)]
    lib.b2g_sac_destroy.argtypes = [vp]
    lib.b2g_sac_dp_export.argtypes = [vp, vp]
    lib.b2g_sac_dp_connect.argtypes = [vp, vp, C.c_int]
    lib.b2g_debug_dp_stamps.argtypes = [vp, C.POINTER(C.c_longlong)]
    lib.b2g_debug_compact_host.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.b2g_sync.argtypes = [vp]
    lib.b2g_param_count.argtypes = [vp]
    lib.b2g_param_info.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                                   C.POINTER(C.c_int64)]
    for f in ("b2g_get_param", "b2g_set_param", "b2g_get_grad"):
        getattr(lib, f).argtypes = [vp, C.c_char_p, fp, C.c_size_t]
    lib.b2g_get_adam.argtypes = [vp, C.c_char_p, fp, fp, C.c_size_t]
    lib.b2g_reset_optimizer.argtypes = [vp]
    lib.b2g_replay_add.argtypes = [vp, fp, fp, fp, fp, fp, C.c_int64]
    lib.b2g_replay_size.argtypes = [vp]
    lib.b2g_replay_size.restype = C.c_int64
    lib.b2g_replay_get.argtypes = [vp, C.c_int64, fp, fp, fp, fp, fp]
    lib.b2g_get_last_batch.argtypes = [vp, C.POINTER(C.c_int32), fp, fp, fp]
    lib.b2g_set_norm_stats.argtypes = [vp, dp, dp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]
    lib.b2g_sac_step.argtypes = [vp, C.c_int, C.c_float, C.POINTER(SacMetrics)]
    lib.b2g_sac_step_async.argtypes = [vp, C.c_int, C.c_float]
    lib.b2g_sac_step_explicit.argtypes = [vp, fp, fp, fp, fp, fp, fp, C.c_float, C.c_int, C.POINTER(SacMetrics), fp, fp]
    lib.b2g_sac_step_host_pipelined.argtypes = [vp, fp, fp, fp, fp, fp, fp, C.c_float, C.POINTER(SacMetrics), C.POINTER(C.c_int)]
    lib.b2g_sac_pipeline_flush.argtypes = [vp, C.POINTER(SacMetrics)]
    lib.b2g_sac_act.argtypes = [vp, fp, C.c_int, C.c_int, fp]
    lib.b2g_launches_per_step.argtypes = [vp]
    lib.b2g_last_step_ms.argtypes = [vp]
    lib.b2g_last_step_ms.restype = C.c_float
    lib.b2g_profile_step.argtypes = [vp, C.c_float, C.POINTER(C.c_char_p), fp, C.c_int]
    lib.b2g_bdq_create.argtypes = [C.POINTER(BdqCfg), C.POINTER(vp)]
    lib.b2g_bdq_destroy.argtypes = [vp]
    lib.b2g_bdq_param_count.argtypes = [vp]
    lib.b2g_bdq_param_info.argtypes = [vp, C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
    for f in ("b2g_bdq_get_param", "b2g_bdq_set_param", "b2g_bdq_get_grad"):
        getattr(lib, f).argtypes = [vp, C.c_char_p, fp, C.c_size_t]
    lib.b2g_bdq_replay_add.argtypes = [vp, fp, fp, fp, fp, fp, C.c_int64]
    lib.b2g_bdq_replay_size.argtypes = [vp]
    lib.b2g_bdq_replay_size.restype = C.c_int64
    lib.b2g_bdq_set_norm_stats.argtypes = [vp, dp, dp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]
    lib.b2g_bdq_step.argtypes = [vp, C.c_int, C.c_float, C.POINTER(BdqMetrics)]
    lib.b2g_bdq_step_explicit.argtypes = [vp, fp, fp, fp, fp, fp, fp, C.c_float, C.c_int, C.POINTER(BdqMetrics), fp]
    lib.b2g_bdq_act.argtypes = [vp, fp, C.c_int, C.POINTER(C.c_int32)]
    lib.b2g_bdq_set_per_beta.argtypes = [vp, C.c_float]
    lib.b2g_bdq_get_last_per.argtypes = [vp, C.POINTER(C.c_int32), fp, fp]
    lib.b2g_encoder_create.argtypes = [C.POINTER(EncoderCfg), C.POINTER(vp)]
    lib.b2g_encoder_destroy.argtypes = [vp]
    lib.b2g_encoder_n_layers.argtypes = [vp]
    lib.b2g_encoder_layer_shape.argtypes = [vp, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.b2g_encoder_set_weights.argtypes = [vp, C.c_int, fp, C.c_size_t, fp, C.c_size_t]
    lib.b2g_encoder_encode.argtypes = [vp, fp, C.c_int, fp]
    lib.b2g_debug_gemm.argtypes = [C.c_int, C.c_int, C.c_int, fp, fp, fp, C.c_int, C.c_int]
    _lib = lib
    return lib


def check(rc: int):
    if rc < 0:
        raise B2GError(f"libb200grasp error {rc}: {load().b2g_last_error().decode(errors='replace')}")
    return rc


def default_nccl_lib():
    """Prefer the NCCL bundled with torch (2.28.9) over the system one (2.27.3)."""
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec and spec.submodule_search_locations:
            p = os.path.join(list(spec.submodule_search_locations)[0], "lib", "libnccl.so.2")
            if os.path.exists(p):
                return p
    except Exception:
        pass
    return None
