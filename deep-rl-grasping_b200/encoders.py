"""Host mirror of the reference's perception encoder.

Same names and call shapes as /root/reference/manipulation_main/gripperEnv/encoders.py:
``SimpleAutoEncoder(config)`` (:67-136), ``load_weights(model_dir)`` (:27-31), ``encode(imgs)`` (:59-61),
``encoding_shape`` (:63-65).  ``encode`` runs on the GPU through libb200grasp (csrc/encoder.cu); the decoder half
(``predict``, only used for the OpenCV debug view at sensor.py:223-228) and ``train``/``test``/``plot`` are outside
the scope table and raise ``NotImplementedError``.  ``model.h5`` is read by ``h5min`` (no h5py/keras needed).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List

import numpy as np

from . import _lib, h5min


def keras_encoder_arrays(weights: Dict[str, np.ndarray], n_conv: int):
    """Orders Keras auto-named layers the way the reference builds them: encoder convs are the first ``n_conv``
    ``conv2d_*`` layers, the encoder's Dense is the first ``dense_*`` layer (encoders.py:92-104 precede :110-128).
    Returns [(kernel, bias)] for conv_0..conv_{n-1}, dense."""
    def ordered(prefix):
        idx = sorted({int(k.split("/")[0].split("_")[-1]) for k in weights if k.startswith(prefix)})
        return [f"{prefix}{i}" for i in idx]
    convs, denses = ordered("conv2d_"), ordered("dense_")
    if len(convs) < n_conv or not denses:
        raise ValueError(f"model.h5 holds {len(convs)} conv / {len(denses)} dense layers; config needs {n_conv} / 1")
    return [(weights[f"{n}/kernel"], weights[f"{n}/bias"]) for n in convs[:n_conv] + denses[:1]]


class Encoder(object):
    """Base class for learning abstract representations of image observations."""

    def __init__(self, config, max_batch: int = 1, device: int = 0):
        self._handle = C.c_void_p()
        self._max_batch, self._device = int(max_batch), int(device)
        self._lib = _lib.load()
        self._build(config)

    def _build(self, config):
        raise NotImplementedError

    def train(self, *a, **k):
        raise NotImplementedError("auto-encoder training is outside the hot-path scope (DESIGN.md section 1)")

    test = plot = train

    def predict(self, imgs):
        raise NotImplementedError("the decoder half is not built (only used for the reference's debug view, sensor.py:223)")


class SimpleAutoEncoder(Encoder):
    """Vanilla autoencoder -- encoder half."""

    input_shape = (64, 64, 1)          # encoders.py:87

    def _build(self, config):
        self.network: List[dict] = list(config["network"])
        self.encoding_dim = int(config["encoding_dim"])
        self.alpha = float(config.get("alpha", 0.1))
        cfg = _lib.EncoderCfg()
        cfg.height, cfg.width, cfg.channels = self.input_shape
        cfg.n_layers = len(self.network)
        if cfg.n_layers > _lib.ENC_MAX_LAYERS:
            raise ValueError("too many conv layers")
        for i, layer in enumerate(self.network):
            cfg.filters[i], cfg.kernel[i], cfg.strides[i] = int(layer["filters"]), int(layer["kernel_size"]), int(layer["strides"])
        cfg.encoding_dim, cfg.alpha = self.encoding_dim, self.alpha
        cfg.max_batch, cfg.device = self._max_batch, self._device
        _lib.check(self._lib.b2g_encoder_create(C.byref(cfg), C.byref(self._handle)))

    def set_weights(self, arrays):
        """arrays: [(kernel, bias)] for each conv then the dense layer (Keras layouts)."""
        n = self._lib.b2g_encoder_n_layers(self._handle)
        if len(arrays) != n:
            raise ValueError(f"expected {n} (kernel, bias) pairs, got {len(arrays)}")
        fp = C.POINTER(C.c_float)
        for i, (k, b) in enumerate(arrays):
            k = np.ascontiguousarray(k, np.float32)
            b = np.ascontiguousarray(b, np.float32)
            _lib.check(self._lib.b2g_encoder_set_weights(self._handle, i, k.ctypes.data_as(fp), k.size, b.ctypes.data_as(fp), b.size))

    def load_weights(self, model_dir):
        model_dir = os.path.expanduser(model_dir)
        weights = h5min.load_keras_weights(os.path.join(model_dir, "model.h5"))
        self.set_weights(keras_encoder_arrays(weights, len(self.network)))

    def encode(self, imgs):
        imgs = np.ascontiguousarray(imgs, np.float32)
        if imgs.ndim != 4 or imgs.shape[1:] != self.input_shape:
            raise ValueError(f"expected imgs of shape (n, {self.input_shape}), got {imgs.shape}")
        n = imgs.shape[0]
        out = np.empty((n, self.encoding_dim), np.float32)
        fp = C.POINTER(C.c_float)
        for s in range(0, n, self._max_batch):          # keras predict() batches internally too (batch_size=32)
            e = min(n, s + self._max_batch)
            _lib.check(self._lib.b2g_encoder_encode(self._handle, imgs[s:e].ctypes.data_as(fp), e - s, out[s:e].ctypes.data_as(fp)))
        return out

    @property
    def encoding_shape(self):
        return (self.encoding_dim,)

    def close(self):
        if self._handle:
            self._lib.b2g_encoder_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
