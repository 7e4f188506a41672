"""PyTorch-CPU restatement of the reference's convolutional auto-encoder (encoder + decoder).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Follows
/root/reference/manipulation_main/gripperEnv/encoders.py:84-128 (Keras 2.2.4 on TF 1.14):

  encoder: for layer in network: Conv2D(filters, kernel_size, strides, padding='same') -> LeakyReLU(alpha)   (:91-96)
           Flatten (NHWC order) -> Dense(encoding_dim) -> LeakyReLU(alpha)                                 (:100-102)
  decoder: Dense(prod(shape)) -> LeakyReLU -> Reshape -> for i reversed: UpSampling2D(strides_i) ->
           Conv2D(filters_{i-1}, kernel_i, 'same') -> LeakyReLU; UpSampling2D -> Conv2D(1, kernel_0, 'same') (:110-124)

TensorFlow 'same' padding: out = ceil(in / s); pad_total = max((out - 1) * s + k - in, 0); floor(pad_total / 2) goes
in FRONT (top / left), the rest behind.  UpSampling2D is nearest-neighbour repetition.

PINNING: the WEIGHTS are the reference's (model.h5, read by h5min and committed as tests/golden/encoder_weights.npz);
the reference holds no stored encodings, and Keras/TF are not installable here, so outputs are PARITY UNPINNED except
for one sanity anchor: the restated auto-encoder's reconstruction error on depth-like synthetic scenes has the
magnitude of the shipped training history (history.csv val_loss ~1.2e-3) -- see tests/golden/make_fixtures.py.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


def same_pad(size: int, k: int, s: int) -> Tuple[int, int]:
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


def conv_same(x: torch.Tensor, w_hwio: torch.Tensor, b: torch.Tensor, s: int) -> torch.Tensor:
    """x [N,C,H,W]; Keras kernel [kh,kw,in,out]."""
    kh, kw = w_hwio.shape[:2]
    pt, pb = same_pad(x.shape[2], kh, s)
    pl, pr = same_pad(x.shape[3], kw, s)
    return F.conv2d(F.pad(x, (pl, pr, pt, pb)), w_hwio.permute(3, 2, 0, 1), b, stride=s)


def encode(imgs: np.ndarray, arrays: Sequence[Tuple[np.ndarray, np.ndarray]], strides: List[int], alpha: float = 0.1,
           dtype=torch.float32) -> np.ndarray:
    """imgs [N,H,W,C]; arrays = [(kernel, bias)] convs then dense.  Returns [N, encoding_dim]."""
    x = torch.tensor(np.asarray(imgs), dtype=dtype).permute(0, 3, 1, 2)
    for (k, b), s in zip(arrays[:-1], strides):
        x = F.leaky_relu(conv_same(x, torch.tensor(k, dtype=dtype), torch.tensor(b, dtype=dtype), s), alpha)
    flat = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)           # Keras Flatten on channels_last
    kd, bd = arrays[-1]
    return F.leaky_relu(flat @ torch.tensor(kd, dtype=dtype) + torch.tensor(bd, dtype=dtype), alpha).numpy()


def decode(z: np.ndarray, dec_arrays: Sequence[Tuple[np.ndarray, np.ndarray]], strides: List[int], shape_hwc, alpha: float = 0.1,
           dtype=torch.float32) -> np.ndarray:
    """dec_arrays = [(dense_2), (conv for i = L-1 .. 1), (final conv)] in model.h5 order.  Returns [N,H,W,1]."""
    kd, bd = dec_arrays[0]
    h = F.leaky_relu(torch.tensor(z, dtype=dtype) @ torch.tensor(kd, dtype=dtype) + torch.tensor(bd, dtype=dtype), alpha)
    h = h.reshape(-1, *shape_hwc).permute(0, 3, 1, 2)
    L = len(strides)
    for j, i in enumerate(reversed(range(1, L))):
        h = F.interpolate(h, scale_factor=strides[i], mode="nearest")
        k, b = dec_arrays[1 + j]
        h = F.leaky_relu(conv_same(h, torch.tensor(k, dtype=dtype), torch.tensor(b, dtype=dtype), 1), alpha)
    h = F.interpolate(h, scale_factor=strides[0], mode="nearest")
    k, b = dec_arrays[-1]
    return conv_same(h, torch.tensor(k, dtype=dtype), torch.tensor(b, dtype=dtype), 1).permute(0, 2, 3, 1).numpy()
