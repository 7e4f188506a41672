"""Perception encoder forward on the GPU, through the C ABI, against oracle/encoder_ref.py with
the reference's own weights (tests/golden/encoder_weights.npz <- encoder_files/new_gripper_encoder/model.h5)."""
import os

import numpy as np
import pytest
import torch

import b200grasp  # noqa: F401
from b200grasp import _lib, synth
from b200grasp.encoders import SimpleAutoEncoder, keras_encoder_arrays
from oracle import encoder_ref as E
from tests.test_encoder_cpu import load_fixture
from tests.util import GOLD

pytestmark = pytest.mark.gpu

TOL = 1e-4      # relative to the largest encoding magnitude (fp32 FFMA vs float64 oracle)


def _enc(max_batch):
    w, cfg = load_fixture()
    enc = SimpleAutoEncoder(cfg, max_batch=max_batch)
    arr = keras_encoder_arrays(w, len(cfg["network"]))
    enc.set_weights(arr)
    return enc, arr, cfg


def test_golden_encodings():
    enc, _, _ = _enc(8)
    g = np.load(os.path.join(GOLD, "golden_encoder.npz"))
    z = enc.encode(synth.make_depth_scenes(8, seed=11))
    assert z.shape == (8, 100) and z.dtype == np.float32
    assert np.abs(z - g["z"]).max() <= TOL * np.abs(g["z"]).max()
    assert enc.encoding_shape == (100,)


@pytest.mark.parametrize("n", [1, 3, 64, 100])
def test_ragged_batches_match_oracle(n):
    enc, arr, cfg = _enc(64)                       # n = 100 exercises the chunked path, n = 1 the per-env-step call
    rng = np.random.default_rng(n)
    imgs = synth.make_depth_scenes(n, seed=n) + rng.normal(0, 0.02, (n, 64, 64, 1)).astype(np.float32)   # exercises every border pixel
    ref = E.encode(imgs, arr, [2, 2, 2], cfg["alpha"], torch.float64)
    z = enc.encode(imgs)
    assert np.abs(z - ref).max() <= TOL * np.abs(ref).max()
    # results must not depend on what an earlier, larger call left in the bordered buffers
    z1 = enc.encode(imgs[:1])
    assert np.array_equal(z1[0], z[0])


def test_other_geometry_matches_oracle():
    """odd kernel/stride mix incl. stride 1 and a 4-channel first layer is outside the shipped config but inside the API."""
    cfg = {"network": [{"filters": 8, "kernel_size": 3, "strides": 1}, {"filters": 16, "kernel_size": 4, "strides": 2},
                       {"filters": 12, "kernel_size": 5, "strides": 3}], "encoding_dim": 20, "alpha": 0.2}
    rng = np.random.default_rng(5)
    arr, c, hw = [], 1, 64
    for l in cfg["network"]:
        arr.append((rng.normal(0, 0.2, (l["kernel_size"], l["kernel_size"], c, l["filters"])).astype(np.float32),
                    rng.normal(0, 0.1, l["filters"]).astype(np.float32)))
        c, hw = l["filters"], -(-hw // l["strides"])
    arr.append((rng.normal(0, 0.05, (hw * hw * c, 20)).astype(np.float32), rng.normal(0, 0.1, 20).astype(np.float32)))
    enc = SimpleAutoEncoder(cfg, max_batch=5)
    enc.set_weights(arr)
    imgs = rng.normal(0, 1, (5, 64, 64, 1)).astype(np.float32)
    ref = E.encode(imgs, arr, [1, 2, 3], 0.2, torch.float64)
    z = enc.encode(imgs)
    assert np.abs(z - ref).max() <= TOL * np.abs(ref).max()


def test_errors_are_loud():
    _, cfg = load_fixture()
    enc = SimpleAutoEncoder(cfg, max_batch=2)
    with pytest.raises(_lib.B2GError, match="no weights"):
        enc.encode(np.zeros((1, 64, 64, 1), np.float32))
    with pytest.raises(ValueError):
        enc.encode(np.zeros((1, 32, 32, 1), np.float32))
    with pytest.raises(_lib.B2GError, match="expected kernel numel"):
        enc.set_weights([(np.zeros((3, 3, 1, 32), np.float32), np.zeros(32, np.float32))] * 4)
    with pytest.raises(NotImplementedError):
        enc.predict(np.zeros((1, 64, 64, 1), np.float32))
