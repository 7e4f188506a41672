"""Times the perception encoder (row a12): per-env-step latency (n = 1) and batched throughput through the public
`SimpleAutoEncoder.encode` call (host buffers in, host buffers out), with the CPU oracle beside it."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200grasp  # noqa: E402,F401
from b200grasp import synth  # noqa: E402
from b200grasp.encoders import SimpleAutoEncoder, keras_encoder_arrays  # noqa: E402
from oracle import encoder_ref as E  # noqa: E402
from tests.test_encoder_cpu import load_fixture  # noqa: E402

w, cfg = load_fixture()
arr = keras_encoder_arrays(w, 3)
out = {}
for n in (1, 256, 4096):
    enc = SimpleAutoEncoder(cfg, max_batch=n)
    enc.set_weights(arr)
    imgs = synth.make_depth_scenes(n, seed=1)
    for _ in range(5):
        enc.encode(imgs)
    reps = 200 if n == 1 else 20
    t0 = time.perf_counter()
    for _ in range(reps):
        enc.encode(imgs)
    dt = (time.perf_counter() - t0) / reps
    torch.set_num_threads(16)
    E.encode(imgs[:min(n, 256)], arr, [2, 2, 2])
    t0 = time.perf_counter()
    E.encode(imgs[:min(n, 256)], arr, [2, 2, 2])
    ct = (time.perf_counter() - t0) * (n / min(n, 256))
    out[f"n{n}"] = {"gpu_ms": dt * 1e3, "gpu_frames_per_s": n / dt, "cpu_oracle_ms": ct * 1e3, "cpu_frames_per_s": n / ct}
    enc.close()
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/encoder_bench.json", "w"), indent=1)
