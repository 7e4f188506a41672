"""Times the BDQ learner (row a11): gradient steps/s through `BDQLearner.step` (device replay, losses read back) for the
two shipped network shapes (config/gripper_grasp.yaml:104-118: batch 64; BDQ_8pads 3x8 bins, BDQ_33pads_big 3x33 bins with
a 512/256 trunk), with the CPU oracle (oracle/bdq_ref.py, fp32, 16 threads) beside it."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200grasp  # noqa: E402
from oracle import bdq_ref as Q  # noqa: E402

out = {}
for name, cfg, B in (("bdq_8pads", Q.BDQConfig(100, 3, 8, (64, 64), 32, 32, 0.99), 64),
                     ("bdq_33pads_big", Q.BDQConfig(100, 3, 33, (512, 256), 128, 128, 0.99), 64)):
    rng = np.random.default_rng(0)
    n = 4096
    obs = rng.normal(0.4, 0.2, (n, cfg.obs_dim)).astype(np.float32)
    nxt = rng.normal(0.4, 0.2, (n, cfg.obs_dim)).astype(np.float32)
    act = rng.integers(0, cfg.n_bins, (n, cfg.n_branches)).astype(np.float32)
    rew = rng.choice([0.0, 1.0], n).astype(np.float32)
    done = (rng.random(n) < 0.2).astype(np.float32)
    L = b200grasp.BDQLearner(cfg.obs_dim, cfg.n_branches, cfg.n_bins, (cfg.trunk, (cfg.branch_hidden,), (cfg.value_hidden,)), batch_size=B,
                             buffer_size=n, gamma=cfg.gamma, target_network_update_freq=1000)
    L.load_parameters(Q.init_params(cfg, seed=1))
    L.replay_add(obs, act, rew, nxt, done)
    L.step(50)
    t0 = time.perf_counter()
    L.step(2000)
    gpu = 2000 / (time.perf_counter() - t0)
    t0 = time.perf_counter()
    for _ in range(200):
        L.step(1)                       # one call per step: what learn() does between env steps
    gpu1 = 200 / (time.perf_counter() - t0)
    torch.set_num_threads(16)
    params = Q.init_params(cfg, seed=1)
    opt = {"t": 0, "m": {}, "v": {}}
    bt = dict(obs=obs[:B], next_obs=nxt[:B], act_idx=act[:B].astype(np.int64), rew=rew[:B], done=done[:B])
    Q.bdq_step(params, opt, bt, 1e-4, cfg)
    t0 = time.perf_counter()
    k = 0
    while time.perf_counter() - t0 < 3.0:
        _, _, params, opt = Q.bdq_step(params, opt, bt, 1e-4, cfg)
        k += 1
    cpu = k / (time.perf_counter() - t0)
    out[name] = {"gpu_steps_per_s_batched_call": gpu, "gpu_steps_per_s_one_call_per_step": gpu1, "cpu_oracle_steps_per_s": cpu, "batch": B}
    L.close()
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/bdq_bench.json", "w"), indent=1)
