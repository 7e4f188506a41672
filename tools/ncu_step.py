"""Minimal driver for ncu: a depth-CNN learner (B=256, bf16x3 parity mode), a few warm-up steps, then STEPS graph replays.
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv python tools/ncu_step.py
  ncu --set full --clock-control none --import-source on -k regex:cg_kernel -s 40 -c 12 -o gpurun_out/cg_r2 python tools/ncu_step.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200grasp
from b200grasp import synth

GOLD = os.path.join(ROOT, "tests", "golden")
vn = dict(np.load(os.path.join(GOLD, "vecnorm_sac_depth.npz")))
params = dict(np.load(os.path.join(GOLD, "sac_depth_params.npz")))
L = b200grasp.Learner((64, 64, 2), n_act=5, batch_size=int(os.environ.get("BATCH", "256")), buffer_size=8192, seed=1, precision=1)
L.load_parameters(params)
L.set_norm_stats(vn["obs_mean"], vn["obs_var"], float(vn["ret_var"]), float(vn["clip_obs"]), float(vn["clip_reward"]), float(vn["epsilon"]))
for i in range(2):
    tr = synth.make_transitions(2048, vn["obs_mean"], vn["obs_var"], seed=1 + i)
    L.replay_add(tr["obs"], tr["act"], tr["rew"], tr["next_obs"], tr["done"])
L.step(int(os.environ.get("WARM", "2")))
m = L.step(int(os.environ.get("STEPS", "2")))
print(m["qf1_loss"], L.launches_per_step())
L.close()
