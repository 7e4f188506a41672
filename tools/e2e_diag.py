"""Device timeline of b2g_sac_step_host_pipelined (B2G_PIPE_TRACE=1 python tools/e2e_diag.py)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200grasp
from b200grasp import synth
GOLD = os.path.join(ROOT, "tests", "golden")
vn = dict(np.load(os.path.join(GOLD, "vecnorm_sac_depth.npz")))
params = dict(np.load(os.path.join(GOLD, "sac_depth_params.npz")))
B = 256
L = b200grasp.Learner((64, 64, 2), n_act=5, batch_size=B, buffer_size=4096, seed=1, precision=1)
L.load_parameters(params)
L.set_norm_stats(vn["obs_mean"], vn["obs_var"], float(vn["ret_var"]), float(vn["clip_obs"]), float(vn["clip_reward"]), float(vn["epsilon"]))
tr = synth.make_transitions(B, vn["obs_mean"], vn["obs_var"], seed=77)
eps = synth.make_eps(B, seed=78)
pin = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory().numpy() for k, v in dict(tr, eps=eps).items()}
for _ in range(3):
    L.step_host_pipelined(pin["obs"], pin["act"], pin["rew"], pin["next_obs"], pin["done"], pin["eps"])
L.pipeline_flush()
n = int(os.environ.get("N", "60"))
torch.cuda.synchronize()
t0 = time.perf_counter()
ts = []
for _ in range(n):
    L.step_host_pipelined(pin["obs"], pin["act"], pin["rew"], pin["next_obs"], pin["done"], pin["eps"])
    ts.append(time.perf_counter())
L.pipeline_flush()
torch.cuda.synchronize()
el = time.perf_counter() - t0
print(f"{n / el:.1f} steps/s; host call-to-call ms: {np.round(np.diff([t0] + ts) * 1e3, 3).tolist()}", file=sys.stderr)
L.close()
