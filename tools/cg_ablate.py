"""Role ablation of the v2 engine launches: per-group device time of one profiled step with operand loads (bit 0), MMAs
(bit 1) and epilogue stores (bit 2) switched off through B2G_CG_DEBUG (results are garbage; timing only)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200grasp
from b200grasp import synth

GOLD = os.path.join(ROOT, "tests", "golden")
vn = dict(np.load(os.path.join(GOLD, "vecnorm_sac_depth.npz")))
params = dict(np.load(os.path.join(GOLD, "sac_depth_params.npz")))
B = int(os.environ.get("BATCH", "256"))
tr = synth.make_transitions(2048, vn["obs_mean"], vn["obs_var"], seed=1)
rows = {}
for dbg in [int(x) for x in os.environ.get("DBGS", "0,1,2,4,3,5,6,7").split(",")]:
    os.environ["B2G_CG_DEBUG"] = str(dbg)
    L = b200grasp.Learner((64, 64, 2), n_act=5, batch_size=B, buffer_size=4096, seed=1, precision=1)
    L.load_parameters(params)
    L.set_norm_stats(vn["obs_mean"], vn["obs_var"], float(vn["ret_var"]), float(vn["clip_obs"]), float(vn["clip_reward"]), float(vn["epsilon"]))
    L.replay_add(tr["obs"], tr["act"], tr["rew"], tr["next_obs"], tr["done"])
    L.step(5)
    acc = None
    for _ in range(5):
        p = L.profile_step()
        acc = p if acc is None else {k: min(acc[k], v) for k, v in p.items()}
    L.step(20)
    ms = L.last_step_ms() / 20
    rows[dbg] = (acc, ms)
    L.close()
names = list(rows[next(iter(rows))][0].keys())
print("%-18s" % "group" + "".join("%9s" % ("dbg=%d" % d) for d in rows))
for n in names:
    print("%-18s" % n + "".join("%9.1f" % (rows[d][0][n] * 1e3) for d in rows))
print("%-18s" % "graph step (us)" + "".join("%9.1f" % (rows[d][1] * 1e3) for d in rows))
