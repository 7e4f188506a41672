import os, sys
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
os.environ["B2G_NO_GRAPH"] = "1"
from tests.util import *
cfg, params, vn = load_case("sac_depth")
tr = b200grasp.synth.make_transitions(2048, vn["obs_mean"], vn["obs_var"])
for dbg in (0, 1, 2, 4, 8, 1|8, 1|2|8, 1|2|4|8):
    os.environ["B2G_TC_DEBUG"] = str(dbg)
    L = make_learner(cfg, vn, 256, params, buffer_size=2048, precision=1)
    L.replay_add(tr["obs"], tr["act"], tr["rew"], tr["next_obs"], tr["done"])
    L.step(3)
    pr = L.profile_step()
    print(f"dbg={dbg:2d} (noload={dbg&1} nomma={(dbg>>1)&1} nostore={(dbg>>2)&1} nosplit={(dbg>>3)&1})", {k: round(v*1e3) for k, v in pr.items() if k.startswith('conv') or k.startswith('fc1')}, flush=True)
    L.close()
