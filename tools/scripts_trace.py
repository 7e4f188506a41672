import os, sys
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from tests.util import *
cfg, params, vn = load_case("sac_depth")
tr = b200grasp.synth.make_transitions(2048, vn["obs_mean"], vn["obs_var"])
L = make_learner(cfg, vn, 256, params, buffer_size=2048, precision=1)
L.replay_add(tr["obs"], tr["act"], tr["rew"], tr["next_obs"], tr["done"])
L.step(5)
for g in os.environ.get("TRACE_GROUPS", "conv1_fwd,conv2_fwd,conv2_wgrad").split(","):
    os.environ["B2G_TC_TRACE"] = g
    L.profile_step()
