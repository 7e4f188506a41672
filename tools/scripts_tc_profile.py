import os, sys
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
os.environ["B2G_NO_GRAPH"] = "1"
from tests.util import *
prec = int(os.environ.get("PREC", "1"))
cfg, params, vn = load_case("sac_depth")
L = make_learner(cfg, vn, 256, params, buffer_size=2048, precision=prec)
tr = b200grasp.synth.make_transitions(2048, vn["obs_mean"], vn["obs_var"])
L.replay_add(tr["obs"], tr["act"], tr["rew"], tr["next_obs"], tr["done"])
print(L.step(int(os.environ.get("NSTEPS", "3"))))
