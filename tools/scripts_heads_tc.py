import os, sys, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from tests.util import *
from oracle import sac_ref as R
cfg, params, vn = load_case("sac_depth")
B = 256
raw, norm, eps = make_batch(vn, B)
ref64, grads64, _, _ = R.sac_step(params, R.OptState.zeros(params), norm, eps, 3e-4, cfg, torch.float64)
tr = b200grasp.synth.make_transitions(4096, vn["obs_mean"], vn["obs_var"])
for heads in ("0", "1"):
    os.environ["B2G_TC_HEADS"] = heads
    L = make_learner(cfg, vn, B, params, buffer_size=4096, precision=1)
    out = L.step_explicit(raw["obs"], raw["act"], raw["rew"], raw["next_obs"], raw["done"], eps, lr=3e-4, apply_update=False)
    g = L.get_gradients()
    gerr = {n: rel_err(g[n], grads64[n]) for n in grads64}
    w = sorted(gerr, key=gerr.get)[-3:]
    print("heads_tc", heads, "q1", rel_err(out["q1"], ref64["q1"].reshape(-1)), "logp", rel_err(out["logp"], ref64["logp"].reshape(-1)), "gn", abs(out["grad_norm_pi"]-ref64["grad_norm_pi"])/ref64["grad_norm_pi"], abs(out["grad_norm_values"]-ref64["grad_norm_values"])/ref64["grad_norm_values"], [(x, f"{gerr[x]:.1e}") for x in w])
    L.load_parameters(params)
    L.replay_add(tr["obs"], tr["act"], tr["rew"], tr["next_obs"], tr["done"])
    L.step(10); m = L.step(200)
    print("  ms/step", L.last_step_ms()/200, {k: round(v*1e3) for k, v in L.profile_step().items()})
    L.close()
