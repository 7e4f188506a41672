"""Diagnostic: where does the policy-gradient-norm error of the bf16x3 engine come from?  Takes the first batch the
graph path draws (test_gpu_graph_path setup), replays it through step_explicit on the fp32 and bf16x3 engines and
prints gradient-norm and per-tensor gradient errors against the float64 oracle."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200grasp import synth
from oracle import sac_ref as R
from tests.util import load_case, make_learner, rel_err
from tests.test_gpu_graph_path import _norm_batch, _run_graph_steps, LR

cfg, params, vn = load_case("sac_depth")
B, NS = 256, 4096
tr = synth.make_transitions(NS, vn["obs_mean"], vn["obs_var"], seed=9001)
rows, _ = _run_graph_steps(cfg, params, vn, tr, B, 1, precision=1)
m, lb = rows[0]
idx = lb["indices"].astype(np.int64)
norm = _norm_batch(tr, idx, vn)
ref, grads, _, _ = R.sac_step(params, R.OptState.zeros(params), norm, lb["eps"], LR, cfg, torch.float64)
print("graph  bf16x3: gn_pi err %.2e gn_v err %.2e" % (abs(m["grad_norm_pi"] - float(ref["grad_norm_pi"])) / float(ref["grad_norm_pi"]),
                                                       abs(m["grad_norm_values"] - float(ref["grad_norm_values"])) / float(ref["grad_norm_values"])))
for prec in (0, 1):
    L = make_learner(cfg, vn, B, params, precision=prec)
    out = L.step_explicit(tr["obs"][idx], tr["act"][idx], tr["rew"][idx], tr["next_obs"][idx], tr["done"][idx], lb["eps"], lr=LR, apply_update=False)
    g = L.get_gradients()
    print("explicit prec=%d: gn_pi err %.2e gn_v err %.2e" % (prec, abs(out["grad_norm_pi"] - float(ref["grad_norm_pi"])) / float(ref["grad_norm_pi"]),
                                                              abs(out["grad_norm_values"] - float(ref["grad_norm_values"])) / float(ref["grad_norm_values"])))
    tot = np.sqrt(sum(float((grads[n] ** 2).sum()) for n in grads if n.startswith("model/pi/")))
    for n in grads:
        if n.startswith("model/pi/"):
            gn, rn = float(np.linalg.norm(g[n])), float(np.linalg.norm(grads[n]))
            print("   %-28s |g| %.3e (share %.3f)  rel_err %.2e  norm_err %+.2e" % (n, rn, rn / tot, rel_err(g[n], grads[n]), (gn - rn) / rn))
    L.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez(os.path.join(ROOT, "gpurun_out", "diag_batch.npz"), idx=idx, eps=lb["eps"])
