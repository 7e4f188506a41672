import os, sys, time, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from tests.util import *
from oracle import sac_ref as R
cfg, params, vn = load_case("sac_depth")
B = int(os.environ.get("BISECT_B", "32"))
raw, norm, eps = make_batch(vn, B)
ref64, grads64, newp64, _ = R.sac_step(params, R.OptState.zeros(params), norm, eps, 3e-4, cfg, torch.float64)
groups = ["none", "conv1_fwd", "conv2_fwd", "fc1_fwd", "fc1_wgrad", "fc1_dgrad", "conv3_wgrad", "conv3_dgrad", "conv2_wgrad", "conv2_dgrad", "conv1_wgrad", "all"]
for prec in (1, 2):
    for gsel in groups:
        os.environ["B2G_TC_GROUPS"] = gsel
        try:
            L = make_learner(cfg, vn, B, params, precision=prec)
            out = L.step_explicit(raw["obs"], raw["act"], raw["rew"], raw["next_obs"], raw["done"], eps, lr=3e-4, apply_update=False)
            g = L.get_gradients()
            gerr = {n: rel_err(g[n], grads64[n]) for n in grads64}
            worst = sorted(gerr, key=gerr.get)[-3:]
            print(f"prec={prec} tc={gsel:12s} q1 {rel_err(out['q1'], ref64['q1'].reshape(-1)):.2e} v {rel_err(out['v'], ref64['v'].reshape(-1)):.2e} logp {rel_err(out['logp'], ref64['logp'].reshape(-1)):.2e} vt {rel_err(out['v_targ'], ref64['v_targ'].reshape(-1)):.2e} gn_pi {abs(out['grad_norm_pi']-ref64['grad_norm_pi'])/ref64['grad_norm_pi']:.2e} gn_v {abs(out['grad_norm_values']-ref64['grad_norm_values'])/ref64['grad_norm_values']:.2e} worst:", [(w.replace('model/',''), f"{gerr[w]:.1e}") for w in worst], flush=True)
            L.close()
        except Exception as e:
            print("prec", prec, gsel, "EXC", str(e)[:300], flush=True)
os.environ["B2G_TC_GROUPS"] = "all"
for prec in (0, 1, 2):
    L = make_learner(cfg, vn, 256, params, buffer_size=4096, precision=prec)
    tr = b200grasp.synth.make_transitions(4096, vn["obs_mean"], vn["obs_var"])
    L.replay_add(tr["obs"], tr["act"], tr["rew"], tr["next_obs"], tr["done"])
    L.step(10)
    m = L.step(100); print("prec", prec, L.last_step_ms()/100, "ms/step", {k: round(v, 5) for k, v in m.items() if k in ("qf1_loss","policy_loss","value_loss")})
    print({k: round(v, 4) for k, v in L.profile_step().items()})
    L.close()
