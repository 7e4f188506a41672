import sys, time, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import tests.test_gpu_parity as T
from tests.util import *
for name in ["test_encoder_mlp_trained_weights_b64", "test_depth_cnn_trained_weights_b32", "test_rgbd_cnn_fresh_init_b16", "test_depth_cnn_fresh_init_b256", "test_two_steps_optimizer_state", "test_sampled_step_from_replay_and_policy_act"]:
    t = time.time()
    try:
        getattr(T, name)(); print("PASS", name, f"{time.time()-t:.1f}s", flush=True)
    except Exception as e:
        import traceback; traceback.print_exc(); print("FAIL", name, str(e)[:2000], flush=True)
# timing
cfg, params, vn = load_case("sac_depth")
B = 256
raw, norm, eps = make_batch(vn, 4096)
L = make_learner(cfg, vn, B, params, buffer_size=4096)
L.replay_add(raw["obs"], raw["act"], raw["rew"], raw["next_obs"], raw["done"])
L.step(5)
for n in (20, 100):
    m = L.step(n); print(n, "steps:", L.last_step_ms()/n, "ms/step", m)
print(L.profile_step())
