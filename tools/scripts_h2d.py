import sys, time, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from tests.util import *
cfg, params, vn = load_case("sac_depth")
B = 256
L = make_learner(cfg, vn, B, params, buffer_size=4096, precision=1)
tr = b200grasp.synth.make_transitions(B, vn["obs_mean"], vn["obs_var"])
eps = b200grasp.synth.make_eps(B)
pin = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory().numpy() for k, v in dict(tr, eps=eps).items()}
x = torch.empty(16789504 // 4, dtype=torch.float32).pin_memory(); y = torch.empty_like(x, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(50): y.copy_(x, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 50
print("torch pinned H2D 16.8MB: %.3f ms = %.1f GB/s" % (dt * 1e3, 16.79e6 / dt / 1e9))
t0 = time.perf_counter()
for _ in range(50): L.replay_add(pin["obs"], pin["act"], pin["rew"], pin["next_obs"], pin["done"])
dt = (time.perf_counter() - t0) / 50
print("replay_add 256 transitions (16.8MB): %.3f ms" % (dt * 1e3))
L.step(5)
for n in (1, 2):
    t0 = time.perf_counter()
    for _ in range(100): L.step(n)
    print("step(%d) sync loop: %.3f ms/call" % (n, (time.perf_counter() - t0) / 100 * 1e3))
for _ in range(3): L.step_host_pipelined(pin["obs"], pin["act"], pin["rew"], pin["next_obs"], pin["done"], pin["eps"])
L.pipeline_flush()
t0 = time.perf_counter()
for _ in range(100): L.step_host_pipelined(pin["obs"], pin["act"], pin["rew"], pin["next_obs"], pin["done"], pin["eps"])
L.pipeline_flush(); print("pipelined: %.3f ms/step" % ((time.perf_counter() - t0) / 100 * 1e3))
