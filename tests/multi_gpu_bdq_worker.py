"""2-rank BDQ data-parallel parity worker (tests/test_gpu_bdq.py; BASELINE config 4)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200grasp  # noqa: E402
from oracle import bdq_ref as Q  # noqa: E402
from tests.util import rel_err  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    ids = [b200grasp.Learner.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, 0)
    cfg = Q.BDQConfig(100, 5, 33, (64, 64), 32, 32, 0.99)
    B = 32
    params = Q.init_params(cfg, seed=1)
    rng = np.random.default_rng(3)
    bt = dict(obs=rng.normal(0.4, 0.2, (B * world, cfg.obs_dim)).astype(np.float32), next_obs=rng.normal(0.4, 0.2, (B * world, cfg.obs_dim)).astype(np.float32),
              act_idx=rng.integers(0, cfg.n_bins, (B * world, cfg.n_branches)), rew=rng.choice([0.0, 1.0], B * world).astype(np.float32),
              done=(rng.random(B * world) < 0.2).astype(np.float32))
    L = b200grasp.BDQLearner(cfg.obs_dim, cfg.n_branches, cfg.n_bins, (cfg.trunk, (cfg.branch_hidden,), (cfg.value_hidden,)), batch_size=B,
                             buffer_size=256, gamma=cfg.gamma, target_network_update_freq=1000, device=local, rank=rank, nranks=world, nccl_id=ids[0])
    L.load_parameters(params)
    sl = slice(rank * B, (rank + 1) * B)
    out = L.step_explicit(bt["obs"][sl], bt["act_idx"][sl].astype(np.float32), bt["rew"][sl], bt["next_obs"][sl], bt["done"][sl], lr=1e-3)
    ref, grads, newp, _ = Q.bdq_step(params, {"t": 0, "m": {}, "v": {}}, bt, 1e-3, cfg, torch.float64)
    g = L.get_gradients()
    gerr = max(rel_err(g[n] / world, grads[n]) for n in grads)
    loss_err = abs(out["loss"] - ref["loss"]) / abs(ref["loss"])
    # sampled (graph) steps with the collective inside the captured graph
    L.replay_add(bt["obs"][sl], bt["act_idx"][sl].astype(np.float32), bt["rew"][sl], bt["next_obs"][sl], bt["done"][sl])
    m = L.step(3, lr=1e-3)
    mine = np.concatenate([a.reshape(-1) for a in L.get_parameters().values()])
    allp = [None] * world
    dist.all_gather_object(allp, mine.tobytes())
    same = all(b == allp[0] for b in allp)
    print(f"rank {rank}: loss_err {loss_err:.2e} worst-grad {gerr:.2e} replicas_identical {same} n_updates {m['n_updates']}", flush=True)
    ok = loss_err <= 1e-4 and gerr <= 1e-3 and same and m["n_updates"] == 4 and np.isfinite(m["loss"])
    L.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
