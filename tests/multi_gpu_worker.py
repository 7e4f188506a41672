"""2-rank data-parallel parity worker (launched by tests/test_gpu_multi.py or by hand:
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_worker.py).

Each rank takes its half of a seeded 2B batch through the C ABI; one NCCL all-reduce averages the
gradients; the result must equal the ORACLE's single step on the concatenated 2B batch
(SURVEY.md section 8e parity definition)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200grasp  # noqa: E402
from oracle import sac_ref as R  # noqa: E402
from tests.util import load_case, make_batch, rel_err  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    ids = [b200grasp.Learner.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, 0)
    cfg, params, vn = load_case("sac_depth")
    B = 16
    raw, norm, eps = make_batch(vn, B * world)
    prec = int(os.environ.get("PREC", "0"))
    L = b200grasp.Learner(cfg.obs_shape, n_act=cfg.n_act, batch_size=B, buffer_size=64, device=local, rank=rank, nranks=world,
                          nccl_id=ids[0], precision=prec)
    L.set_norm_stats(vn["obs_mean"], vn["obs_var"], float(vn["ret_var"]), float(vn["clip_obs"]), float(vn["clip_reward"]),
                     float(vn["epsilon"]))
    dp = os.environ.get("DP", "nccl")
    if dp == "p2p":
        L.dp_connect_torch()       # the optimiser launch becomes the collective (peer-memory reduce-scatter + Adam + all-gather)
    L.load_parameters(params)
    sl = slice(rank * B, (rank + 1) * B)
    out = L.step_explicit(raw["obs"][sl], raw["act"][sl], raw["rew"][sl], raw["next_obs"][sl], raw["done"][sl], eps[sl], lr=3e-4)
    ref, grads, newp, _ = R.sac_step(params, R.OptState.zeros(params), norm, eps, 3e-4, cfg, torch.float64)
    errs = {k: abs(out[k] - float(ref[k])) / abs(float(ref[k])) for k in
            ("policy_loss", "qf1_loss", "qf2_loss", "value_loss", "grad_norm_pi", "grad_norm_values")}
    if dp == "p2p":
        # the gradients are never materialised as a whole: compare the updated parameters (what the step produces) instead,
        # as the UPDATE each one received -- lr * m / (sqrt(v) + eps) of the first Adam step is +-lr wherever the gradient is not ~0
        newd = L.get_parameters()
        upd_ref = np.concatenate([(np.asarray(newp[n], np.float64) - np.asarray(params[n], np.float64)).reshape(-1) for n in grads])
        upd_dev = np.concatenate([(newd[n].astype(np.float64) - np.asarray(params[n], np.float64)).reshape(-1) for n in grads])
        big = np.concatenate([np.abs(np.asarray(grads[n], np.float64)).reshape(-1) for n in grads]) > 1e-7     # Adam's first step is sign-like
        gerr = float(np.linalg.norm((upd_dev - upd_ref)[big]) / np.linalg.norm(upd_ref[big]))
    else:
        g = L.get_gradients()
        gerr = max(rel_err(g[n], grads[n]) for n in grads)
    # replicas must stay bit-identical
    mine = np.concatenate([a.reshape(-1) for a in L.get_parameters().values()])
    allp = [None] * world
    dist.all_gather_object(allp, mine.tobytes())
    same = all(b == allp[0] for b in allp)
    q_err = rel_err(out["q1"], np.asarray(ref["q1"]).reshape(-1)[sl])
    print(f"rank {rank}: errs {errs} worst-grad {gerr:.2e} q1 {q_err:.2e} replicas_identical {same}", flush=True)
    ok = all(v <= 1e-4 for v in errs.values()) and gerr <= 1e-3 and same and q_err <= 1e-4
    L.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
