"""Phase timestamps (%globaltimer, ns) of the peer-memory optimiser kernel on every rank:
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29540 tools/dp_trace.py"""
import ctypes as C, os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200grasp
from b200grasp import synth, _lib

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("gloo")
ids = [b200grasp.Learner.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, 0)
GOLD = os.path.join(ROOT, "tests", "golden")
vn = dict(np.load(os.path.join(GOLD, "vecnorm_sac_depth.npz")))
params = dict(np.load(os.path.join(GOLD, "sac_depth_params.npz")))
L = b200grasp.Learner((64, 64, 2), n_act=5, batch_size=256, buffer_size=4096, seed=1, device=rank, rank=rank, nranks=world, nccl_id=ids[0], precision=1)
L.dp_connect_torch()
L.load_parameters(params)
L.set_norm_stats(vn["obs_mean"], vn["obs_var"], float(vn["ret_var"]), float(vn["clip_obs"]), float(vn["clip_reward"]), float(vn["epsilon"]))
tr = synth.make_transitions(2048, vn["obs_mean"], vn["obs_var"], seed=1 + rank)
L.replay_add(tr["obs"], tr["act"], tr["rew"], tr["next_obs"], tr["done"])
L.step(20)
for it in range(3):
    dist.barrier()
    L.step(50)
    st = (C.c_longlong * 5)()
    _lib.check(L.lib.b2g_debug_dp_stamps(L.h, st))
    t = [int(x) for x in st]
    print(f"rank {rank} it {it}: step {L.last_step_ms() / 50 * 1e3:.1f} us | start {t[0] % 10**9} | flags-in +{t[1]-t[0]} ns | slice done (CTA 0) +{t[2]-t[1]} | last CTA arrives +{t[3]-t[2]} | "
          f"peers' slices in +{t[4]-t[3]} | total {t[4]-t[0]} ns", flush=True)
L.close()
dist.barrier()
dist.destroy_process_group()
