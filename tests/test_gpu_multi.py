"""N>1 path on real GPUs (skipped on a 1-GPU box; the CPU-side sharding logic is covered by
tests/test_host_cpu.py with gloo)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("prec", [0, 1])
def test_two_rank_allreduce_matches_oracle_on_concatenated_batch(prec):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, PREC=str(prec))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(29511 + prec), os.path.join(ROOT, "tests", "multi_gpu_worker.py")],
                       env=env, capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0


def test_two_rank_peer_memory_optimiser_matches_oracle():
    """b2g_sac_dp_connect: the optimiser launch reduces the gradients, updates each rank's slice and writes the new parameters into
    every replica over NVLink peer memory.  The updated parameters and the averaged losses must match the oracle's single step on the
    concatenated batch, and the replicas must be bit-identical (tests/multi_gpu_worker.py, DP=p2p)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, PREC="1", DP="p2p")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29515", os.path.join(ROOT, "tests", "multi_gpu_worker.py")],
                       env=env, capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
