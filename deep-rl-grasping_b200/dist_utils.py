"""Host-side helpers of the data-parallel learner (one process per GPU, torch.distributed for the
control plane only; the gradient all-reduce itself is issued by libb200grasp on its own stream)."""
from __future__ import annotations

import os
from typing import Callable, Optional


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def share_nccl_id(make_id: Callable[[], bytes]) -> Optional[bytes]:
    """Rank 0 creates the 128-byte ncclUniqueId, every rank receives it (works on gloo and nccl groups)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return None
    box = [make_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(box, 0)
    assert isinstance(box[0], (bytes, bytearray)) and len(box[0]) == 128
    return bytes(box[0])


def shard_seed(seed: int, rank: int) -> int:
    """Distinct replay-index / policy-noise streams per rank (mirrors the rank mixing in csrc/sac.cu)."""
    return (seed + 0x9E3779B97F4A7C15 * rank) & 0xFFFFFFFFFFFFFFFF


def weak_scaling_value(sync_steps_per_s: float, world: int) -> float:
    """Metric definition: every rank processes its own 256-sample minibatch per
    synchronous step, so the job advances `world` 256-sample step-equivalents per synchronous step."""
    return sync_steps_per_s * world
