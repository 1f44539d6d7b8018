"""Independent-session batching across GPUs: the recursion does not shard, so N GPUs
run N filters ("replicas only"); the one collective is an all-gather of each session's 8-double pose
record (t, x, y, z, qw, qx, qy, qz) after a vision update."""
from __future__ import annotations

import torch
import torch.distributed as dist


def session_seed(n_features: int, rank: int) -> int:
    """Synthetic-sequence seed of session `rank` (BASELINE config 5: seeds 1000 + N + rank)."""
    return 1000 + n_features + rank


def gather_pose_records(rec: torch.Tensor) -> torch.Tensor:
    """All-gather one (8,) float64 record per rank into a (world, 8) tensor on the same device.
    NCCL when `rec` is a CUDA tensor (64 B per rank over NVLink: latency-bound), gloo on CPU."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    flat = torch.empty(world * rec.numel(), dtype=rec.dtype, device=rec.device)
    if world == 1:
        flat.copy_(rec.reshape(-1))
    else:
        dist.all_gather_into_tensor(flat, rec.contiguous().reshape(-1))
    return flat.view(world, rec.numel())
