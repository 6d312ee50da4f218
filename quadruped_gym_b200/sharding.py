"""Env-index sharding across the GPUs of one box (one process per GPU, launched by torchrun).

Environments are independent (the reference's own parallelism is ``SubprocVecEnv`` over OS processes,
/root/reference/src/train_quadruped.py:49-50), so the physics path needs NO collective: rank r owns the
contiguous global env ids ``[r*n_local, (r+1)*n_local)`` and passes ``env_offset`` to the library so
per-env random streams (reset yaw) are keyed on the GLOBAL id -- results do not depend on the world size.
The only collective is an optional all-reduce of a small rollout-statistics vector.
"""
from __future__ import annotations

import os
from typing import Dict, Tuple

import torch
import torch.distributed as dist


def rank_world() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice of global env ids owned by `rank` (sizes differ by at most one)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of size {world}")
    base, rem = divmod(n_total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


STAT_KEYS = ("physics_steps", "contacts", "efc_rows", "newton_iters", "ls_evals", "verts_tested", "diverged",
             "contact_overflow", "episodes", "reward_sum", "env_steps")


def reduce_rollout_stats(stats: Dict[str, float], device=None) -> Dict[str, float]:
    """Sum a small statistics dict over all ranks (NCCL on GPUs, gloo on CPU). No-op without a process group."""
    vec = torch.tensor([float(stats.get(k, 0.0)) for k in STAT_KEYS], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM)
    return {k: float(v) for k, v in zip(STAT_KEYS, vec.tolist())}


def max_over_ranks(seconds: float, device=None) -> float:
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
