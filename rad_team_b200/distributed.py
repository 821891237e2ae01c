"""Multi-GPU plumbing: one process per GPU, environments sharded in contiguous blocks, no cross-GPU
traffic in `step`.  The only exchange is once per rollout: every rank contributes six doubles —
its reading moments (count, mean, M2) and its episode statistics (sum of returns, sum of lengths,
agents that reached the source) — to ONE all-gather (NCCL over NVLink on GPUs, gloo in the CPU tests),
and every rank then merges the rows in rank order, so all ranks hold bit-identical results
(SURVEY.md §5 / §8e).  Moments are never all-reduce-summed as (sum x, sum x^2): that cancels
catastrophically and its order is not deterministic.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

PACK = 6  # count, mean, M2, sum_return, sum_len, n_done


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [base, base + count) of global env ids owned by `rank`; the first
    n_total % world ranks own one extra env."""
    if not (0 <= rank < world) or n_total < 0:
        raise ValueError("bad shard request")
    q, r = divmod(n_total, world)
    count = q + (1 if rank < r else 0)
    base = rank * q + min(rank, r)
    return base, count


def pack_stats(moments_state: torch.Tensor, episode_stats: torch.Tensor) -> torch.Tensor:
    """[6] float64 on the tensors' device; no host round trip."""
    return torch.cat([moments_state.reshape(3), episode_stats.reshape(3)]).contiguous()


def gather_stats(packed: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All-gather the per-rank [6] rows into [world, 6] (row r = rank r).  Works for any backend."""
    if not (dist.is_available() and dist.is_initialized()):
        return packed.reshape(1, PACK).clone()
    world = dist.get_world_size(group)
    out = torch.empty((world, PACK), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out.view(-1), packed.contiguous(), group=group)
    return out


def merge_rollout_stats(moments, episode_stats: torch.Tensor, group: Optional[dist.ProcessGroup] = None):
    """Once per rollout on every rank: gather, then rank-ordered Chan merge of the moments ON THE
    DEVICE (rt_moments_merge) and an in-order sum of the episode statistics.  Returns
    (gathered [world, 6], merged episode stats [3]); `moments.state` now holds the global moments."""
    gathered = gather_stats(pack_stats(moments.state, episode_stats), group)
    moments.merge_gathered(gathered)
    ep = gathered[0, 3:].clone()
    for r in range(1, gathered.shape[0]):  # fixed order; integer-valued doubles, exact anyway
        ep += gathered[r, 3:]
    return gathered, ep
