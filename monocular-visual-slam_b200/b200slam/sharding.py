"""Multi-GPU plumbing: frame pairs are independent, so ranks take contiguous blocks of
the pair batch and the only collective is one all-gather of fixed-size result records
(SURVEY.md §8e).  Works with NCCL (one process per B200) and with gloo on CPU (tests).
"""
from __future__ import annotations

import numpy as np

RECORD_WIDTH = 4  # per pair: (n_matches, best_h, inlier_count, pair_id)


def shard_bounds(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`; sizes differ by at most one."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_records(pair_ids, n_matches, best_h, inlier_count):
    """-> (n_local, RECORD_WIDTH) int32 array."""
    return np.stack([np.asarray(n_matches, np.int32), np.asarray(best_h, np.int32),
                     np.asarray(inlier_count, np.int32), np.asarray(pair_ids, np.int32)], axis=1)


def gather_records(local, n_items: int, group=None):
    """All-gather per-pair result records (torch tensor, (n_local, W) int32, on the
    backend's device) from every rank; returns the (n_items, W) tensor ordered by pair id.
    Blocks are padded to the largest shard so a single fixed-size collective suffices."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    cap = max(shard_bounds(n_items, r, world)[1] - shard_bounds(n_items, r, world)[0] for r in range(world))
    width = local.shape[1]
    send = torch.full((cap, width), -1, dtype=local.dtype, device=local.device)
    send[: local.shape[0]] = local
    recv = torch.empty((world * cap, width), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(n_items, r, world)
        parts.append(recv[r * cap: r * cap + (hi - lo)])
    return torch.cat(parts, dim=0)
