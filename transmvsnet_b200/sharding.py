"""Reference-view sharding across the GPUs of one box.

A dataset item is one (scan, ref_view, src_views) tuple and nothing crosses items
(models/TransMVSNet.py:141-226), so the path is embarrassingly parallel over reference views:
one process per GPU, no collective inside the path.  The only exchange is gathering the
per-view depth / confidence maps on rank 0 (the inference analogue of the reference's
DistributedSampler, train.py:377-381, which exists for training only).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_views(n_views: int, rank: int, world_size: int) -> List[int]:
    """Round-robin assignment: view v goes to rank v % world_size (balanced to within one view)."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return list(range(rank, n_views, world_size))


def views_per_rank(n_views: int, world_size: int) -> List[int]:
    return [len(range(r, n_views, world_size)) for r in range(world_size)]


def gather_maps(local_maps: torch.Tensor, n_views: int, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """Gather per-view maps onto rank `dst` in global view order.

    local_maps: [n_local, K, H, W] (K = 2: depth, confidence) for this rank's views, in shard order.
    Returns [n_views, K, H, W] on rank dst, None elsewhere.  Ranks with fewer views pad to the
    maximum so the collective has equal message sizes (NCCL/gloo all_gather).
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = views_per_rank(n_views, world)
    n_max = max(counts)
    if local_maps.shape[0] != counts[rank]:
        raise ValueError(f"rank {rank} holds {local_maps.shape[0]} views, expected {counts[rank]}")
    pad = local_maps
    if local_maps.shape[0] < n_max:
        fill = torch.zeros((n_max - local_maps.shape[0],) + tuple(local_maps.shape[1:]), dtype=local_maps.dtype,
                           device=local_maps.device)
        pad = torch.cat([local_maps, fill], 0)
    pad = pad.contiguous()
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    if rank != dst:
        return None
    out = torch.empty((n_views,) + tuple(local_maps.shape[1:]), dtype=local_maps.dtype, device=local_maps.device)
    for r in range(world):
        for j, v in enumerate(range(r, n_views, world)):
            out[v] = bufs[r][j]
    return out
