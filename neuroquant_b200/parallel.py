"""Host-side partitioning for the multi-GPU paths (one process per GPU, torch.distributed):

  * calibration   frames of every global mini-batch are sharded across ranks; ONE all-reduce (sum) of the
                  flat dW/db buffer per iteration; the loss is normalised by the GLOBAL pixel count, so the
                  sum of rank gradients equals the reference's single-process gradient
  * bit_assign    candidate bit configurations are farmed out round-robin, one all-gather of scalars
  * decode        contiguous frame ranges per rank, no collective on the data path

These helpers are pure index arithmetic so that the CPU (gloo) tests exercise exactly what the GPU
path uses.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


def world_info(group=None):
    """(rank, world_size, group) -- (0, 1, None) outside torch.distributed."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group), group
    return 0, 1, None


def shard_indices(idx: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """This rank's frames of one global mini-batch: positions rank, rank + world, ...."""
    if world == 1:
        return idx
    if idx.numel() % world:
        raise ValueError(f"global batch of {idx.numel()} frames does not split over {world} ranks")
    return idx[rank::world]


def candidates_of_rank(n_candidates: int, rank: int, world: int) -> List[int]:
    """bit_assign farming: candidate i is scored by rank i % world."""
    return list(range(rank, n_candidates, world))


def frame_range_of_rank(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """decode sharding: contiguous ranges, sizes differing by at most one frame."""
    base, rem = divmod(n_frames, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_scores(local: Sequence[Tuple[int, float]], n_candidates: int, group=None) -> List[float]:
    """All ranks receive the full score list (one all_gather_object of a few floats)."""
    import torch.distributed as dist

    rank, world, group = world_info(group)
    if world == 1:
        pairs = list(local)
    else:
        bucket = [None] * world
        dist.all_gather_object(bucket, list(local), group=group)
        pairs = [p for part in bucket for p in part]
    out = [float("nan")] * n_candidates
    for i, s in pairs:
        out[i] = s
    return out
