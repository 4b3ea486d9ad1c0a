"""Point-range sharding of one MSM across ranks.

Rank g owns ck[first .. first+count) resident in its HBM and receives the matching scalar slice; each
rank runs a full local Pippenger and contributes ONE Jacobian point (96 bytes).  Elliptic-curve
addition is not an NCCL reduction, so the partial sums are all-gathered (12 x int64 per rank) and
added on every rank's GPU (vimz_point_sum).  Independent transformations need no collective at all
(replicas, one prover context per GPU).
"""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of [0, n): the first n % world ranks get one extra element."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n, world)
    first = rank * base + min(rank, rem)
    count = base + (1 if rank < rem else 0)
    return first, count


def gather_partial_points(dist, local_point: np.ndarray, device=None) -> np.ndarray:
    """all_gather of one Jacobian point per rank -> (world, 12) uint64 array (identical on every rank)."""
    import torch
    world = dist.get_world_size()
    t = torch.from_numpy(np.ascontiguousarray(local_point, dtype=np.uint64).view(np.int64).copy())
    if device is not None:
        t = t.to(device)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    return torch.stack(parts).cpu().numpy().view(np.uint64).reshape(world, 12)


def sharded_commit(dist, n: int, local_commit: Callable[[int, int], np.ndarray], point_sum: Callable[[np.ndarray], np.ndarray],
                   device=None) -> np.ndarray:
    """commit over n points split by point range: local_commit(first, count) -> Jacobian point of this rank's
    shard; point_sum adds the gathered partial sums."""
    first, count = shard_range(n, dist.get_rank(), dist.get_world_size())
    local = local_commit(first, count)
    return point_sum(gather_partial_points(dist, local, device))
