"""Data-parallel exchange helpers (host logic, backend-agnostic so the gloo tests cover them).

Training is batch-sharded: rank r owns samples [r*B, (r+1)*B); BatchNorm statistics stay local
(the reference has no SyncBN); the only exchange per step is a sum all-reduce of the flat
gradient arena, cut into a few contiguous buckets so the tail of the arena (decoder + belly,
whose gradients are final first) can be reduced while the encoder's backward still runs.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def bucket_ranges(param_sizes: Sequence[int], n_buckets: int = 2) -> List[Tuple[int, int]]:
    """Contiguous [start, end) element ranges over the flat arena, cut at parameter boundaries,
    roughly equal in size, returned in arena order."""
    total = sum(param_sizes)
    if n_buckets <= 1 or len(param_sizes) <= 1:
        return [(0, total)]
    bounds, acc, target = [0], 0, total / n_buckets
    for s in param_sizes[:-1]:
        acc += s
        if acc >= target * len(bounds) and len(bounds) < n_buckets:
            bounds.append(acc)
    bounds.append(total)
    return [(a, b) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]


def allreduce_buckets(flat: torch.Tensor, ranges: Sequence[Tuple[int, int]], group=None, reverse: bool = True):
    """Sum all-reduce each bucket (last bucket first: that is the order backward finishes them)."""
    order = list(ranges)[::-1] if reverse else list(ranges)
    works = [dist.all_reduce(flat[a:b], group=group, async_op=True) for a, b in order]
    for w in works:
        w.wait()
    return flat
