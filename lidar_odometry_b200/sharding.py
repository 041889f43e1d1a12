"""Host-side logic of the two multi-GPU modes.  Pure Python/torch.distributed plumbing: no geometry here.

* independent sequences: rank r runs sequence r on its own GPU; nothing is exchanged, times are max-reduced.
* point-sharded scan: rank r owns the contiguous query slice ``shard_bounds(m, world, r)``; per Gauss-Newton iteration the
  ranks exchange (count, sum r, sum r^2), the <=128-entry GMM sample and the 28 normal-equation sums.
"""
from __future__ import annotations

import math


def shard_bounds(m: int, world: int, rank: int):
    """Contiguous slice [lo, hi) of m queries for `rank`; the first m % world ranks get one extra query, so the
    concatenation over ranks is the original order (the PKO sample is drawn by position in that order)."""
    base, extra = divmod(m, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_plan(counts, rank: int):
    """Global offset of this rank's accepted correspondences and the global count, from the all-gathered counts."""
    counts = [int(c) for c in counts]
    return sum(counts[:rank]), sum(counts)


def scale_from_moments(total: int, sum_r: float, sum_r2: float) -> float:
    """Residual normalisation of iteration 0 (ICP.cpp:304-316): population sigma / 6 from the global raw moments."""
    mean = sum_r / total
    return math.sqrt(max(sum_r2 / total - mean * mean, 0.0)) / 6.0


def sequence_seed(base_seed: int, rank: int) -> int:
    return base_seed + rank


def whole_job_rate(units_per_rank: int, world: int, max_seconds: float) -> float:
    """Units processed by all ranks divided by the slowest rank's time."""
    return world * units_per_rank / max_seconds
