"""Partitioning of independent problems over ranks (ensemble mode):
contiguous batch ranges, no data-path collective - only statistics are reduced."""
from __future__ import annotations


def shard_range(n, rank, world):
    """[lo, hi) of rank `rank` when n items are split contiguously over `world` ranks,
    the first n % world ranks taking one extra item."""
    if world < 1 or not (0 <= rank < world) or n < 0:
        raise ValueError("shard_range: bad arguments")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def slab_range(nx, rank, world):
    """Row slab [i_lo, i_hi) of a single large grid split along the slow axis i
    (halo lines are contiguous ny*8 B): same contiguous rule as shard_range."""
    return shard_range(nx, rank, world)
