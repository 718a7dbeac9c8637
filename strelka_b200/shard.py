"""Region sharding across ranks: contiguous, equal-count blocks of loci; rank r owns [begin, end).
Regions are independent given the pre-rolled context the caller already includes in each region's reference window, exactly how
the reference splits genome segments across processes (applications/starling/starling_run.cpp:335-341)."""
from __future__ import annotations

from typing import List, Tuple


def shard_range(n_units: int, rank: int, world: int) -> Tuple[int, int]:
    assert 0 <= rank < world
    return (n_units * rank) // world, (n_units * (rank + 1)) // world


def shard_ranges(n_units: int, world: int) -> List[Tuple[int, int]]:
    return [shard_range(n_units, r, world) for r in range(world)]


def gathered_offsets(counts: List[int]) -> List[int]:
    """Record offset of each rank's block in the gathered array (rank order == locus order, so no re-sort is needed)."""
    off, out = 0, []
    for c in counts:
        out.append(off)
        off += c
    return out
