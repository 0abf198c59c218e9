"""Query-chunk sharding across GPUs.

The unit of work is one SeedAndFilter call = one (strand, 250 kb chunk) of a query block
(src/seeder.cpp:48-51,:89-90).  Units are independent: the reference hands each one
to whichever GPU is free (src/seed_filter.cu:699-708).  With one process per GPU the same
independence lets a static partition give byte-identical output; no data-path collective exists.
"""
from __future__ import annotations


def shard_units(num_units: int, rank: int, world_size: int) -> range:
    """Contiguous, balanced slice of [0, num_units) owned by `rank` (sizes differ by <= 1)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(num_units, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def shard_intervals(intervals: list, rank: int, world_size: int) -> list:
    """north_star's 'query intervals shard one-per-GPU': round-robin over the 10 Mb intervals
    (src/main.cpp:380-393), which keeps every rank busy for the same number of rounds."""
    return [iv for i, iv in enumerate(intervals) if i % world_size == rank]


def merge_ranked(per_rank: list) -> list:
    """Concatenate per-rank [(unit_index, payload)] lists back into unit order."""
    out = [x for part in per_rank for x in part]
    out.sort(key=lambda t: t[0])
    return out
