"""Batch sharding across ranks / devices: problems are independent
(reference src/solver.jl:40-46 touches only its own Problem / SolverState), so a batch is cut into
contiguous shards and no collective is needed on the data path.  The same rule is used by the
library for the devices of one process (socp_b200_create: ceil(B / ndev) problems per device) and
by bench.py for one-process-per-GPU runs."""
from __future__ import annotations

from typing import List, Tuple


def shard_range(batch: int, world: int, rank: int) -> Tuple[int, int]:
    """[first, last) of `rank`'s contiguous shard of a batch of `batch` problems."""
    assert 0 <= rank < world and batch >= 0
    per = (batch + world - 1) // world
    first = min(batch, rank * per)
    return first, min(batch, first + per)


def weak_shard(per_rank: int, rank: int) -> Tuple[int, int]:
    """Weak scaling (bench.py): every rank owns `per_rank` problems of the seeded sequence."""
    return rank * per_rank, (rank + 1) * per_rank


def gather_plan(batch: int, world: int) -> List[Tuple[int, int]]:
    """All shards, in rank order; they tile [0, batch) exactly."""
    return [shard_range(batch, world, r) for r in range(world)]
