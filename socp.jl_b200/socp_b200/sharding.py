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


def _parse_cpulist(text: str) -> List[int]:
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def numa_node_of_pci_device(bus_id: str, sysfs: str = "/sys") -> int:
    """NUMA node of a PCI device ("0000:3b:00.0"), -1 when the platform does not say."""
    import os
    try:
        with open(os.path.join(sysfs, "bus", "pci", "devices", bus_id.lower(), "numa_node")) as f:
            return int(f.read().strip())
    except (OSError, ValueError):
        return -1


def bind_host_to_pci_device(bus_id: str, sysfs: str = "/sys") -> int:
    """One process per GPU: run this process on the CPUs of its GPU's NUMA node, so that the pinned staging buffers it
    allocates afterwards (first touch) sit on the socket the GPU hangs off -- with eight ranks uploading at once the
    host side of the H2D copies otherwise crosses the socket interconnect.  Returns the node, or -1 when nothing was
    changed (single-node host, unknown topology, no permission)."""
    import os
    node = numa_node_of_pci_device(bus_id, sysfs)
    if node < 0:
        return -1
    try:
        with open(os.path.join(sysfs, "devices", "system", "node", f"node{node}", "cpulist")) as f:
            cpus = set(_parse_cpulist(f.read()))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return -1
        os.sched_setaffinity(0, allowed)
    except (OSError, ValueError, AttributeError):
        return -1
    return node
