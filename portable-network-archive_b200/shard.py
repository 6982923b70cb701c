"""Entry sharding across the GPUs of one box: entries are independent units (own IV, own zstd/zlib stream,
own chunks; the key is shared read-only), so a rank takes a subset of the entries and there is NO data-path collective.
Greedy longest-processing-time on compressed bytes; deterministic, so every rank computes the same partition from the
index pass alone (the index is 12 bytes per chunk and is replicated, not communicated)."""
from __future__ import annotations

import heapq


def lpt_partition(weights, world: int):
    """Returns `world` lists of entry indices.  Heaviest first onto the currently lightest rank; ties by index so the
    result does not depend on dict/heap ordering.  Each list is returned in ascending entry order (archive order)."""
    if world <= 0:
        raise ValueError("world must be positive")
    order = sorted(range(len(weights)), key=lambda i: (-int(weights[i]), i))
    heap = [(0, r) for r in range(world)]
    heapq.heapify(heap)
    parts = [[] for _ in range(world)]
    for i in order:
        load, r = heapq.heappop(heap)
        parts[r].append(i)
        heapq.heappush(heap, (load + int(weights[i]), r))
    return [sorted(p) for p in parts]


def rank_entries(weights, rank: int, world: int):
    return lpt_partition(weights, world)[rank]
