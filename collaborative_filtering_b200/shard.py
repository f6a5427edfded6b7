"""User sharding across the GPUs of one box.  Users are independent units of the path
(each compute_eigens() task reads only the shared read-only weights,
precompute_local_threads.cpp:25,120-123), so the data path needs no collective: a rank owns a set
of users, W is replicated once (NCCL broadcast), results stay on the producing GPU."""
from __future__ import annotations

import heapq

import numpy as np


def user_cost(deg) -> np.ndarray:
    """Cost model of one user: the eigensolve is cubic in the number of rated movies."""
    d = np.asarray(deg, dtype=np.float64)
    return d ** 3 + 64.0 * d ** 2


def lpt_assign(deg, n_shards: int) -> np.ndarray:
    """Longest-processing-time-first greedy assignment; returns owner[u] in [0, n_shards).
    Deterministic (ties by user index), identical on every rank."""
    cost = user_cost(deg)
    order = np.lexsort((np.arange(len(cost)), -cost))
    owner = np.empty(len(cost), dtype=np.int32)
    heap = [(0.0, s) for s in range(n_shards)]
    heapq.heapify(heap)
    for u in order:
        load, s = heapq.heappop(heap)
        owner[u] = s
        heapq.heappush(heap, (load + cost[u], s))
    return owner


def shard_users(deg, rank: int, world: int) -> np.ndarray:
    """0-based user indices owned by `rank` (ascending)."""
    return np.nonzero(lpt_assign(deg, world) == rank)[0]


def imbalance(deg, owner, n_shards: int) -> float:
    """max shard cost / mean shard cost."""
    cost = user_cost(deg)
    loads = np.bincount(owner, weights=cost, minlength=n_shards)
    return float(loads.max() / loads.mean())
