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


# ---- per-movie variant (local_calc.cpp): the independent unit is the MOVIE vertex -- its local graph, its eigensolve and
# all predictions for its test users live together (vertex_program::apply, local_calc.cpp:262-526) ---------------------
def movie_cost(n_nodes, n_pairs) -> np.ndarray:
    """Cost model of one movie vertex: a full eigensolve of its local graph (n nodes) plus, per test user, about twenty
    Lanczos steps of one n x n GEMV each (the fast path of gsi_local_calc_host).  Movies without pairs cost nothing."""
    n = np.asarray(n_nodes, dtype=np.float64)
    p = np.asarray(n_pairs, dtype=np.float64)
    return np.where(p > 0, 9.0 * n ** 3 + 40.0 * n ** 2 * p, 0.0)


def shard_movies(n_nodes, n_pairs, world: int) -> np.ndarray:
    """owner[movie] in [0, world): LPT over movie_cost, deterministic, identical on every rank."""
    cost = movie_cost(n_nodes, n_pairs)
    order = np.lexsort((np.arange(len(cost)), -cost))
    owner = np.zeros(len(cost), dtype=np.int32)
    heap = [(0.0, s) for s in range(world)]
    heapq.heapify(heap)
    for m in order:
        if cost[m] == 0.0:
            break
        load, s = heapq.heappop(heap)
        owner[m] = s
        heapq.heappush(heap, (load + cost[m], s))
    return owner


def local_calc_pair_mask(items, owner, rank: int, base_mask=None) -> np.ndarray:
    """pair_mask for gsi_local_calc_host on `rank`: the pairs whose movie this rank owns (and that `base_mask`, e.g. the
    --pct sample, keeps).  Every rank passes the whole test CSR, so a user's ratings of the other movies stay visible as
    known ratings; outputs of the other ranks' pairs come back as GSI_PRED_SKIPPED and are merged by the host."""
    items = np.asarray(items)
    mask = (np.asarray(owner)[items] == rank).astype(np.uint8)
    if base_mask is not None:
        mask &= np.asarray(base_mask, dtype=np.uint8)
    return mask
