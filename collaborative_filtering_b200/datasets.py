"""Synthetic MovieLens / Netflix shaped inputs and the user-fold splitter (host side, numpy).

Recipe: SURVEY.md section 8d (seed 31413 = the reference generator's own seed,
make_synthetic_als_data.cpp:125).  User degrees are log-normal, clipped and rescaled to hit the
rating count; items are drawn without replacement proportionally to a Zipf(1.0) popularity over a
random permutation of item ids 1..I; ratings are clip(round(3.5 + b_u + b_i + N(0,1)), 1, 5).
Movie ids start at 1 (row/col 0 of the weight table is unused, precompute_local.cpp:153).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

SEED = 31413

#            users   items   nnz         lo  hi     med  sigma  half_stars
SHAPES = {
    "ml-100k": (943, 1682, 100_000, 20, 737, 65, 0.989, False),
    "ml-1m": (6040, 3706, 1_000_209, 20, 2314, 96, 1.044, False),
    "ml-10m": (71_567, 10_681, 10_000_054, 20, 7359, 69, 1.207, True),
    "netflix": (480_189, 17_770, 100_480_507, 1, 17_653, 96, 1.248, False),
}


@dataclass
class Ratings:
    """CSR by user.  user ids are 1..U (file ids; user' = INT_MAX - id downstream)."""

    n_users: int
    n_items: int
    offsets: np.ndarray   # int64 [U+1]
    items: np.ndarray     # int32 [nnz], ascending inside a user
    ratings: np.ndarray   # float32 [nnz]

    @property
    def nnz(self) -> int:
        return int(self.offsets[-1])

    def degrees(self) -> np.ndarray:
        return np.diff(self.offsets)

    def triples(self):
        """(user, item, rating) rows, user ids 1..U."""
        u = np.repeat(np.arange(1, self.n_users + 1, dtype=np.int64), self.degrees())
        return u, self.items.astype(np.int64), self.ratings


def _degrees(rng, n_users, nnz, lo, hi, med, sigma):
    raw = np.exp(rng.normal(np.log(med), sigma, size=n_users))
    scale = 1.0
    for _ in range(60):  # rescale so that the clipped, rounded degrees sum to nnz
        d = np.clip(np.rint(raw * scale), lo, hi)
        s = d.sum()
        if abs(s - nnz) <= 0.0005 * nnz:
            break
        scale *= nnz / s
    return d.astype(np.int64)


def _draw_items(rng, deg, n_items, p):
    """Successive sampling without replacement proportional to p == i.i.d. draws from p keeping
    first occurrences.  Vectorised over users; heavy users fall back to exponential races."""
    n_users = len(deg)
    cdf = np.cumsum(p)
    cdf[-1] = 1.0
    chosen = [None] * n_users
    todo = np.arange(n_users)
    factor = 1.6
    for _round in range(4):
        if len(todo) == 0:
            break
        light = todo[deg[todo] <= n_items // 6]
        heavy = todo[deg[todo] > n_items // 6]
        for u in heavy:  # exponential race: n smallest of E_i / p_i
            keys = rng.exponential(size=n_items) / p
            chosen[u] = np.argpartition(keys, deg[u] - 1)[: deg[u]] if deg[u] < n_items else np.arange(n_items)
        if len(light) == 0:
            todo = light
            break
        m = np.ceil(deg[light] * factor).astype(np.int64) + 16
        owner = np.repeat(np.arange(len(light)), m)
        draws = np.searchsorted(cdf, rng.random(owner.shape[0]), side="right").astype(np.int64)
        draws = np.minimum(draws, n_items - 1)
        key = owner * n_items + draws
        _, first = np.unique(key, return_index=True)
        first.sort()  # keep draw order
        o, it = owner[first], draws[first]
        starts = np.searchsorted(o, np.arange(len(light)), side="left")
        ends = np.searchsorted(o, np.arange(len(light)), side="right")
        redo = []
        for i, u in enumerate(light):
            if ends[i] - starts[i] >= deg[u]:
                chosen[u] = it[starts[i]: starts[i] + deg[u]]
            else:
                redo.append(u)
        todo = np.array(redo, dtype=np.int64)
        factor *= 2.5
    for u in todo:
        keys = rng.exponential(size=n_items) / p
        chosen[u] = np.argpartition(keys, deg[u] - 1)[: deg[u]]
    return chosen


def make_ratings(shape: str | tuple = "ml-100k", seed: int = SEED, n_users: int | None = None) -> Ratings:
    """Ratings of the named shape.  ``n_users`` (optional) keeps the per-user degree law and the
    item universe but generates only that many users (nnz scales proportionally)."""
    users, items, nnz, lo, hi, med, sigma, half = SHAPES[shape] if isinstance(shape, str) else shape
    if n_users is not None and n_users != users:
        nnz = int(round(nnz * n_users / users))
        users = n_users
    rng = np.random.default_rng(seed)
    hi = min(hi, items)
    deg = _degrees(rng, users, nnz, lo, hi, med, sigma)
    perm = rng.permutation(items) + 1                      # popularity rank -> item id (1-based)
    p = 1.0 / np.arange(1, items + 1, dtype=np.float64)    # Zipf(1.0)
    p /= p.sum()
    chosen = _draw_items(rng, deg, items, p)
    offsets = np.zeros(users + 1, dtype=np.int64)
    np.cumsum(deg, out=offsets[1:])
    it = np.empty(offsets[-1], dtype=np.int32)
    for u in range(users):
        ids = perm[chosen[u]]
        ids.sort()
        it[offsets[u]: offsets[u + 1]] = ids
    b_u = rng.normal(0.0, 0.5, size=users)
    b_i = rng.normal(0.0, 0.5, size=items + 1)
    raw = 3.5 + np.repeat(b_u, deg) + b_i[it] + rng.normal(0.0, 1.0, size=it.shape[0])
    r = np.clip(np.rint(raw * 2) / 2 if half else np.rint(raw), 1.0, 5.0).astype(np.float32)
    if half:
        r = np.maximum(r, 0.5).astype(np.float32)
    return Ratings(users, items, offsets, it, r)


def make_weights(n_items: int, density: float = 0.9, seed: int = SEED, dtype=np.float64) -> np.ndarray:
    """Kernel-bench item-similarity table: symmetric (N+1)^2, zero diagonal,
    row/col 0 unused, a fraction ``density`` of pairs carry a weight U(0.5,1] rounded to the 6
    significant digits that survive the out_fin_ text format (knn2.cpp:157-160)."""
    rng = np.random.default_rng(seed + 1)
    n1 = n_items + 1
    w = rng.random((n1, n1))
    keep = rng.random((n1, n1)) < density
    w = np.where(keep, np.round(1.0 - 0.5 * w, 6), 0.0)
    w = np.triu(w, 1)
    w = w + w.T
    w[0, :] = 0.0
    w[:, 0] = 0.0
    return np.ascontiguousarray(w, dtype=dtype)


def fold_split(r: Ratings, num_div: int = 5, seed: int = SEED):
    """fold_cross_validation.py:11-56 with a seeded shuffle: user-disjoint folds, a fold is cut
    every time its user count exceeds num_usr / num_div (:41).  Returns a list of arrays of user
    indices (0-based) -- fold i is the test (``.validate``) set of run i, the rest is train."""
    rng = np.random.default_rng(seed + 2)
    keys = rng.permutation(r.n_users)
    folds, cur = [], []
    for k in keys:
        cur.append(k)
        if len(cur) > r.n_users / num_div:
            folds.append(np.array(cur, dtype=np.int64))
            cur = []
    folds.append(np.array(cur, dtype=np.int64))
    return folds


def subset(r: Ratings, user_idx: np.ndarray):
    """(user_ids 1-based, CSR) restricted to the given 0-based users, in the given order."""
    user_idx = np.asarray(user_idx, dtype=np.int64)
    deg = r.degrees()[user_idx]
    offsets = np.zeros(len(user_idx) + 1, dtype=np.int64)
    np.cumsum(deg, out=offsets[1:])
    items = np.empty(offsets[-1], dtype=np.int32)
    rat = np.empty(offsets[-1], dtype=np.float32)
    for i, u in enumerate(user_idx):
        items[offsets[i]: offsets[i + 1]] = r.items[r.offsets[u]: r.offsets[u + 1]]
        rat[offsets[i]: offsets[i + 1]] = r.ratings[r.offsets[u]: r.offsets[u + 1]]
    return user_idx + 1, offsets, items, rat


def write_rating_file(path: str, user_ids, offsets, items, ratings) -> None:
    """GraphLab-style ``user<TAB>item<TAB>rating`` lines (fold_cross_validation.py:39)."""
    with open(path, "w") as f:
        for i, u in enumerate(user_ids):
            for j in range(offsets[i], offsets[i + 1]):
                f.write("%d\t%d\t%g\n" % (u, items[j], ratings[j]))
