"""Host-side multi-GPU logic on CPU: LPT user sharding, and a world_size-2 gloo run that checks
the ranks' shards partition the users and that W replication by broadcast is bit-identical."""
import os
import subprocess
import sys

import numpy as np

from collaborative_filtering_b200 import datasets as D
from collaborative_filtering_b200 import shard as SH

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lpt_partition_and_balance():
    r = D.make_ratings("ml-1m")
    deg = r.degrees()
    for world in (1, 2, 4, 8):
        owner = SH.lpt_assign(deg, world)
        assert owner.min() == 0 and owner.max() == world - 1
        parts = [SH.shard_users(deg, k, world) for k in range(world)]
        allu = np.sort(np.concatenate(parts))
        assert np.array_equal(allu, np.arange(r.n_users))           # every user exactly once
        # the heavy tail is spread: imbalance bounded by the single heaviest user
        heaviest = SH.user_cost(deg).max() / (SH.user_cost(deg).sum() / world)
        assert SH.imbalance(deg, owner, world) <= max(1.02, heaviest + 1e-9)


_WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, %r)
from collaborative_filtering_b200 import datasets as D, shard as SH
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% os.environ["PORT"],
                        rank=int(os.environ["RANK"]), world_size=2)
rank = dist.get_rank()
r = D.make_ratings("ml-100k")
mine = SH.shard_users(r.degrees(), rank, 2)
# W replicated by broadcast from rank 0
w = torch.from_numpy(D.make_weights(64)) if rank == 0 else torch.zeros((65, 65), dtype=torch.float64)
dist.broadcast(w, src=0)
flag = torch.zeros(r.n_users, dtype=torch.int32)
flag[torch.from_numpy(mine)] = 1
dist.all_reduce(flag)
cost = torch.tensor([float(SH.user_cost(r.degrees()[mine]).sum())], dtype=torch.float64)
costs = [torch.zeros(1, dtype=torch.float64) for _ in range(2)]
dist.all_gather(costs, cost)
if rank == 0:
    assert int(flag.min()) == 1 and int(flag.max()) == 1, "shards must partition the users"
    assert np.array_equal(w.numpy(), D.make_weights(64))
    c = [float(x) for x in costs]
    assert max(c) / (sum(c) / 2) < 1.05, c
    print("OK", c)
else:
    assert np.array_equal(w.numpy(), D.make_weights(64))
dist.destroy_process_group()
""" % ROOT


def test_two_rank_gloo_sharding(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "OK" in outs[0]


_WORKER2 = r"""
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
import bench
from collaborative_filtering_b200 import datasets as D
rank = int(os.environ["RANK"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% os.environ["PORT"], rank=rank, world_size=2)
r = D.make_ratings("ml-100k", n_users=120)
deg = np.diff(r.offsets)
mine = bench.deal_users(deg, rank, 2)
other = bench.deal_users(deg, 1 - rank, 2)
assert len(np.intersect1d(mine, other)) == 0 and np.array_equal(np.sort(np.concatenate([mine, other])), np.arange(len(deg)))
assert abs(int(deg[mine].sum()) - int(deg[other].sum())) <= int(deg.max())          # dealt by descending n: balanced
pairs = int(deg[mine].sum())
sums, maxes = bench.reduce_predict_stats([pairs, 2.0 * pairs, 0.5 * pairs, pairs - rank], [10.0 + rank, 0.1 * (rank + 1)], torch.device("cpu"))
tot = int(deg.sum())
assert sums[0] == tot and sums[1] == 2.0 * tot and abs(sums[2] - 0.5 * tot) < 1e-9 and sums[3] == tot - 1
assert maxes[0] == 11.0 and abs(maxes[1] - 0.2) < 1e-12                                          # the slowest rank's times
out = sums
print("OK", out[0])
dist.destroy_process_group()
""" % ROOT


def test_two_rank_gloo_predict_sample_reduction(tmp_path):
    """bench.py's predictions/s sample at N > 1: the users are dealt over the ranks and the counts / squared errors are summed,
    the times maxed (world_size-2 gloo on CPU)."""
    script = tmp_path / "worker2.py"
    script.write_text(_WORKER2)
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "OK" in outs[0] and "OK" in outs[1]


def test_movie_sharding_for_local_calc():
    """Per-movie variant: movies are dealt by cost, the ranks' pair masks partition the requested pairs."""
    r = D.make_ratings("ml-100k")
    rng = np.random.default_rng(7)
    n_nodes = rng.integers(1, 1200, size=r.n_items + 1)
    n_pairs = np.bincount(r.items, minlength=r.n_items + 1)
    base = (rng.random(len(r.items)) < 0.5).astype(np.uint8)
    for world in (1, 2, 8):
        owner = SH.shard_movies(n_nodes, n_pairs, world)
        assert owner.min() >= 0 and owner.max() <= world - 1
        masks = [SH.local_calc_pair_mask(r.items, owner, k, base) for k in range(world)]
        assert np.array_equal(np.sum(masks, axis=0).astype(np.uint8), base)      # each kept pair on exactly one rank
        loads = np.bincount(owner, weights=SH.movie_cost(n_nodes, n_pairs), minlength=world)
        heaviest = SH.movie_cost(n_nodes, n_pairs).max() / loads.mean()
        assert loads.max() / loads.mean() <= max(1.05, heaviest + 1e-9)
        for k in range(world):                                                  # all pairs of a movie on its owner
            assert set(np.unique(r.items[masks[k] == 1])) <= set(np.nonzero(owner == k)[0])


_WORKER_LC = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, %r)
from collaborative_filtering_b200 import datasets as D, shard as SH
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% os.environ["PORT"],
                        rank=int(os.environ["RANK"]), world_size=2)
rank = dist.get_rank()
r = D.make_ratings("ml-100k")
n_pairs = np.bincount(r.items, minlength=r.n_items + 1)
n_nodes = (np.arange(r.n_items + 1) * 7919) %% 900 + 3          # same on both ranks (a stand-in for the graph's out-degrees)
owner = SH.shard_movies(n_nodes, n_pairs, 2)
mask = SH.local_calc_pair_mask(r.items, owner, rank)
# what the host merge does: every pair computed once, squared errors and counts all-reduced
done = torch.from_numpy(mask.astype(np.int32))
dist.all_reduce(done)
stats = torch.tensor([float(mask.sum()), float((r.ratings[mask == 1].astype(np.float64) ** 2).sum())], dtype=torch.float64)
dist.all_reduce(stats)
if rank == 0:
    assert int(done.min()) == 1 and int(done.max()) == 1, "pair masks must partition the pairs"
    assert stats[0].item() == len(r.items)
    assert abs(stats[1].item() - float((r.ratings.astype(np.float64) ** 2).sum())) < 1e-6
    print("OK", stats.tolist())
dist.destroy_process_group()
""" % ROOT


def test_gloo_world2_movie_deal_and_reduce(tmp_path):
    script = tmp_path / "worker_lc.py"
    script.write_text(_WORKER_LC)
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = []
    for rk in range(2):
        env = dict(os.environ, RANK=str(rk), PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=300)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "OK" in outs[0]


def test_local_calc_shard_masks_through_the_api(monkeypatch):
    """Context.local_calc_shard without a device: the call it forwards to is replaced by a recorder; the masks of the
    ranks partition the requested pairs, movie ids beyond the table are left out, a movie's pairs stay together."""
    from collaborative_filtering_b200.api import Context
    r = D.make_ratings("ml-100k", n_users=60)
    items = r.items.copy()
    items[::97] = r.n_items + 50                          # ids outside the weight table
    n_nodes = (np.arange(r.n_items + 1) * 31) % 400 + 1
    base = (np.arange(len(items)) % 3 != 0).astype(np.uint8)
    seen = []
    ctx = Context.__new__(Context)                         # no gsi_create: only the host-side dealing is exercised
    monkeypatch.setattr(Context, "local_calc", lambda self, off, it, rat, pair_mask=None: seen.append(np.array(pair_mask)) or {})
    monkeypatch.setattr(Context, "__del__", lambda self: None, raising=False)
    for rank in range(4):
        ctx.local_calc_shard(r.offsets, items, r.ratings.astype(np.float64), n_nodes, rank, 4, pair_mask=base)
    total = np.sum(seen, axis=0)
    inside = items < len(n_nodes)
    assert np.array_equal(total[inside], base[inside]) and (total[~inside] == 0).all()
    owner_of = {}
    for rank, m in enumerate(seen):
        for movie in np.unique(items[m == 1]):
            assert owner_of.setdefault(int(movie), rank) == rank
