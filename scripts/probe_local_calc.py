"""Times the per-movie variant (gsi_local_calc_host, local_calc.cpp) on one ML-100K-shaped fold and checks a
bounded sample of it against the oracle.

ML-100K shape (943 x 1682, 100k integer ratings, seed 31413), 5-fold user split; the item graph comes from the knn2
stage on the train folds, the test ratings are the validation fold.  A seeded PCT % of the movie vertices is
computed (the tool's own --pct switch, local_calc.cpp:266).  Prints one JSON line.

    python scripts/probe_local_calc.py [pct] [n_oracle_movies] [shape]
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/probe_local_calc.py [pct] 0 [shape]

Under torchrun (WORLD_SIZE > 1) the movie vertices are dealt over the ranks (shard.shard_movies through
Context.local_calc_shard), every rank computes the pairs of its own movies, pair counts and squared errors are
all-reduced and the wall time is the slowest rank's; the oracle comparison runs on rank 0 over its own pairs only.
(The multi-GPU mode has been exercised on CPU for its dealing logic only -- tests/test_shard.py.)
"""
import os
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from collaborative_filtering_b200 import datasets as D  # noqa: E402
from collaborative_filtering_b200.api import Context  # noqa: E402
from oracle import gsi_oracle as O  # noqa: E402

pct = float(sys.argv[1]) if len(sys.argv) > 1 else 5.0
n_oracle = int(sys.argv[2]) if len(sys.argv) > 2 else 3
shape = sys.argv[3] if len(sys.argv) > 3 else "ml-100k"
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl")
r = D.make_ratings(shape)
folds = D.fold_split(r, 5)
val_idx = np.sort(folds[0])
trn_idx = np.sort(np.concatenate(folds[1:]))
_, v_off, v_items, v_rat = D.subset(r, val_idx)
_, t_off, t_items, t_rat = D.subset(r, trn_idx)
ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
a, b, w = ctx.knn_build(t_off, t_items, t_rat, r.n_items + 1, install_weights=False)
ctx.set_weights_edges(a, b, w.astype(np.float64))
kept = w.astype(np.float32).astype(np.float64) > 0.1
nb = np.bincount(a[kept], minlength=r.n_items + 1)                      # out-degree in the thresholded item graph
rng = np.random.default_rng(31413)
movies = np.unique(v_items)
chosen = movies[rng.random(len(movies)) * 100.0 < pct]
mask = np.isin(v_items, chosen).astype(np.uint8)
ctx.local_calc(v_off, v_items, v_rat.astype(np.float64), pair_mask=np.zeros_like(mask))   # warm-up: neighbour lists only
if world > 1:
    dist.barrier()
t0 = time.time()
if world > 1:
    out = ctx.local_calc_shard(v_off, v_items, v_rat.astype(np.float64), nb + 1, rank, world, pair_mask=mask)
else:
    out = ctx.local_calc(v_off, v_items, v_rat.astype(np.float64), pair_mask=mask)
wall = time.time() - t0
done = out["status"] != 4
if world > 1:                      # merge: slowest rank's wall time, pair counts and squared errors summed
    okr = out["status"] == 0
    st = torch.tensor([float(done.sum()), float(okr.sum()), float(out["err"][okr].astype(np.float64).sum())], dtype=torch.float64, device="cuda")
    tw = torch.tensor([wall], dtype=torch.float64, device="cuda")
    dist.all_reduce(st)
    dist.all_reduce(tw, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps(dict(shape=shape, pct=pct, n_gpus=world, pairs=int(st[0].item()), wall_s=round(tw.item(), 3),
                              pairs_per_s=round(st[0].item() / tw.item(), 1), rmse_ok=float(np.sqrt(st[2].item() / max(st[1].item(), 1.0))))))
    dist.destroy_process_group()
    ctx.close()
    sys.exit(0)
ok = out["status"] == 0
res = dict(shape=shape + " fold 0", pct=pct, movies=int(len(chosen)), pairs=int(done.sum()), wall_s=round(wall, 3),
           pairs_per_s=round(float(done.sum()) / wall, 1),
           local_graph_nodes=dict(mean=float(nb[chosen].mean() + 1), max=int(nb[chosen].max() + 1)),
           status=np.bincount(out["status"], minlength=5).tolist(),
           rmse_ok=float(np.sqrt(np.mean(out["err"][ok]))) if ok.any() else None,
           mean_lim=float(out["cols"][ok].mean()) if ok.any() else None, mean_kk=float(out["kk"][ok].mean()) if ok.any() else None)
# oracle on the smallest chosen local graphs (two dense eigensolves per pair on the CPU)
fin = [(int(x), int(y), float(z)) for x, y, z in zip(a, b, w)]
gw = O.item_graph_weights(fin)
users = O.UIMAX - (val_idx + 1)
test_rat = {}
for ui in range(len(val_idx)):
    for t in range(v_off[ui], v_off[ui + 1]):
        test_rat.setdefault(int(v_items[t]), {})[int(users[ui])] = float(v_rat[t])
pos = {(int(v_items[t]), int(users[ui])): t for ui in range(len(val_idx)) for t in range(v_off[ui], v_off[ui + 1])}
small = sorted((int(m) for m in chosen if nb[m] + 1 >= 3), key=lambda m: nb[m])[:n_oracle]
t0 = time.time()
rows = [row for m in small for row in O.local_calc_movie(m, gw, test_rat)]
t_or = time.time() - t0
dw = dp = 0.0
n_cmp = mism = 0
for (m, u, err, kk, pred, status, lim, w_lim, gap) in rows:
    t = pos[(m, u)]
    if kk == 0:
        continue
    dw = max(dw, abs(out["w_lim"][t] - w_lim))
    if out["cols"][t] != lim or (out["status"][t] == 0) != (status == O.PRED_OK):
        mism += 1
    elif status == O.PRED_OK and gap > 1e-6:
        dp = max(dp, abs(out["pred"][t] - pred))
        n_cmp += 1
res["oracle"] = dict(movies=small, nodes=[int(nb[m] + 1) for m in small], pairs=len(rows), cpu_s=round(t_or, 2),
                     cpu_pairs_per_s=round(len(rows) / max(t_or, 1e-9), 2), max_dw_lim=dw, max_dpred=dp, compared=n_cmp,
                     lim_or_status_mismatch=mism)
# CPU cost of the reference's algorithm at a typical size: the chosen movie closest to the mean local graph, a few users
mid = min((int(m) for m in chosen if nb[m] + 1 >= 3), key=lambda m: abs(nb[m] + 1 - res["local_graph_nodes"]["mean"]))
few = {mid: dict(sorted(test_rat[mid].items())[:4])}
few.update({m: d for m, d in test_rat.items() if m != mid})
t0 = time.time()
rows_mid = O.local_calc_movie(mid, gw, few)
t_mid = time.time() - t0
res["oracle_typical"] = dict(movie=mid, nodes=int(nb[mid] + 1), pairs=len(rows_mid), cpu_s=round(t_mid, 2),
                             cpu_pairs_per_s=round(len(rows_mid) / max(t_mid, 1e-9), 2), threads="numpy/LAPACK default",
                             max_dw_lim=max(abs(out["w_lim"][pos[(x[0], x[1])]] - x[7]) for x in rows_mid if x[3] > 0),
                             max_dpred=max(abs(out["pred"][pos[(x[0], x[1])]] - x[4]) for x in rows_mid if x[3] > 0 and x[5] == 0 and x[8] > 1e-6))
ctx.close()
print(json.dumps(res))
