"""CPU baseline of the per-movie variant: the C++ restatement of local_calc.cpp's vertex program (oracle/cpu_ref.cpp, the
reference's own solver class and its dense inverse / products) on a bounded sample of the ML-100K-shaped fold that
scripts/probe_local_calc.py runs on the GPU, with one worker per host thread (the GraphLab engine runs vertex programs
concurrently).  The item graph is the knn2 formula evaluated densely in numpy (float32 sums are exact for integer ratings).
Prints one JSON line.  CPU only.

    python scripts/cpu_baseline_local_calc.py [n_movies] [threads]
"""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from collaborative_filtering_b200 import datasets as D  # noqa: E402
from oracle import cpu_ref  # noqa: E402

n_movies = int(sys.argv[1]) if len(sys.argv) > 1 else 16
threads = int(sys.argv[2]) if len(sys.argv) > 2 else cpu_ref.hardware_threads()
r = D.make_ratings("ml-100k")
folds = D.fold_split(r, 5)
val_idx = np.sort(folds[0])
trn_idx = np.sort(np.concatenate(folds[1:]))
_, v_off, v_items, v_rat = D.subset(r, val_idx)
_, t_off, t_items, t_rat = D.subset(r, trn_idx)
rows = r.n_items + 1
R = np.zeros((len(trn_idx), rows), dtype=np.float32)
for u in range(len(trn_idx)):
    R[u, t_items[t_off[u]:t_off[u + 1]]] = t_rat[t_off[u]:t_off[u + 1]]
B = (R > 0).astype(np.float32)
num, den1, cnt = R.T @ R, (R * R).T @ B, B.T @ B
with np.errstate(all="ignore"):
    W = np.where(cnt > 5, num / (np.sqrt(den1) * np.sqrt(den1.T)), 0).astype(np.float32)      # knn2.cpp:127-146
np.fill_diagonal(W, 0)
W = np.where(W > 0.01, W, 0).astype(np.float64)                                                # knn2.cpp:157-160
deg = ((W.astype(np.float32).astype(np.float64)) > 0.1).sum(axis=1)
movies = np.unique(v_items)
movies = movies[deg[movies] + 1 >= 3]
order = movies[np.argsort(np.abs(deg[movies] + 1 - np.mean(deg[movies] + 1)))]               # closest to the mean local graph first
chosen = order[:n_movies]
mask = np.isin(v_items, chosen).astype(np.uint8)
out = cpu_ref.local_calc(W, v_off, v_items, v_rat.astype(np.float64), pair_mask=mask, n_threads=threads, honest=True)
ok = out["status"] == 0
print(json.dumps(dict(shape="ml-100k fold 0", movies=int(len(chosen)), nodes=[int(deg[m] + 1) for m in chosen[:8]],
                      mean_nodes_all=float(np.mean(deg[movies] + 1)), pairs=out["pairs"], threads=threads,
                      seconds=round(out["seconds"], 2), pairs_per_s=round(out["pairs"] / out["seconds"], 2),
                      rmse_ok=float(np.sqrt(np.mean(out["err"][ok]))) if ok.any() else None, kind="port (C++ restatement, honest)")))
