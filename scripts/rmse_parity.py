"""RMSE parity on one full fold (the "RMSE delta" third of BASELINE.json's metric, north_star: agree to 1e-4).

ML-100K shape (943 x 1682, 100k integer ratings, seed 31413), user-disjoint 5-fold split with the semantics of
fold_cross_validation.py; fold F is the validation role, the other four are train.  Stages:
  knn + knn2 on the GPU (train ratings)            -> item graph, weights (w > 0.01)
  precompute on the GPU (validation users)         -> records (sig_min, lambda, U)
  predictor on the GPU: every (movie, validation user) pair
  oracle predictor (oracle/gsi_oracle.py, numpy restatement of local_calc_precomp.cpp:217-380) on the SAME records
  (stage-wise protocol, SURVEY.md H1: the signed column drop makes end-to-end results depend on eigenvector signs).
Prints one JSON line: RMSE of both sides over the pairs both classify as well-posed, their difference, the largest
prediction difference, the status agreement, and the eigenvalue / sig_min agreement of the records with the oracle's own.

    python scripts/rmse_parity.py [fold] [n_users]
"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from collaborative_filtering_b200 import datasets as D  # noqa: E402
from collaborative_filtering_b200.api import Context  # noqa: E402
from oracle import gsi_oracle as O  # noqa: E402

fold = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n_users = int(sys.argv[2]) if len(sys.argv) > 2 else None
r = D.make_ratings("ml-100k", n_users=n_users)
folds = D.fold_split(r, 5)
val_idx = np.sort(folds[fold])
trn_idx = np.sort(np.concatenate([f for i, f in enumerate(folds) if i != fold]))
_, v_off, v_items, v_rat = D.subset(r, val_idx)
_, t_off, t_items, t_rat = D.subset(r, trn_idx)

ctx = Context(0)
t0 = time.time()
a, b, w = ctx.knn_build(t_off, t_items, t_rat, r.n_items + 1, install_weights=False)      # knn2 weights of the train folds
t_knn = time.time() - t0
wd = np.zeros((r.n_items + 1, r.n_items + 1))                                              # what precompute_local reads back from out_fin_
wd[a, b] = w
ctx.set_weights(wd)
t0 = time.time()
recs = ctx.precompute(v_off, v_items)
t_pre = time.time() - t0
t0 = time.time()
out = ctx.predict(recs, v_rat.astype(np.float64))
t_pred = time.time() - t0

# oracle side: same graph (edges with (float)w > 0.1), same records
graph = O.item_graph([(int(x), int(y), float(z)) for x, y, z in zip(a, b, w)])
dl, ds = 0.0, 0.0
se_g = se_o = 0.0
n_ok = n_pairs = status_mismatch = 0
dpred = 0.0
t0 = time.time()
for ui in range(len(val_idx)):
    its = v_items[v_off[ui]: v_off[ui + 1]]
    ref = O.precompute_user(0, its, wd)                                                # the oracle's own record, for the record parity
    if len(ref.lam) == recs.k[ui]:
        dl = max(dl, float(np.abs(ref.lam - recs.lam_of(ui)).max()))
    ds = max(ds, float(np.abs(ref.sigs_min - recs.sig_of(ui)).max()))
    ud = dict(items=its.astype(np.int64), row_of={int(m): j for j, m in enumerate(its)},
              sigs_min=recs.sig_of(ui), lam=recs.lam_of(ui), vec=recs.vec_of(ui))
    ur = {int(m): float(v_rat[v_off[ui] + j]) for j, m in enumerate(its)}
    for j, m in enumerate(its):
        m = int(m)
        err, kk, pred, status, c = O.predict_pair(ud, m, graph.get(m, set()), ur, ur[m])
        g = v_off[ui] + j
        n_pairs += 1
        ok_g, ok_o = out["status"][g] == 0, status == O.PRED_OK
        if ok_g != ok_o:
            status_mismatch += 1
        if ok_g and ok_o:
            n_ok += 1
            se_g += float(out["err"][g])
            se_o += float(err)
            dpred = max(dpred, abs(float(out["pred"][g]) - float(pred)))
t_oracle = time.time() - t0
ctx.close()
rg, ro = float(np.sqrt(se_g / n_ok)), float(np.sqrt(se_o / n_ok))
print(json.dumps({"shape": "ml-100k", "fold": fold, "validation_users": int(len(val_idx)), "pairs": n_pairs, "well_posed_pairs": n_ok,
                  "rmse_gpu": rg, "rmse_oracle": ro, "rmse_delta": abs(rg - ro), "max_abs_pred_diff": dpred,
                  "status_mismatches": status_mismatch, "max_abs_lambda_diff_vs_oracle_records": dl,
                  "max_abs_sig_min_diff_vs_oracle_records": ds,
                  "seconds": {"knn+knn2 (gpu, wall)": round(t_knn, 3), "precompute (gpu, wall)": round(t_pre, 3),
                              "predict (gpu, wall)": round(t_pred, 3), "oracle predictor (cpu)": round(t_oracle, 1)}}))
