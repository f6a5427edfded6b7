"""BASELINE.json configs[1]: the reference's own workflow (run_test_precompute.sh) through the drop-in CLI tools on a
MovieLens-shaped synthetic data set: u.data -> 5 user-disjoint folds -> per fold  knn; knn2; precompute_local_threads N;
local_calc_precomp --pct P  -> RMSE over out_res.  Prints one JSON line per fold with the wall seconds of every tool
(process start to exit: text parsing, GPU work and text / binary output included).

    python scripts/cli_pipeline.py ml-1m 0 20 [binary]      # shape, fold, pct
"""
import json, os, sys, tempfile, time
sys.path.insert(0, ".")
from collaborative_filtering_b200 import datasets as D
from collaborative_filtering_b200 import workflow as WF

shape = sys.argv[1] if len(sys.argv) > 1 else "ml-1m"
fold = int(sys.argv[2]) if len(sys.argv) > 2 else 0
pct = int(sys.argv[3]) if len(sys.argv) > 3 else 20
if len(sys.argv) > 4 and sys.argv[4] == "binary":
    os.environ["GSI_EIGEN_BINARY"] = "1"
r = D.make_ratings(shape)
tmp = tempfile.mkdtemp(prefix="gsi_cli_")
src = os.path.join(tmp, "u.data")
users, items, ratings = r.triples()
with open(src, "w") as f:
    for u, i, x in zip(users, items, ratings):
        f.write("%d\t%d\t%d\n" % (u, i, int(x)))
t0 = time.time()
nf = WF.fold_cross_validation(src, 5, os.path.join(tmp, "cross_validation"))
res = WF.run_pipeline(os.path.join(tmp, "cross_validation"), os.path.join(tmp, "work"), folds=[fold], pct=pct,
                      precompute_tool="precompute_local_threads", threads=os.cpu_count() or 8, log=open(os.devnull, "w"), stage_timeout=900)
out = dict(res[0], shape=shape, folds=nf, pct=pct, records="binary" if os.environ.get("GSI_EIGEN_BINARY") else "text",
           out_eigen_bytes=sum(os.path.getsize(os.path.join(tmp, "work", f)) for f in os.listdir(os.path.join(tmp, "work")) if f.startswith("out_eigen_")))
print(json.dumps(out))
