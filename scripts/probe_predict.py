"""Times the predictor (gsi_predict_host) on an ML-100K-shaped data set with the item graph built by the
knn2 stage itself (cosine weights over common raters): precompute, then one prediction per
(user, rated movie) pair of the users with n <= NMAX."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from collaborative_filtering_b200 import datasets as D
from collaborative_filtering_b200.api import Context

nmax = int(sys.argv[1]) if len(sys.argv) > 1 else 256
shape = sys.argv[2] if len(sys.argv) > 2 else "ml-100k"
r = D.make_ratings(shape)
ctx = Context(0)
t0 = time.time()
a, b, w = ctx.knn_build(r.offsets, r.items, r.ratings, r.n_items + 1, install_weights=True)
print("knn2: %d edges in %.3fs (density %.3f)" % (len(w), time.time() - t0, len(w) / float(r.n_items) ** 2))
deg = np.diff(r.offsets)
sel = np.nonzero(deg <= nmax)[0]
_, s_off, s_items, s_rat = D.subset(r, sel)
t0 = time.time(); recs = ctx.precompute(s_off, s_items); t1 = time.time()
print("precompute %d users in %.3fs, mean k/n %.2f" % (len(sel), t1 - t0, float(np.mean(recs.k / np.diff(s_off)))))
ctx.timing_enable(True); ctx.timing_reset()
t0 = time.time(); out = ctx.predict(recs, s_rat.astype(np.float64)); t1 = time.time()
tm = ctx.timing()["predict"]
npairs = int(s_off[-1])
st = np.bincount(out["status"], minlength=5)
print("predict %d pairs: wall %.3fs, kernel %.1f ms (%d launches) -> %.0f pairs/s kernel-only; status ok/empty/under/sing/skip = %s"
      % (npairs, t1 - t0, tm["ms"], tm["launches"], npairs / (tm["ms"] * 1e-3), st.tolist()))
ok = out["status"] == 0
print("rmse over ok pairs: %.4f (n=%d); mean cols %.1f mean kk %.1f" % (np.sqrt(np.mean(out["err"][ok])), ok.sum(), out["cols"][ok].mean(), out["kk"][ok].mean()))
ctx.close()
