"""Heavy-chunk probe for profiling the block-Jacobi kernels: a few users of one size."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from collaborative_filtering_b200 import datasets as D
from collaborative_filtering_b200.api import Context

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
nu = int(sys.argv[2]) if len(sys.argv) > 2 else 6
n_items = 10681
w = D.make_weights(n_items)
ctx = Context(0)
ctx.set_workspace_limit(32 << 30)
ctx.set_weights(w)
rng = np.random.default_rng(0)
lists = [np.sort(rng.choice(np.arange(1, n_items + 1), n, replace=False)).astype(np.int32) for _ in range(nu)]
offs = np.arange(nu + 1, dtype=np.int64) * n
items = np.concatenate(lists)
ctx.timing_enable(True)
for rep in range(2):
    ctx.timing_reset()
    t = time.time(); recs = ctx.precompute(offs, items); dt = time.time() - t
    tm = ctx.timing()
    print("n=%d users=%d wall %.3fs" % (n, nu, dt), {k: (round(v["ms"] * v["launches"] / max(1, v["samples"]), 1), v["launches"]) for k, v in tm.items() if v["launches"]}, flush=True)
flop = 9.0 * n ** 3 * nu
bj = sum(tm[k]["ms"] * tm[k]["launches"] / max(1, tm[k]["samples"]) for k in ("bj_gram", "bj_inner", "bj_update"))
print("algorithmic TF/s over bj kernels: %.2f" % (flop / (bj * 1e-3) / 1e12))
