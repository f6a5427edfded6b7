"""First-light probe: correctness spot checks + raw timings per size bucket (not the bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from collaborative_filtering_b200 import datasets as D
from collaborative_filtering_b200.api import Context

ctx = Context(0)
print(ctx.version, "fp64 fma TF/s", ctx.measure_fp64_tflops(False), "dmma TF/s", ctx.measure_fp64_tflops(True), flush=True)
shape = sys.argv[1] if len(sys.argv) > 1 else "ml-100k"
r = D.make_ratings(shape)
w = D.make_weights(r.n_items)
ctx.set_weights(w)
ctx.timing_enable(True)
for rep in range(2):
    ctx.timing_reset()
    t = time.time()
    recs = ctx.precompute(r.offsets, r.items)
    dt = time.time() - t
    print("rep", rep, shape, "users", r.n_users, "host-api seconds %.3f -> %.0f users/s" % (dt, r.n_users / dt), flush=True)
    for k, v in ctx.timing().items():
        if v["launches"]:
            print("   %-10s ms=%.2f launches=%d samples=%d" % (k, v["ms"], v["launches"], v["samples"]))
# n sweep (uniform sizes)
rng = np.random.default_rng(0)
for n in [16, 32, 64, 96, 128, 160, 200, 400, 800, 1600]:
    if n > r.n_items: break
    nu = max(2, min(4000, int(2e9 / n**3)))
    lists = [np.sort(rng.choice(np.arange(1, r.n_items + 1), n, replace=False)).astype(np.int32) for _ in range(nu)]
    offs = np.arange(nu + 1, dtype=np.int64) * n
    items = np.concatenate(lists)
    ctx.timing_reset()
    t = time.time(); recs = ctx.precompute(offs, items); dt = time.time() - t
    tm = ctx.timing()
    dev_ms = sum(v["ms"] * (v["launches"] / max(1, v["samples"])) for v in tm.values())
    print("n=%d users=%d wall %.3fs dev-est %.1f ms -> %.0f users/s (dev) ; 9n^3 GF/s %.1f" % (n, nu, dt, dev_ms, nu / (dev_ms / 1e3), 9 * n**3 * nu / (dev_ms / 1e3) / 1e9), flush=True)
