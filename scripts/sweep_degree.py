"""BASELINE.json configs[4]: per-user degree sweep n = 8 .. 4096: batch = max(1, ceil(2^26 / n^3)) * #SM
users per point (capped), synthetic W density 0.9, seed 31413 + n.  Device-resident users/s of the precompute path and the
algorithmic 9 n^3 TFLOP/s per bucket; run once per large-path variant:

    python scripts/sweep_degree.py                 # n <= 160: CTA Jacobi, n > 160: Householder + D&C
    GSI_LARGE=bj python scripts/sweep_degree.py    # n > 160: block Jacobi (previous path)
"""
import json, math, os, sys
import numpy as np
import torch
sys.path.insert(0, ".")
from collaborative_filtering_b200.api import Context, upper_bounds

N_ITEMS = 10681
points = [8, 16, 32, 64, 128, 160, 161, 256, 512, 1024, 2048, 4096]
if len(sys.argv) > 1:
    points = [int(x) for x in sys.argv[1:]]
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(31414)
n1 = N_ITEMS + 1
u = torch.rand((n1, n1), generator=g, device=dev, dtype=torch.float64)
keep = torch.rand((n1, n1), generator=g, device=dev) < 0.9
wv = torch.where(keep, torch.round((1.0 - 0.5 * u) * 1e6) / 1e6, torch.zeros_like(u)).triu(1)
d_w = wv + wv.T
d_w[0, :] = 0; d_w[:, 0] = 0
del u, keep, wv
stream = torch.cuda.current_stream()
ctx = Context(0, stream=stream.cuda_stream)
ctx.set_workspace_limit(64 << 30)
ctx.set_weights(d_w)
variant = os.environ.get("GSI_LARGE", "hh")
for n in points:
    users = min(max(1, math.ceil(2 ** 26 / n ** 3)) * 148, 40000)
    if os.environ.get("GSI_SWEEP_USERS"):          # probe: a fixed number of users per point (e.g. fewer than the SM count)
        users = int(os.environ["GSI_SWEEP_USERS"])
    if variant == "bj" and n > 160:
        users = min(users, 148 if n <= 1024 else 8)
    rng = np.random.default_rng(31413 + n)
    items = np.empty((users, n), dtype=np.int32)
    for i in range(users):
        items[i] = np.sort(rng.choice(N_ITEMS, size=n, replace=False) + 1)
    offsets = np.arange(users + 1, dtype=np.int64) * n
    lam_cap, vec_cap = upper_bounds(offsets)
    d_items = torch.from_numpy(items.reshape(-1)).to(dev)
    d_sig = torch.empty(users * n, dtype=torch.float64, device=dev)
    d_k = torch.empty(users, dtype=torch.int32, device=dev)
    d_lo = torch.empty(users, dtype=torch.int64, device=dev)
    d_vo = torch.empty(users, dtype=torch.int64, device=dev)
    d_lam = torch.empty(lam_cap, dtype=torch.float64, device=dev)
    d_vec = torch.empty(vec_cap, dtype=torch.float64, device=dev)
    ctx.precompute_device(offsets, d_items, d_sig, d_k, d_lo, d_vo, d_lam, d_vec)      # warm-up
    reps = 3 if n <= 1024 else 2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        ctx.precompute_device(offsets, d_items, d_sig, d_k, d_lo, d_vo, d_lam, d_vec)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    path = "eig_cta (Jacobi)" if n <= ctx.small_max else ("block Jacobi" if variant == "bj" else "Householder + D&C")
    print(json.dumps({"n": n, "users": users, "path": path, "ms": ms, "users_per_s": users / (ms * 1e-3),
                      "alg_tflops_9n3": 9.0 * n ** 3 * users / (ms * 1e-3) / 1e12,
                      "mean_k_over_n": float(d_k.double().mean().item()) / n}), flush=True)
    del d_items, d_sig, d_k, d_lo, d_vo, d_lam, d_vec
ctx.close()
