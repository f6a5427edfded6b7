"""Generates the committed golden fixtures from the CPU oracle (run from the repo root:
``python tests/golden/make_golden.py``).  The reference has no golden vectors for this path and
cannot be built here, so these vectors pin *our oracle's* behaviour
(parity unpinned with respect to a genuine Eigen/GraphLab run) and let the GPU tests run without
recomputing it.  Deterministic: seed 31413 (make_synthetic_als_data.cpp:125)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import gsi_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def tiny_dataset(rng, n_train=70, n_val=14, n_items=60, half=False):
    """Dense-ish toy ratings so that many item pairs share > 5 raters (knn2.cpp:142)."""
    rows_t, rows_v = [], []
    b_i = rng.normal(0, 0.6, n_items + 1)
    for u in range(1, n_train + n_val + 1):
        n = int(rng.integers(8, 26))
        items = np.sort(rng.choice(np.arange(1, n_items + 1), size=n, replace=False))
        b_u = rng.normal(0, 0.5)
        raw = 3.5 + b_u + b_i[items] + rng.normal(0, 1.0, n)
        r = np.clip(np.rint(raw * 2) / 2 if half else np.rint(raw), 1, 5)
        dst = rows_t if u <= n_train else rows_v
        for m, x in zip(items, r):
            dst.append((u, int(m), float(x)))
    return rows_t, rows_v


def run_case(name, rng, **kw):
    out = os.path.join(HERE, name)
    os.makedirs(os.path.join(out, "movielens"), exist_ok=True)
    train, val = tiny_dataset(rng, **kw)
    with open(os.path.join(out, "movielens", "u0.train"), "w") as f:
        f.writelines("%d\t%d\t%g\n" % t for t in train)
    with open(os.path.join(out, "movielens", "u0.validate"), "w") as f:
        f.writelines("%d\t%d\t%g\n" % t for t in val)
    rat, test_rat, edg = O.knn1(train, val)
    fin = O.knn2(rat, edg)
    fin_text = O.format_fin(fin)
    open(os.path.join(out, "out_rat_1_of_1"), "w").write(O.format_rat(rat))
    open(os.path.join(out, "out_test_rat_1_of_1"), "w").write(O.format_rat(test_rat))
    open(os.path.join(out, "out_edg_1_of_1"), "w").write(O.format_edg(edg))
    open(os.path.join(out, "out_fin_1_of_1"), "w").write(fin_text)
    # downstream stages read the 6-digit text, exactly as the reference pipeline does
    fin_rt = O.parse_fin(fin_text)
    weights = O.weights_from_fin(fin_rt)
    users = O.users_from_validate(val)
    recs = O.precompute_all(users, weights)
    O.write_out_eigen(os.path.join(out, "out_eigen_"), recs)
    test_rt = O.parse_rat(O.format_rat(test_rat))
    graph = O.item_graph(fin_rt)
    res = {}
    for b1 in (True, False):
        ud = O.parse_out_eigen(os.path.join(out, "out_eigen_"), bug_b1=b1)
        rows = O.local_calc_precomp(ud, graph, test_rt)
        res[b1] = rows
        open(os.path.join(out, "out_res_b1_%s" % ("on" if b1 else "off")), "w").write(O.format_res(rows))
    mse, _ = O.knn3(fin_rt, test_rt)
    full = dict(
        weights=weights,
        user=np.array([r.user for r in recs], dtype=np.int64),
        offsets=np.concatenate([[0], np.cumsum([len(r.items) for r in recs])]).astype(np.int64),
        items=np.concatenate([r.items for r in recs]).astype(np.int32),
        sig_min=np.concatenate([r.sigs_min for r in recs]),
        k=np.array([len(r.lam) for r in recs], dtype=np.int32),
        lam=np.concatenate([r.lam for r in recs]),
        vec=np.concatenate([r.vec.reshape(-1) for r in recs]),
        fin_a=np.array([e[0] for e in fin], dtype=np.int32),
        fin_b=np.array([e[1] for e in fin], dtype=np.int32),
        fin_w=np.array([e[2] for e in fin], dtype=np.float64),
    )
    for b1 in (True, False):
        tag = "on" if b1 else "off"
        rows = res[b1]
        full["res_%s_movie" % tag] = np.array([r[0] for r in rows], dtype=np.int64)
        full["res_%s_user" % tag] = np.array([r[1] for r in rows], dtype=np.int64)
        full["res_%s_err" % tag] = np.array([r[2] for r in rows], dtype=np.float32)
        full["res_%s_kk" % tag] = np.array([r[3] for r in rows], dtype=np.int32)
        full["res_%s_pred" % tag] = np.array([r[4] for r in rows], dtype=np.float64)
        full["res_%s_status" % tag] = np.array([r[5] for r in rows], dtype=np.int32)
        full["res_%s_c" % tag] = np.array([r[6] for r in rows], dtype=np.int32)
    np.savez_compressed(os.path.join(out, "oracle.npz"), **full)
    meta = dict(knn3_avg_mse=mse, n_users=len(recs), n_edges=len(fin),
                rmse_b1_on=O.rmse_of(res[True]), rmse_b1_off=O.rmse_of(res[False]))
    json.dump(meta, open(os.path.join(out, "meta.json"), "w"), indent=1)
    print(name, meta)


def run_precompute_case(name, rng):
    """Precompute-only vectors on a random directed-asymmetric table: exercises k > 2, n = 1, 2, 3,
    ids beyond the table (precompute_local.cpp:187-188), isolated items (degree 0 -> 1, :204-205)
    and the bucket edges of the GPU solver."""
    out = os.path.join(HERE, name)
    os.makedirs(out, exist_ok=True)
    n_items = 220
    w = np.round(1.0 - 0.5 * rng.random((n_items + 1, n_items + 1)), 6)
    w = np.where(rng.random(w.shape) < 0.45, w, 0.0)
    w = np.triu(w, 1)
    w = w + w.T
    asym = rng.random(w.shape) < 0.05                      # B9: last-digit asymmetry
    w = np.where(asym & (w > 0), np.round(w + 1e-6, 6), w)
    w[0, :] = 0
    w[:, 0] = 0
    w[17, :] = 0                                           # isolated items
    w[:, 17] = 0
    w[101, :] = 0
    w[:, 101] = 0
    sizes = [1, 2, 3, 5, 8, 16, 31, 32, 33, 47, 63, 64, 65, 96, 127, 128, 129, 160, 161, 200]
    users = {}
    for i, n in enumerate(sizes):
        ids = np.sort(rng.choice(np.arange(1, n_items + 30), size=n, replace=False))  # some ids > N
        if n >= 5:
            ids[1] = 17 if 17 not in ids else ids[1]
            ids = np.unique(ids)
        users[O.UIMAX - (i + 1)] = ids.astype(np.int64)
    users[O.UIMAX - 100] = np.array([17, 101, 240], dtype=np.int64)   # all isolated / out of table
    recs = O.precompute_all(users, w)
    np.savez_compressed(
        os.path.join(out, "oracle.npz"), weights=w,
        user=np.array([r.user for r in recs], dtype=np.int64),
        offsets=np.concatenate([[0], np.cumsum([len(r.items) for r in recs])]).astype(np.int64),
        items=np.concatenate([r.items for r in recs]).astype(np.int32),
        sig_min=np.concatenate([r.sigs_min for r in recs]),
        k=np.array([len(r.lam) for r in recs], dtype=np.int32),
        lam=np.concatenate([r.lam for r in recs]),
        vec=np.concatenate([r.vec.reshape(-1) for r in recs]))
    print(name, "k =", [len(r.lam) for r in recs])


def run_cheby_case(name, seed=31413 + 7, nv=48, density=0.25, ncoef=9):
    """cheby.cpp inputs (coeff, graph_topology, graph_signal) and the oracle's graph_filtered_signal: sub-threshold weights,
    a duplicate line, an isolated vertex, a vertex that only appears in the topology."""
    rng = np.random.default_rng(seed)
    d = os.path.join(HERE, name)
    os.makedirs(d, exist_ok=True)
    ids = np.sort(rng.choice(np.arange(1, 4 * nv), size=nv, replace=False))
    lines = []
    for i in range(nv):
        for j in range(i + 1, nv):
            if rng.random() < density:
                lines.append((int(ids[i]), int(ids[j]), float(np.round(rng.uniform(0.02, 1.0), 6))))
    lines.append(lines[3])
    lines.append((int(ids[0]), 4 * nv + 50, 0.75))                             # 4 nv + 50 has no graph_signal line
    topo = "".join("%d %d %s\n" % (a, b, O.fmt_g(w)) for a, b, w in lines)
    signal = {int(v): float(np.round(rng.normal(3.5, 1.0), 4)) for v in ids}
    signal[4 * nv + 7] = 2.5                                                   # isolated
    sig_text = "".join("%d %s\n" % (v, O.fmt_g(x)) for v, x in signal.items())
    coeff_text = " ".join(O.fmt_g(float(np.round(c, 5))) for c in rng.normal(0, 1, ncoef)) + "\n"
    open(os.path.join(d, "graph_topology"), "w").write(topo)
    open(os.path.join(d, "graph_signal"), "w").write(sig_text)
    open(os.path.join(d, "coeff"), "w").write(coeff_text)
    out = O.cheby_filter(O.cheby_parse_topology(topo), O.cheby_parse_signal(sig_text), O.cheby_parse_coeff(coeff_text))
    open(os.path.join(d, "graph_filtered_signal_1_of_1"), "w").write(O.cheby_format(out))
    np.savez(os.path.join(d, "oracle.npz"), vertex=np.array(sorted(out), dtype=np.int64), value=np.array([out[v] for v in sorted(out)]))
    print(name, len(out), "vertices")


def run_local_calc_case(name):
    """Per-movie variant (local_calc.cpp) on the out_fin_ / out_test_rat_ texts of an existing case."""
    d = os.path.join(HERE, name)
    fin_rt = O.parse_fin(open(os.path.join(d, "out_fin_1_of_1")).read())
    test_rt = O.parse_rat(open(os.path.join(d, "out_test_rat_1_of_1")).read())
    rows = O.local_calc(fin_rt, test_rt)
    open(os.path.join(d, "out_res_local_calc"), "w").write(O.format_res(rows))
    np.savez_compressed(
        os.path.join(d, "local_calc.npz"),
        movie=np.array([r[0] for r in rows], dtype=np.int64), user=np.array([r[1] for r in rows], dtype=np.int64),
        err=np.array([r[2] for r in rows], dtype=np.float32), kk=np.array([r[3] for r in rows], dtype=np.int32),
        pred=np.array([r[4] for r in rows], dtype=np.float64), status=np.array([r[5] for r in rows], dtype=np.int32),
        lim=np.array([r[6] for r in rows], dtype=np.int32), w_lim=np.array([r[7] for r in rows], dtype=np.float64),
        gap=np.array([r[8] for r in rows], dtype=np.float64))
    print(name, "local_calc:", len(rows), "rows, rmse", O.rmse_of(rows))


if __name__ == "__main__":
    run_cheby_case("cheby_small")
    rng = np.random.default_rng(31413)
    run_case("tiny_int", rng, half=False)
    run_case("tiny_half", rng, half=True, n_train=90, n_val=10, n_items=50)
    run_precompute_case("precompute_rand", rng)
    run_local_calc_case("tiny_int")
    run_local_calc_case("tiny_half")
