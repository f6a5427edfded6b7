"""GPU parity tests of the knn chain against the oracle's golden vectors.
Bar: neighbour lists (out_edg) and the out_fin_ edge set exact, weights bit-exact as float32 (the
ratings are integer / half-star, so the float accumulators are order independent), the dense table
bit-exact after the 6-digit text round trip, knn3 average MSE within 1e-6 relative."""
import json
import os

import numpy as np
import pytest

from oracle import gsi_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from collaborative_filtering_b200.api import Context
    c = Context(0)
    yield c
    c.close()


def _csr(triples):
    by = {}
    for u, m, r in triples:
        by.setdefault(u, {})[m] = r
    users = sorted(by)
    offsets = np.zeros(len(users) + 1, dtype=np.int64)
    items, rat = [], []
    for i, u in enumerate(users):
        ms = sorted(by[u])
        items += ms
        rat += [by[u][m] for m in ms]
        offsets[i + 1] = len(items)
    return users, offsets, np.array(items, dtype=np.int32), np.array(rat, dtype=np.float32)


@pytest.mark.parametrize("case", ["tiny_int", "tiny_half"])
def test_knn_chain_matches_golden(ctx, golden_dir, case):
    d = os.path.join(golden_dir, case)
    z = np.load(os.path.join(d, "oracle.npz"))
    meta = json.load(open(os.path.join(d, "meta.json")))
    train, val = O.read_rating_files(os.path.join(d, "movielens"))
    rows = z["weights"].shape[0]
    _, t_off, t_items, t_rat = _csr(train)
    a, b, w = ctx.knn_build(t_off, t_items, t_rat, rows, install_weights=True)
    assert np.array_equal(a, z["fin_a"]) and np.array_equal(b, z["fin_b"])          # edge set exact
    assert np.array_equal(w, z["fin_w"].astype(np.float32))                        # float bit-exact
    # the text a host writes from these edges is the oracle's out_fin_ byte for byte
    text = "".join("%d %d %s\n" % (x, y, O.fmt_g(v)) for x, y, v in zip(a, b, w))
    assert text == open(os.path.join(d, "out_fin_1_of_1")).read()
    # installed table == what precompute_local parses back from that text
    import torch
    import ctypes
    p, r = ctypes.c_void_p(), ctypes.c_int(0)
    ctx._check(ctx._lib.gsi_get_weights(ctx._h, ctypes.byref(p), ctypes.byref(r)))
    assert r.value == rows
    recs = ctx.precompute(z["offsets"], z["items"])
    assert np.array_equal(recs.sig_min, z["sig_min"]) and np.array_equal(recs.k, z["k"])
    # knn step 1: co-rated lists through train and validate users
    _, a_off, a_items, _ = _csr(train + val)
    co = ctx.knn_corated(a_off, a_items, rows)
    edg_text = "".join("%d %s\n" % (m, "".join("%d " % j for j in np.nonzero(co[m])[0]))
                       for m in sorted({t[1] for t in train + val}))
    assert edg_text == open(os.path.join(d, "out_edg_1_of_1")).read()
    # knn3
    _, v_off, v_items, v_rat = _csr(val)
    mse, err, cnt, has = ctx.knn3(v_off, v_items, v_rat)
    assert abs(mse - meta["knn3_avg_mse"]) <= 1e-6 * max(1.0, abs(meta["knn3_avg_mse"]))


def test_knn_build_vs_oracle_ml100k_sample(ctx):
    from collaborative_filtering_b200 import datasets as D
    r = D.make_ratings("ml-100k", n_users=300)
    u, it, rat = r.triples()
    train = [(int(a), int(b), float(c)) for a, b, c in zip(u, it, rat)]
    ratm, _, edg = O.knn1(train, [])
    fin = O.knn2(ratm, edg)
    a, b, w = ctx.knn_build(r.offsets, r.items, r.ratings, r.n_items + 1, install_weights=False)
    assert len(fin) == len(a)
    assert np.array_equal(a, np.array([e[0] for e in fin])) and np.array_equal(b, np.array([e[1] for e in fin]))
    assert np.array_equal(w, np.array([e[2] for e in fin], dtype=np.float32))
