"""Workflow tools: fold generator with the semantics of fold_cross_validation.py:11-56,
RMSE aggregation over out_res lines (local_calc_precomp.cpp:393-404), and the per-fold pipeline of
run_test_precompute.sh:9-20 over the drop-in binaries."""
import math
import os

import numpy as np
import pytest

from collaborative_filtering_b200 import datasets as D
from collaborative_filtering_b200 import workflow as WF


def _write_udata(path, r):
    users, items, ratings = r.triples()
    with open(path, "w") as f:
        for u, i, x in zip(users, items, ratings):
            f.write("%d\t%d\t%d\n" % (u, i, int(x)))


def _read(path):
    rows = []
    with open(path) as f:
        for line in f:
            u, i, x = line.rstrip("\n").split("\t")
            rows.append((int(u), int(i), int(x)))
    return rows


def test_fold_cross_validation_semantics(tmp_path):
    r = D.make_ratings("ml-100k", n_users=57)
    src = str(tmp_path / "u.data")
    _write_udata(src, r)
    everything = sorted(_read(src))
    out = str(tmp_path / "cross_validation")
    nf = WF.fold_cross_validation(src, 5, out, seed=11)
    assert nf == 5                                           # 57 users: a fold is cut after 12 users (> 57/5), the last holds 9
    tests = [_read(os.path.join(out, "u%d.test" % i)) for i in range(nf)]
    users = [set(u for u, _, _ in t) for t in tests]
    assert [len(s) for s in users] == [12, 12, 12, 12, 9]
    for i in range(nf):
        for j in range(i + 1, nf):
            assert not (users[i] & users[j])                 # user-disjoint folds
    assert sorted(sum(tests, [])) == everything
    for i in range(nf):
        train = _read(os.path.join(out, "u%d.train" % i))
        assert sorted(train + tests[i]) == everything        # train = the other folds
        assert not (set(u for u, _, _ in train) & users[i])
    # a user's ratings stay in file order; the shuffle is seeded
    out2 = str(tmp_path / "again")
    WF.fold_cross_validation(src, 5, out2, seed=11)
    assert open(os.path.join(out, "u3.train")).read() == open(os.path.join(out2, "u3.train")).read()
    out3 = str(tmp_path / "other_seed")
    WF.fold_cross_validation(src, 5, out3, seed=12)
    assert open(os.path.join(out, "u0.test")).read() != open(os.path.join(out3, "u0.test")).read()
    with pytest.raises(FileExistsError):                     # os.mkdir in the reference
        WF.fold_cross_validation(src, 5, out, seed=11)


def test_fold_count_when_divisible(tmp_path):
    # 10 users, 5 folds: the cut needs MORE than 2 users, so folds hold 3,3,3,1 (fold_cross_validation.py:40)
    src = str(tmp_path / "u.data")
    with open(src, "w") as f:
        for u in range(1, 11):
            f.write("%d\t%d\t%d\n" % (u, 1, 3))
    nf = WF.fold_cross_validation(src, 5, str(tmp_path / "cv"), seed=1)
    sizes = [len(_read(str(tmp_path / "cv" / ("u%d.test" % i)))) for i in range(nf)]
    assert nf == 4 and sizes == [3, 3, 3, 1]


def test_rmse_from_out_res(tmp_path):
    p1, p2 = str(tmp_path / "out_res_1_of_2"), str(tmp_path / "out_res_2_of_2")
    open(p1, "w").write("3 2147483000 0.25 4\n7 2147483001 nan 0\n")
    open(p2, "w").write("9 2147483002 2.25 6\n9 2147483003 1 2\n\n")
    s = WF.rmse_from_out_res([p1, p2])
    assert s["predictions"] == 3 and s["nan"] == 1 and s["files"] == 2
    assert math.isclose(s["mse"], (0.25 + 2.25 + 1.0) / 3) and math.isclose(s["rmse"], math.sqrt(3.5 / 3))
    assert math.isclose(s["mean_kk"], 4.0)
    assert WF.rmse_from_out_res(str(tmp_path / "out_res_*"))["predictions"] == 3
    assert WF.main(["rmse", p1]) == 0


@pytest.mark.gpu
def test_pipeline_two_folds(tmp_path):
    r = D.make_ratings("ml-100k", n_users=40)
    src = str(tmp_path / "u.data")
    _write_udata(src, r)
    cv = str(tmp_path / "cross_validation")
    nf = WF.fold_cross_validation(src, 4, cv, seed=3)
    assert nf == 4
    work = str(tmp_path / "work")
    res = WF.run_pipeline(cv, work, folds=[0, 2], pct=100, seed=5, log=open(os.devnull, "w"))
    assert [x["fold"] for x in res] == [0, 2]
    for x in res:
        assert os.path.exists(os.path.join(work, "out_res.%d" % x["fold"]))
        n_test = len(_read(os.path.join(cv, "u%d.test" % x["fold"])))
        assert x["predictions"] + x["nan"] == n_test         # --pct 100: one line per validation rating
        assert x["predictions"] > 0 and 0.0 <= x["rmse"] <= 4.0
    # seeded --pct: same sample, same file
    a = WF.run_pipeline(cv, work, folds=[2], pct=50, seed=5, log=open(os.devnull, "w"))[0]
    text = open(os.path.join(work, "out_res.2")).read()
    b = WF.run_pipeline(cv, work, folds=[2], pct=50, seed=5, log=open(os.devnull, "w"))[0]
    strip = lambda d: {k: v for k, v in d.items() if k != "seconds"}      # wall seconds of the tools differ from run to run
    assert text == open(os.path.join(work, "out_res.2")).read() and strip(a) == strip(b)
    assert set(a["seconds"]) == {"knn", "knn2", "precompute_local", "local_calc_precomp"}
    assert 0 < a["predictions"] + a["nan"] < res[1]["predictions"] + res[1]["nan"]


def test_run_pipeline_rejects_unknown_variant(tmp_path):
    with pytest.raises(ValueError):
        WF.run_pipeline(str(tmp_path), str(tmp_path / "w"), folds=[0], variant="nope")


def test_mega_graph_generator(tmp_path):
    """mega_graph.py:27-40: conn * size^2 distinct directed links without self links, weights with two decimals, one
    signal line per vertex; seeded, so two runs write identical files."""
    d = str(tmp_path / "g")
    n = WF.mega_graph(300, 0.02, d, seed=4)
    assert n == int(0.02 * 300 * 300)
    topo = [l.split() for l in open(os.path.join(d, "graph_topology.txt"))]
    assert len(topo) == n and len({(a, b) for a, b, _ in topo}) == n
    assert all(a != b and 1 <= int(a) <= 300 and 1 <= int(b) <= 300 for a, b, _ in topo)
    assert all(len(w.split(".")[1]) == 2 and 0.0 <= float(w) <= 1.0 for _, _, w in topo)
    sig = [l.split() for l in open(os.path.join(d, "graph_signal.txt"))]
    assert [int(v) for v, _ in sig] == list(range(1, 301)) and all(0.0 <= float(x) < 10.0 for _, x in sig)
    first = open(os.path.join(d, "graph_topology.txt")).read()
    WF.mega_graph(300, 0.02, d, seed=4)
    assert first == open(os.path.join(d, "graph_topology.txt")).read()
    with pytest.raises(ValueError):
        WF.mega_graph(10, 1.0, d)


def test_cheby_scale_driver_with_a_stub_tool(tmp_path):
    """scale2.sh:5-36: the sweep driver around the cheby tool -- checked on CPU with a stand-in executable that prints the
    two timing lines the script greps and records the inputs it was given."""
    bin_dir = tmp_path / "bin"
    bin_dir.mkdir()
    stub = bin_dir / "cheby"
    stub.write_text("#!/bin/sh\n"
                    "echo \"$(wc -l < graph_topology.txt) $(wc -l < graph_signal.txt) $(wc -w < coeff.txt)\" >> calls.txt\n"
                    "echo 'Loading graph. Finished in 0.25'\n"
                    "echo 'Final Runtime (seconds):   0.5'\n")
    stub.chmod(0o755)
    work = str(tmp_path / "work")
    res = WF.cheby_scale(work, nodes=200, conn=0.01, coeffs=6, sweep_coeffs=(3, 5), sweep_conn=(0.01, 0.02),
                         sweep_nodes=(100, 200), seed=9, bin_dir=str(bin_dir), log=open(os.devnull, "w"))
    assert [(r["nodes"], r["conn"], r["coeffs"]) for r in res] == [(200, 0.01, 3), (200, 0.01, 5), (200, 0.01, 6), (200, 0.02, 6),
                                                                    (100, 0.01, 6), (200, 0.01, 6)]
    assert all(r["load_s"] == 0.25 and r["runtime_s"] == 0.5 for r in res)
    calls = [tuple(int(x) for x in l.split()) for l in open(os.path.join(work, "calls.txt"))]
    assert calls == [(400, 200, 3), (400, 200, 5), (400, 200, 6), (800, 200, 6), (100, 100, 6), (400, 200, 6)]
    lines = open(os.path.join(work, "scale_res2.jsonl")).read().splitlines()
    assert len(lines) == 6
