"""Multi-GPU inside the product: the device group behind the C ABI (gsi_group_*) and the tools
that use it.  The reference's tool spreads the users over all workers of the box by itself
(precompute_local_threads.cpp:300-314).  On a single-GPU box the dealing / merging logic is exercised with the same
device listed twice (GSI_GROUP_ALLOW_DUP=1, small users only); the tests that need two devices skip there and run under
`gpurun --gpus 2`."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import gsi_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "collaborative_filtering_b200", "bin")


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _batch(sizes, n_items, seed):
    rng = np.random.default_rng(seed)
    offsets = np.zeros(len(sizes) + 1, dtype=np.int64)
    np.cumsum(sizes, out=offsets[1:])
    items = np.concatenate([np.sort(rng.choice(n_items, size=n, replace=False) + 1) for n in sizes]).astype(np.int32)
    return offsets, items


def _same_records(a, b):
    assert np.array_equal(a.k, b.k) and np.array_equal(a.sig_min, b.sig_min)
    for u in range(len(a.k)):
        assert np.array_equal(a.lam_of(u), b.lam_of(u)), "lam of user %d" % u
        assert np.array_equal(a.vec_of(u), b.vec_of(u)), "vec of user %d" % u


def _group_vs_single(devices, sizes, n_items=400):
    from collaborative_filtering_b200 import datasets as D
    from collaborative_filtering_b200.api import Context, Group
    w = D.make_weights(n_items)
    offsets, items = _batch(sizes, n_items, 17)
    rng = np.random.default_rng(5)
    ratings = rng.integers(1, 6, size=int(offsets[-1])).astype(np.float64)
    c = Context(devices[0])
    g = Group(devices)
    try:
        c.set_weights(w)
        g.set_weights(w)
        assert g.size == len(devices)
        one = c.precompute(offsets, items)
        many = g.precompute(offsets, items)
        _same_records(one, many)
        p1 = c.predict(one, ratings)
        p2 = g.predict(many, ratings)
        for key in ("kk", "status", "cols"):
            assert np.array_equal(p1[key], p2[key]), key
        assert np.array_equal(p1["pred"], p2["pred"], equal_nan=True) and np.array_equal(p1["err"], p2["err"], equal_nan=True)
        return g.broadcast_path
    finally:
        g.close()
        c.close()


def test_group_deal_and_merge_on_one_device(monkeypatch):
    """Two members on device 0: the LPT deal, the per-member threads, the sink serialisation, the user_index / sig_min
    remapping and the predictor's scatter -- records and predictions bit-identical to one context."""
    monkeypatch.setenv("GSI_GROUP_ALLOW_DUP", "1")
    path = _group_vs_single([0, 0], [160, 150, 120, 100, 90, 70, 60, 45, 44, 33, 20, 8, 3, 2, 1] + [25] * 40)
    assert path == "peer"


def test_group_refuses_a_device_twice():
    from collaborative_filtering_b200.api import Group, GsiError
    with pytest.raises(GsiError) as e:
        Group([0, 0])
    assert e.value.code == 1


def test_group_on_two_gpus():
    """Needs 2 devices.  W replicated device to device (NCCL broadcast, or peer copies), users up to n = 2,500 dealt over
    both GPUs: bit-identical to the single-device records (SURVEY.md section 7 test (h))."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    path = _group_vs_single([0, 1], [2500, 1500, 1100, 700, 300, 300, 130, 90, 50, 33, 20, 8, 3, 2, 1] + [60] * 50, n_items=3000)
    assert path in ("nccl", "peer")


def _run(tool, cwd, *args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([os.path.join(BIN, tool), *args], cwd=cwd, env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    assert p.returncode == 0, p.stdout.decode()
    return p.stdout.decode()


def _pipeline_inputs(tmp_path, golden_dir, case="tiny_int"):
    g = os.path.join(golden_dir, case)
    cwd = str(tmp_path)
    shutil.copytree(os.path.join(g, "movielens"), os.path.join(cwd, "movielens"))
    _run("knn", cwd)
    _run("knn2", cwd)
    return cwd


def test_out_eigen_is_written_in_ascending_user_order_and_feeds_the_predictor(tmp_path, golden_dir):
    """by default (B1 kept) local_calc_precomp takes every cutoff from the running concatenation of
    all records read so far (local_calc_precomp.cpp:414,437,440,271), so its output depends on the record ORDER of
    out_eigen_.  The tools write ascending user' (defined order); this test feeds the tool's OWN out_eigen_
    to the predictor with B1 on and off and compares with the oracle's reader + predictor on that same file."""
    cwd = _pipeline_inputs(tmp_path, golden_dir)
    _run("precompute_local_threads", cwd, "3", env={"GSI_DEVICE": "0"})
    path = os.path.join(cwd, "out_eigen_")
    users = [int(line.split()[0]) for i, line in enumerate(open(path)) if i % 3 == 0]
    assert users == sorted(users) and len(users) == len(set(users)) > 3
    fin = O.parse_fin(open(os.path.join(cwd, "out_fin_1_of_1")).read())
    test_rt = O.parse_rat(open(os.path.join(cwd, "out_test_rat_1_of_1")).read())
    graph = O.item_graph(fin)
    for b1, extra in ((True, []), (False, ["--fix-b1"])):
        _run("local_calc_precomp", cwd, "--pct", "100", *extra, env={"GSI_DEVICE": "0"})
        rows = {}
        for line in open(os.path.join(cwd, "out_res_1_of_1")):
            m, u, e, kk = line.split()
            rows[(int(m), int(u))] = (float(e), int(kk))
        ref = O.local_calc_precomp(O.parse_out_eigen(path, bug_b1=b1), graph, test_rt)
        assert len(ref) == len(rows) > 0
        se_g = se_o = 0.0
        ok = 0
        for (m, u, err, kk, pred, status, c) in ref:
            e, kg = rows[(m, u)]
            assert kg == kk
            if status == O.PRED_OK:
                assert abs(e - float(err)) <= 2e-5 * max(1.0, float(err)), (m, u, e, float(err), b1)
                se_g += e
                se_o += float(err)
                ok += 1
        assert ok > 0 and abs(np.sqrt(se_g / ok) - np.sqrt(se_o / ok)) <= 1e-4


def test_tools_on_two_gpus_write_the_same_files(tmp_path, golden_dir):
    """Needs 2 devices: precompute_local_threads and local_calc_precomp over GSI_DEVICES=0,1 produce byte-identical
    out_eigen_ / out_res_ to the single-device run (the file order is defined: ascending user')."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    cwd = _pipeline_inputs(tmp_path, golden_dir)
    out = _run("precompute_local_threads", cwd, "4", env={"GSI_DEVICES": "0,1"})
    assert "Devices: 2" in out
    two = open(os.path.join(cwd, "out_eigen_")).read()
    _run("local_calc_precomp", cwd, "--pct", "100", env={"GSI_DEVICES": "0,1"})
    res_two = open(os.path.join(cwd, "out_res_1_of_1")).read()
    _run("precompute_local_threads", cwd, "4", env={"GSI_DEVICE": "0"})
    assert open(os.path.join(cwd, "out_eigen_")).read() == two
    _run("local_calc_precomp", cwd, "--pct", "100", env={"GSI_DEVICE": "0"})
    assert open(os.path.join(cwd, "out_res_1_of_1")).read() == res_two
