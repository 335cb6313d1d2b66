"""CPU: the oracle against the golden vectors produced by the reference's own
code (oracle/gen_golden.py) and the two oracle restatements against each other."""
import json
import os

import numpy as np
import pytest

from oracle import c_oracle, flat_ip, maxsim


def _load(golden_dir, name):
    with open(os.path.join(golden_dir, name)) as f:
        return json.load(f)


def _case_inputs(c):
    r = np.random.default_rng(c["seed"])
    q = r.standard_normal((c["Lq"], c["H"])).astype(np.float32) * np.float32(c["q_scale"])
    d = r.standard_normal((c["Ld"], c["H"])).astype(np.float32) * np.float32(c["d_scale"])
    return q, d


def test_stage2_matches_reference_golden(golden_dir):
    g = _load(golden_dir, "stage2_reference.json")
    n = 0
    for c in g["cases"]:
        if c["kind"] != "numpy":
            continue
        q, d = _case_inputs(c)
        assert maxsim.maxsim_score(q, d) == pytest.approx(c["maxsim"], rel=1e-5, abs=1e-6)
        assert maxsim.colbert_score(q, d) == pytest.approx(c["colbert"], rel=1e-5, abs=1e-6)
        assert c_oracle.maxsim(q, d, 0) == pytest.approx(c["maxsim"], rel=1e-5, abs=1e-6)
        assert c_oracle.maxsim(q, d, 1) == pytest.approx(c["colbert"], rel=1e-5, abs=1e-6)
        n += 1
    assert n >= 10


def test_stage2_torch_seed0_probe(golden_dir):
    """torch seed-0 randn(1,32,128) x randn(1,180,128)."""
    import torch

    g = _load(golden_dir, "stage2_reference.json")
    c = [c for c in g["cases"] if c["kind"] == "torch_seed0"][0]
    torch.manual_seed(0)
    q = torch.randn(1, 32, 128).numpy()[0]
    d = torch.randn(1, 180, 128).numpy()[0]
    assert maxsim.maxsim_score(q, d) == pytest.approx(c["maxsim"], rel=1e-5)
    assert maxsim.colbert_score(q, d) == pytest.approx(c["colbert"], rel=1e-5)
    assert c["maxsim"] == pytest.approx(0.24198836, rel=1e-6)
    assert c["colbert"] == pytest.approx(0.24319050, rel=1e-6)


def test_normalize_matches_reference_golden(golden_dir):
    g = _load(golden_dir, "stage1_normalize.json")
    x = (np.random.default_rng(g["seed"]).standard_normal(g["shape"]) * g["scale"]).astype(np.float32)
    x[g["zero_row"]] = 0.0
    y = flat_ip.normalize_rows(x)
    assert str(y.dtype) == g["dtype"]
    np.testing.assert_allclose(y, np.array(g["y"], np.float32), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(c_oracle.normalize_rows(x), np.array(g["y"], np.float32), rtol=1e-6, atol=1e-7)
    assert np.isfinite(y).all() and (y[g["zero_row"]] == 0).all()


def test_flat_ip_golden_and_semantics(golden_dir):
    g = _load(golden_dir, "stage1_flat_ip.json")
    rng = np.random.default_rng(g["seed"])
    x = flat_ip.normalize_rows(rng.standard_normal((g["n"], g["d"])).astype(np.float32)).astype(np.float32)
    q = flat_ip.normalize_rows(rng.standard_normal((g["nq"], g["d"])).astype(np.float32)).astype(np.float32)
    x[g["dup"][0]] = x[g["dup"][1]]
    idx = flat_ip.IndexFlatIP(g["d"])
    idx.add(x[:17])
    idx.add(x[17:])                                   # multi-add == single add
    D, I = idx.search(q, 5)
    assert I.tolist() == g["k5"]["I"]
    np.testing.assert_allclose(D, np.array(g["k5"]["D"], np.float32), rtol=1e-6)
    assert I.dtype == np.int64 and D.dtype == np.float32
    assert (np.diff(D, axis=1) <= 0).all()
    D2, I2 = idx.search(q, 50)                        # k > ntotal: -1 labels, lowest-float scores
    assert int((I2[0] >= 0).sum()) == g["k50_valid"] == g["n"]
    assert I2[0].tolist() == g["k50_I0"]
    assert D2[0, -1] == np.float32(g["k50_pad_score"]) == flat_ip.LOWEST_F32
    # duplicate rows tie exactly: lower id first
    for b in range(len(q)):
        row = I2[b, : g["n"]].tolist()
        assert row.index(g["dup"][1]) < row.index(g["dup"][0])


@pytest.mark.parametrize("n,d,B,k", [(1, 8, 1, 3), (257, 24, 5, 10), (5000, 64, 4, 100), (3000, 32, 2, 500)])
def test_numpy_and_c_oracles_agree(n, d, B, k):
    rng = np.random.default_rng(n + d)
    X = flat_ip.normalize_rows(rng.standard_normal((n, d)).astype(np.float32)).astype(np.float32)
    Q = flat_ip.normalize_rows(rng.standard_normal((B, d)).astype(np.float32)).astype(np.float32)
    idx = flat_ip.IndexFlatIP(d)
    idx.add(X)
    D, I = idx.search(Q, k)
    Db, Ib = idx.search(Q, k, block=701)              # blocked scan == one-shot
    Dc, Ic = c_oracle.flat_ip_search(X, Q, k)
    assert (I == Ib).all()
    S = Q @ X.T
    assert not flat_ip.check_topk(Dc, Ic, lambda b, ids: S[b, ids], D, I, rel=1e-5)
    assert not flat_ip.check_topk(D, I, lambda b, ids: S[b, ids], D, I, rel=1e-6)


def test_check_topk_flags_real_errors():
    rng = np.random.default_rng(3)
    X = rng.standard_normal((500, 16)).astype(np.float32)
    Q = rng.standard_normal((2, 16)).astype(np.float32)
    S = Q @ X.T
    D, I = flat_ip.topk_desc(S, 10)
    sc = lambda b, ids: S[b, ids]                     # noqa: E731
    assert not flat_ip.check_topk(D, I, sc, D, I)
    bad_I = I.copy()
    bad_I[0, 3] = int(np.argmin(S[0]))                # a far-away row
    assert flat_ip.check_topk(D, bad_I, sc, D, I)
    bad_D = D.copy()
    bad_D[1, 0] *= 1.01
    assert flat_ip.check_topk(bad_D, I, sc, D, I)
    swapped = I.copy()
    swapped[0, [0, 9]] = swapped[0, [9, 0]]           # order inversion far outside the tie band
    assert flat_ip.check_topk(D, swapped, sc, D, I)


def test_rescore_order_is_stable():
    s = np.array([0.5, 0.9, 0.5, 0.9, 0.1], np.float32)
    assert maxsim.rescore_order(s, 4).tolist() == [1, 3, 0, 2]


def test_maxsim_batch_matches_loop():
    rng = np.random.default_rng(11)
    q = rng.standard_normal((32, 64)).astype(np.float32)
    lens = rng.integers(1, 40, size=20)
    tok = rng.standard_normal((int(lens.sum()), 64)).astype(np.float32)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    for mode in (0, 1):
        got = c_oracle.maxsim_batch(q, tok, off, mode)
        ref = maxsim.score_candidates(q, [tok[off[i]:off[i + 1]] for i in range(len(lens))], mode)
        np.testing.assert_allclose(got, ref, rtol=2e-5, atol=2e-6)


def test_ivf_golden_and_semantics(golden_dir):
    """oracle/ivf.py against its committed vectors (FAISS is absent: this pins the restated IndexIVFFlat against
    itself) and the properties the kernels are tested for: every row in exactly one list, probes best first,
    results only from probed lists, -1 padding past the probed rows, all lists probed == the flat index."""
    import json

    from oracle import ivf as oivf

    with open(os.path.join(golden_dir, "stage1_ivf.json")) as f:
        g = json.load(f)
    x, q = np.asarray(g["x"], np.float32), np.asarray(g["q"], np.float32)
    cent = oivf.kmeans_ip(x, g["nlist"], niter=10, seed=1234)
    assert np.allclose(cent, np.asarray(g["centroids"], np.float32), atol=1e-6)
    assign = oivf.assign_lists(x, cent)
    assert assign.tolist() == g["assign"] and assign.min() >= 0 and assign.max() < g["nlist"]
    lists, lscores = oivf.coarse_probe(q, cent, g["nprobe"])
    assert lists.tolist() == g["lists"] and np.allclose(lscores, g["lscores"], atol=1e-6)
    assert (np.diff(lscores, axis=1) <= 0).all()
    D, I = oivf.ivf_search(x, q, assign, lists, 8)
    assert I.tolist() == g["k8"]["I"] and np.allclose(D, g["k8"]["D"], atol=1e-6)
    for b in range(2):
        ok = I[b] >= 0
        assert np.isin(assign[I[b][ok]], lists[b]).all() and (np.diff(D[b][ok]) <= 0).all()
    Dp, Ip = oivf.ivf_search(x, q, assign, lists[:, :1], 40)
    assert [int((Ip[b] >= 0).sum()) for b in range(2)] == g["one_list_k40_valid"] and Ip[0].tolist() == g["one_list_k40_I0"]
    n0 = g["one_list_k40_valid"][0]
    assert n0 == int((assign == lists[0, 0]).sum()) and (Ip[0, n0:] == -1).all() and (Dp[0, n0:] == flat_ip.LOWEST_F32).all()
    all_lists = np.tile(np.arange(g["nlist"], dtype=np.int32), (2, 1))
    De, Ie = oivf.ivf_search(x, q, assign, all_lists, 8)
    rD, rI = flat_ip.topk_desc(q @ x.T, 8)
    assert (Ie == rI).all() and np.allclose(De, rD, atol=1e-6)
