"""GPU: the approximate Stage-1 mode (ts_ivf_*, csrc/ivf.cu -- the stand-in for the reference's
``faiss.IndexIVFFlat`` branch, /root/reference/src/stage1_retriever.py:262-273) through the C ABI
against ``oracle/ivf.py`` on the same seeded inputs, at the reference's nlist = 100 / nprobe = 10.
The same kernels run at small sizes on the SIMT emulator in tests/test_ivf.py.

Tolerance: ids exact except near-ties within 1e-3 relative score, scores within 1e-3 relative
(bf16 storage, fp32 accumulation) -- the Stage-1 rule of BASELINE.json."""
import os

import numpy as np
import pytest

from test_ivf import clustered

from oracle import flat_ip
from oracle import ivf as oivf
from tristage_rag_b200 import _lib
from tristage_rag_b200 import ivf as ivf_train

pytestmark = pytest.mark.gpu
REL = 1e-3


@pytest.fixture(scope="module")
def corpus(request, cuda_device):
    small = request.config.getoption("--emulate")
    N, d, nlist = (3000, 64, 12) if small else (200_000, 256, 100)
    if os.environ.get("TS_TEST_IVF_SIZE"):                     # "N,d,nlist": rehearse other sizes (e.g. on the emulator)
        N, d, nlist = (int(v) for v in os.environ["TS_TEST_IVF_SIZE"].split(","))
    X, centers = clustered(N, d, nlist, seed=17)
    idx = _lib.Index(d, "bf16", "ip", cuda_device)
    idx.add(X[: N // 2])
    iv = _lib.IVF(idx, nlist)
    cent = ivf_train.train_centroids(X[: N // 2], nlist)       # trained on the first batch, as the reference does
    iv.set_centroids(cent)
    iv.sync()
    idx.add(X[N // 2:])                                        # later batch: assigned only
    iv.sync()
    return X, centers, idx, iv, cent, flat_ip.round_to(X, "bf16")


def _queries(centers, B, d, seed):
    rng = np.random.default_rng(seed)
    q = centers[rng.integers(0, len(centers), size=B)] + 0.3 * rng.standard_normal((B, d)).astype(np.float32) / np.sqrt(d)
    return flat_ip.normalize_rows(q).astype(np.float32)


def test_assignments_match_oracle(corpus):
    X, centers, idx, iv, cent, Xr = corpus
    got = iv.assignments()
    assert got.shape == (len(X),) and iv.nassigned == len(X)
    want, margin = oivf.assign_lists(Xr, cent), oivf.assign_margin(Xr, cent)
    bad = np.nonzero((got != want) & (margin > 1e-5))[0]
    assert bad.size == 0, (bad[:5], got[bad[:5]], want[bad[:5]])
    assert np.array_equal(iv.list_sizes(), np.bincount(got, minlength=iv.nlist))


@pytest.mark.parametrize("B,k", [(1, 100), (32, 100), (5, 500), (200, 10)])
def test_search_matches_oracle_on_the_probed_lists(corpus, B, k):
    X, centers, idx, iv, cent, Xr = corpus
    nprobe = max(1, iv.nlist // 10)
    Q = _queries(centers, B, X.shape[1], seed=B + k)
    lists, lscores = iv.coarse_host(Q, nprobe)
    olists, oscores = oivf.coarse_probe(Q, cent, nprobe)
    assert np.allclose(lscores, oscores, rtol=1e-4, atol=1e-5)
    assert (np.sort(lists, axis=1) == np.sort(olists, axis=1)).mean() > 0.99      # swaps only between near-tied centroids
    a = iv.assignments()
    D, I = iv.search_host(Q, k, nprobe)
    Qr = flat_ip.round_to(Q, "bf16")
    rD, rI = oivf.ivf_search(Xr, Qr, a, lists, k)
    sc = lambda b, ids: Xr[ids].astype(np.float64) @ Qr[b].astype(np.float64)   # noqa: E731
    assert not flat_ip.check_topk(D, I, sc, rD, rI, rel=REL)
    for b in range(B):
        ok = I[b] >= 0
        assert np.isin(a[I[b][ok]], lists[b]).all()
        assert (D[b][~ok] == np.float32(flat_ip.LOWEST_F32)).all()


def test_all_lists_probed_is_the_exact_search_and_recall_is_high(corpus):
    X, centers, idx, iv, cent, Xr = corpus
    Q = _queries(centers, 16, X.shape[1], seed=99)
    k = 50
    eD, eI = idx.search_host(Q, k)
    D, I = iv.search_host(Q, k, iv.nlist)
    Qr = flat_ip.round_to(Q, "bf16")
    sc = lambda b, ids: Xr[ids].astype(np.float64) @ Qr[b].astype(np.float64)   # noqa: E731
    assert not flat_ip.check_topk(D, I, sc, eD, eI, rel=REL)
    # a tenth of the lists on a clustered corpus finds nearly everything the exact scan finds
    aD, aI = iv.search_host(Q, k, max(1, iv.nlist // 10))
    recall = np.mean([len(set(aI[b]) & set(eI[b])) / k for b in range(len(Q))])
    assert recall > 0.8, recall


def test_device_pointer_api(corpus, request):
    if request.config.getoption("--emulate"):
        pytest.skip("needs torch CUDA tensors")
    import torch

    X, centers, idx, iv, cent, Xr = corpus
    Q = _queries(centers, 8, X.shape[1], seed=5)
    nprobe = max(1, iv.nlist // 10)
    D, I = iv.search_host(Q, 100, nprobe)
    s, i = iv.search(torch.from_numpy(Q).cuda(), 100, nprobe)
    assert np.array_equal(i.cpu().numpy(), I) and np.array_equal(s.cpu().numpy(), D)


def test_faiss_shaped_module_drives_both_index_kinds(cuda_device, tmp_path, monkeypatch):
    """tristage_rag_b200/faiss_compat.py (INTEGRATION.md way C) called the way the reference's _create_faiss_index /
    search / save_index / load_index call faiss (src/stage1_retriever.py:262-277,380,436,463)."""
    from tristage_rag_b200 import faiss_compat as faiss

    monkeypatch.setenv("TS_STORAGE_DTYPE", "bf16")
    monkeypatch.setenv("TS_GPU_INDEX", str(cuda_device))
    X, centers = clustered(5000, 64, 10, seed=23)
    Q = _queries(centers, 4, 64, seed=2)
    Xr, Qr = flat_ip.round_to(X, "bf16"), flat_ip.round_to(Q, "bf16")
    sc = lambda b, ids: Xr[ids].astype(np.float64) @ Qr[b].astype(np.float64)   # noqa: E731
    flat = faiss.IndexFlatIP(64)
    flat.add(X)
    D, I = flat.search(Q, 10)
    rD, rI = flat_ip.topk_desc(Qr @ Xr.T, 10)
    assert flat.ntotal == 5000 and not flat_ip.check_topk(D, I, sc, rD, rI, rel=REL)
    ivf = faiss.IndexIVFFlat(faiss.IndexFlatIP(64), 64, 10, faiss.METRIC_INNER_PRODUCT)
    ivf.train(X)
    ivf.add(X)
    ivf.nprobe = 10
    D2, I2 = ivf.search(Q, 10)
    assert ivf.is_trained and ivf.ntotal == 5000 and not flat_ip.check_topk(D2, I2, sc, rD, rI, rel=REL)
    ivf.nprobe = 2
    D3, I3 = ivf.search(Q, 10)
    for kind, index in (("flat", flat), ("ivf", ivf)):
        path = str(tmp_path / f"{kind}.index")
        faiss.write_index(index, path)
        back = faiss.read_index(path)
        assert type(back).__name__ == type(index).__mro__[1].__name__ and back.ntotal == 5000
        bD, bI = back.search(Q, 10)
        assert np.array_equal(bI, I if kind == "flat" else I3) and np.array_equal(bD, D if kind == "flat" else D3)
    with pytest.raises(ValueError):
        faiss.IndexIVFFlat(faiss.IndexFlatIP(32), 64, 10)
