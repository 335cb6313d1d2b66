"""CPU: the approximate Stage-1 mode (csrc/ivf.cu -- SURVEY.md §8f-4, the stand-in for the reference's
``faiss.IndexIVFFlat`` branch, /root/reference/src/stage1_retriever.py:262-273) EXECUTED on the SIMT
emulator (tests/cudasim) and checked against ``oracle/ivf.py``; plus the host-side trainer and the
reader for the index file the reference writes in that mode.  Test infrastructure only: the emulated
library is never used by the package."""
import numpy as np
import pytest

from oracle import flat_ip
from oracle import ivf as oivf
from tristage_rag_b200 import _lib
from tristage_rag_b200 import ivf as ivf_train

REL = 1e-3


@pytest.fixture(autouse=True)
def small_emulated_gpu(monkeypatch):
    """16 SMs instead of 148: the scan grid (~4 CTAs per SM, cut into list segments) stays small enough to emulate
    quickly while lists are still split into several segments."""
    monkeypatch.setenv("HOSTSIM_SM_COUNT", "16")


def clustered(N, d, n_clusters, seed, spread=0.35):
    """Unit rows around n_clusters random directions (a corpus IVF lists make sense for)."""
    rng = np.random.default_rng(seed)
    centers = flat_ip.normalize_rows(rng.standard_normal((n_clusters, d)).astype(np.float32))
    which = rng.integers(0, n_clusters, size=N)
    X = centers[which] + spread * rng.standard_normal((N, d)).astype(np.float32) / np.sqrt(d)
    return flat_ip.normalize_rows(X).astype(np.float32), centers.astype(np.float32)


def build(X, nlist, dtype, seed=0):
    idx = _lib.Index(X.shape[1], dtype, "ip", 0)
    idx.add(X)
    iv = _lib.IVF(idx, nlist)
    cent = ivf_train.train_centroids(X, nlist, seed=seed)
    iv.set_centroids(cent)
    iv.sync()
    return idx, iv, cent


def test_trainer_matches_the_oracle_restatement_and_fills_every_list():
    X, _ = clustered(4000, 24, 9, seed=3)
    cent = ivf_train.train_centroids(X, 16)
    assert cent.shape == (16, 24) and cent.dtype == np.float32 and np.isfinite(cent).all()
    assert np.array_equal(cent, oivf.kmeans_ip(X, 16))
    a = oivf.assign_lists(X, cent)
    assert len(np.unique(a)) >= 9                      # at least the real clusters are populated
    # subsampling cap: more than 256 points per centroid -> trains on a subset, still deterministic
    c2 = ivf_train.train_centroids(X, 4, max_points_per_centroid=100)
    assert np.array_equal(c2, oivf.kmeans_ip(X, 4, max_points_per_centroid=100))
    with pytest.raises(ValueError):
        ivf_train.train_centroids(X[:3], 4)


def test_trainer_reseeds_empty_clusters():
    rng = np.random.default_rng(1)
    X = np.repeat(flat_ip.normalize_rows(rng.standard_normal((3, 8)).astype(np.float32)), 40, axis=0).astype(np.float32)
    cent = ivf_train.train_centroids(X, 6, niter=4)     # 3 distinct points, 6 lists: splits must happen
    assert np.isfinite(cent).all() and len(np.unique(np.round(cent, 6), axis=0)) >= 3


@pytest.mark.parametrize("N,d,nlist,dtype", [(1500, 40, 12, "bf16"), (900, 64, 7, "fp16"), (1100, 20, 5, "fp32"),
                                             (300, 136, 33, "bf16")])
def test_assign_kernel_matches_oracle(sim, N, d, nlist, dtype):
    X, _ = clustered(N, d, max(3, nlist // 2), seed=N)
    idx, iv, cent = build(X, nlist, dtype)
    assert iv.is_trained and iv.nassigned == N
    got = iv.assignments()
    Xr = flat_ip.round_to(X, dtype)
    want = oivf.assign_lists(Xr, cent)
    margin = oivf.assign_margin(Xr, cent)
    bad = np.nonzero((got != want) & (margin > 1e-5))[0]
    assert bad.size == 0, (bad[:5], got[bad[:5]], want[bad[:5]])
    assert np.array_equal(iv.list_sizes(), np.bincount(got, minlength=nlist))
    assert np.allclose(iv.centroids(), cent)


@pytest.mark.parametrize("N,d,B,k,nlist,nprobe,dtype", [
    (1500, 40, 3, 10, 12, 3, "bf16"),
    (1200, 64, 2, 100, 8, 2, "fp16"),
    (1100, 24, 4, 20, 6, 6, "fp32"),        # nprobe == nlist: the exact result
    (700, 72, 1, 128, 10, 1, "bf16"),       # one probed list, k larger than most lists -> -1 padding
    (600, 32, 1, 500, 9, 4, "bf16"),        # k = 500 takes the 1024-entry candidate lists
])
def test_search_matches_oracle_on_the_probed_lists(sim, monkeypatch, N, d, B, k, nlist, nprobe, dtype):
    if B * nprobe <= 4:
        monkeypatch.setenv("HOSTSIM_SM_COUNT", "40")       # 160 segments per list, most of them empty here
    X, centers = clustered(N, d, nlist, seed=N + k)
    rng = np.random.default_rng(N)
    Q = flat_ip.normalize_rows(centers[rng.integers(0, nlist, size=B)]
                               + 0.3 * rng.standard_normal((B, d)).astype(np.float32) / np.sqrt(d)).astype(np.float32)
    idx, iv, cent = build(X, nlist, dtype)
    lists, lscores = iv.coarse_host(Q, nprobe)
    # coarse step: the oracle's lists, except where two centroids score within rounding of each other
    olists, oscores = oivf.coarse_probe(Q, cent, nprobe)
    assert np.allclose(lscores, oscores, rtol=1e-5, atol=1e-6)
    for b in range(B):
        if not np.array_equal(lists[b], olists[b]):
            full = np.sort((Q[b].astype(np.float64) @ cent.astype(np.float64).T))[::-1]
            assert np.min(np.abs(np.diff(full))) < 1e-6, (lists[b], olists[b])
    before = sim.cudasim_launches()
    D, I = iv.search_host(Q, k, nprobe)
    assert sim.cudasim_launches() - before >= 5      # 2 x query prep, coarse, scan, selection
    Xr, Qr = flat_ip.round_to(X, dtype), flat_ip.round_to(Q, dtype)
    rD, rI = oivf.ivf_search(Xr, Qr, iv.assignments(), lists, k)
    sc = lambda b, ids: Xr[ids].astype(np.float64) @ Qr[b].astype(np.float64)   # noqa: E731
    assert not flat_ip.check_topk(D, I, sc, rD, rI, rel=REL)
    scanned = np.array([np.isin(iv.assignments(), lists[b]).sum() for b in range(B)])
    for b in range(B):
        n = min(k, scanned[b])
        assert (I[b, n:] == -1).all() and (D[b, n:] == np.float32(flat_ip.LOWEST_F32)).all()
        assert (I[b, :n] >= 0).all() and np.isin(iv.assignments()[I[b, :n]], lists[b]).all()
    if nprobe == nlist:
        eD, eI = idx.search_host(Q, k, path="stream")
        assert not flat_ip.check_topk(D, I, sc, eD, eI, rel=REL)


def test_incremental_add_autosync_reset_and_id_base(sim):
    X, centers = clustered(700, 32, 6, seed=8)
    idx = _lib.Index(32, "bf16", "ip", 0)
    iv = _lib.IVF(idx, 6)
    Q = centers[:2].copy()
    with pytest.raises(_lib.TristageError, match="no centroids"):
        iv.search_host(Q, 5, 2)
    idx.add(X[:501])                                     # the reference trains on its first batch
    cent = ivf_train.train_centroids(X[:501], 6)
    iv.set_centroids(cent)
    D1, I1 = iv.search_host(Q, 10, 2)                    # syncs by itself
    assert iv.nassigned == 501 and I1.max() < 501
    a1 = iv.assignments()
    idx.add(X[501:])                                     # later batches are only assigned (:313)
    D2, I2 = iv.search_host(Q, 10, 6)
    assert iv.nassigned == 700
    Xr, Qr = flat_ip.round_to(X, "bf16"), flat_ip.round_to(Q, "bf16")
    rD, rI = flat_ip.topk_desc(Qr @ Xr.T, 10)            # all lists probed == exact
    sc = lambda b, ids: Xr[ids].astype(np.float64) @ Qr[b].astype(np.float64)   # noqa: E731
    assert not flat_ip.check_topk(D2, I2, sc, rD, rI, rel=REL)
    a = iv.assignments()
    assert a.shape == (700,) and np.array_equal(a[:501], a1)        # earlier rows keep their lists
    idx.set_id_base(5_000_000_000)
    D3, I3 = iv.search_host(Q, 10, 6)
    assert (I3 == I2 + 5_000_000_000).all() and np.array_equal(D3, D2)
    idx.set_id_base(0)
    # saved lists can be installed again (load_index): same results without the assign kernel
    iv2 = _lib.IVF(idx, 6)
    iv2.set_centroids(cent)
    iv2.set_assignments(a)
    launches = iv2.launches
    D4, I4 = iv2.search_host(Q, 10, 6)
    assert np.array_equal(I4, I2) and np.array_equal(D4, D2) and iv2.launches > launches
    with pytest.raises(_lib.TristageError):
        iv2.set_assignments(a[:-1])
    bad = a.copy()
    bad[0] = 6
    with pytest.raises(_lib.TristageError, match="names list"):
        iv2.set_assignments(bad)
    idx.reset()
    with pytest.raises(_lib.TristageError) as e:
        iv.search_host(Q, 5, 2)
    assert e.value.code == _lib.TS_ERR_EMPTY and "No documents indexed" in e.value.msg
    idx.add(X[:50])
    D5, I5 = iv.search_host(Q, 5, 6)                     # lists rebuilt for the new contents
    assert iv.nassigned == 50 and I5.max() < 50 and (I5 >= 0).all()
    idx.reset()
    idx.add(X[100:400])                                  # MORE rows than before the reset: still nothing stale
    D6, I6 = iv.search_host(Q, 5, 6)
    assert iv.nassigned == 300
    Xs = flat_ip.round_to(X[100:400], "bf16")
    rD6, rI6 = flat_ip.topk_desc(Qr @ Xs.T, 5)
    assert not flat_ip.check_topk(D6, I6, lambda b, ids: Xs[ids].astype(np.float64) @ Qr[b].astype(np.float64), rD6, rI6, rel=REL)
    with pytest.raises(_lib.TristageError):
        _lib.IVF(idx, 0)
    with pytest.raises(_lib.TristageError):
        _lib.IVF(idx, 5000)


@pytest.mark.parametrize("order", ["ascending", "descending", "constant"])
def test_adversarial_orders_overflow_the_per_warp_lists(sim, monkeypatch, order):
    """One list, one segment per probe (a 1-SM "GPU"), scores ascending along the list: every row beats the
    running threshold, so each warp's 256-entry candidate list overflows and is pruned in the scan."""
    monkeypatch.setenv("HOSTSIM_SM_COUNT", "1")
    N, d, k = 2600, 32, 100
    rng = np.random.default_rng(4)
    u = flat_ip.normalize_rows(rng.standard_normal((1, d)).astype(np.float32))[0].astype(np.float32)
    v = {"ascending": np.linspace(0.05, 1.0, N), "descending": np.linspace(1.0, 0.05, N),
         "constant": np.full(N, 0.5)}[order].astype(np.float32)
    X = (v[:, None] * u[None, :]).astype(np.float32)
    Q = np.stack([u, u, -u, -u]).astype(np.float32)       # B * nprobe = 4 -> S = 1
    idx = _lib.Index(d, "bf16", "ip", 0)
    idx.add(X)
    iv = _lib.IVF(idx, 2)
    iv.set_centroids(np.stack([u, -u]))                   # every row lands in list 0
    iv.sync()
    assert iv.list_sizes().tolist() == [N, 0]
    D, I = iv.search_host(Q, k, 1)
    Xr, Qr = flat_ip.round_to(X, "bf16"), flat_ip.round_to(Q, "bf16")
    rD, rI = flat_ip.topk_desc(Qr[:2] @ Xr.T, k)
    sc = lambda b, ids: Xr[ids].astype(np.float64) @ Qr[b].astype(np.float64)   # noqa: E731
    assert not flat_ip.check_topk(D[:2], I[:2], sc, rD, rI, rel=REL)
    assert (I[2:] == -1).all()                            # -u probes the empty list
    if order == "constant":
        assert I[0].tolist() == list(range(k))            # all tied: the k smallest ids, ascending


def test_cosine_metric_scales_by_the_stored_inverse_norms(sim):
    rng = np.random.default_rng(12)
    X = (rng.standard_normal((800, 48)) * rng.uniform(0.2, 5.0, size=(800, 1))).astype(np.float32)
    Q = flat_ip.normalize_rows(rng.standard_normal((3, 48)).astype(np.float32)).astype(np.float32)
    idx = _lib.Index(48, "bf16", "cosine", 0)
    idx.add(X)
    iv = _lib.IVF(idx, 5)
    iv.set_centroids(ivf_train.train_centroids(flat_ip.normalize_rows(X).astype(np.float32), 5))
    D, I = iv.search_host(Q, 20, 5)
    eD, eI = idx.search_host(Q, 20, path="stream")
    assert np.array_equal(I, eI) and np.allclose(D, eD, rtol=1e-5, atol=1e-6)


# --------------------------------------------------------------- drop-in classes ---
def write_faiss_ivf(path, x, centroids, assign, nprobe=10, sparse=False, ids=None):
    """An IndexIVFFlat file in the layout tristage_rag_b200/faiss_io.py documents (produced by this test, not by
    FAISS -- FAISS is not installable here)."""
    import struct

    n, d = x.shape
    nlist = centroids.shape[0]
    hdr = lambda nn: struct.pack("<iqqqBi", d, nn, 1 << 20, 1 << 20, 1, 0)       # noqa: E731
    b = b"IwFl" + hdr(n) + struct.pack("<QQ", nlist, nprobe)
    b += b"IxFI" + hdr(nlist) + struct.pack("<Q", nlist * d) + centroids.astype("<f4").tobytes()
    b += struct.pack("<BQ", 0, 0)
    b += b"ilar" + struct.pack("<QQ", nlist, 4 * d)
    sizes = np.bincount(assign, minlength=nlist).astype("<u8")
    if sparse:
        nz = np.nonzero(sizes)[0]
        b += b"sprs" + struct.pack("<Q", 2 * len(nz)) + np.stack([nz.astype("<u8"), sizes[nz]], axis=1).tobytes()
    else:
        b += b"full" + struct.pack("<Q", nlist) + sizes.tobytes()
    ids = np.arange(n, dtype=np.int64) if ids is None else ids
    for l in range(nlist):
        rows = np.nonzero(assign == l)[0]
        if len(rows):
            b += x[rows].astype("<f4").tobytes() + ids[rows].astype("<i8").tobytes()
    open(path, "wb").write(b)
    return b


def test_faiss_ivf_reader_accepts_only_self_consistent_files(tmp_path):
    from tristage_rag_b200.faiss_io import FaissFormatError, read_faiss_flat, read_faiss_ivf

    X, _ = clustered(300, 16, 4, seed=2)
    cent = ivf_train.train_centroids(X, 5)
    a = oivf.assign_lists(X, cent)
    for sparse in (False, True):
        p = str(tmp_path / f"ok{sparse}")
        blob = write_faiss_ivf(p, X, cent, a, nprobe=3, sparse=sparse)
        got = read_faiss_ivf(p)
        assert (got["vectors"] == X).all() and (got["assign"] == a).all() and (got["centroids"] == cent).all()
        assert got["nlist"] == 5 and got["nprobe"] == 3 and got["metric"] == "ip"
    with pytest.raises(FaissFormatError, match="read_faiss_ivf"):
        read_faiss_flat(p)
    bad_ids = np.arange(300, dtype=np.int64)
    bad_ids[7] = 8                                            # a duplicated id: not insertion positions
    cases = {"short": blob[:-8], "long": blob + b"\0" * 8, "fourcc": b"IwXX" + blob[4:],
             "nlist": blob[:37] + (6).to_bytes(8, "little") + blob[45:]}
    for name, data in cases.items():
        q = str(tmp_path / name)
        open(q, "wb").write(data)
        with pytest.raises(FaissFormatError):
            read_faiss_ivf(q)
    q = str(tmp_path / "dupids")
    write_faiss_ivf(q, X, cent, a, ids=bad_ids)
    with pytest.raises(FaissFormatError, match="insertion positions"):
        read_faiss_ivf(q)


def _retriever(tmp_path, **cfg):
    from oracle.fakes import FakeSentenceEncoder
    from tristage_rag_b200.stage1_retriever import Stage1Config, Stage1Retriever

    config = Stage1Config(cache_dir=str(tmp_path / "models"), index_dir=str(tmp_path / "idx"), enable_bm25=False, **cfg)
    return Stage1Retriever(config, model=FakeSentenceEncoder(32))


def test_retriever_follows_the_reference_index_rule_when_approximate(sim, tmp_path):
    """reference :262-273: more than 1000 rows in the FIRST batch -> IndexIVFFlat(nlist) trained on that batch with
    nprobe from the config; otherwise flat, and the type never changes afterwards (:310-313)."""
    X, centers = clustered(1400, 32, 8, seed=5)
    Q = centers[:3] + 0.01
    r = _retriever(tmp_path, approximate=True, nlist=8, nprobe=2, top_k_candidates=20)
    r.add_embeddings(X[:1100], normalize=False)
    assert r.get_stats()["faiss_index_type"] == "IndexIVFFlat" and r.faiss_index.nprobe == 2
    r.add_embeddings(X[1100:], normalize=False)
    assert r.faiss_index.ntotal == 1400
    got = r.search_batch(np.asarray(Q, np.float32))
    assert r.faiss_index._ivf.nassigned == 1400          # later batches reach their lists with the next search
    assert len(got) == 3 and all(len(g) == 20 for g in got)
    a = r.faiss_index._ivf.assignments()
    qn = r._normalize_embeddings(np.asarray(Q, np.float32))
    lists, _ = r.faiss_index._ivf.coarse_host(qn, 2)
    for b, res in enumerate(got):
        assert all(a[h["doc_id"]] in lists[b] for h in res) and all(h["stage"] == "stage1" for h in res)
        assert [h["score"] for h in res] == sorted((h["score"] for h in res), reverse=True)
        assert isinstance(res[0]["doc_id"], int) and isinstance(res[0]["score"], float)
    # all lists probed == the exact scan of the same rows
    r.faiss_index.nprobe = 8
    full = r.search_batch(np.asarray(Q, np.float32))
    eD, eI = r.faiss_index.exact_search(qn, 20, path="stream")
    for b in range(3):
        ids = [h["doc_id"] for h in full[b]]
        assert set(ids) == set(eI[b].tolist()) or np.allclose(sorted(h["score"] for h in full[b]), np.sort(eD[b]), rtol=1e-5)
    # persistence keeps the lists
    r.faiss_index.nprobe = 2
    r.save_index()
    r2 = _retriever(tmp_path, approximate=True, nlist=8, nprobe=2, top_k_candidates=20)
    r2.load_index()
    assert r2.get_stats()["faiss_index_type"] == "IndexIVFFlat" and r2.faiss_index.nprobe == 2
    again = r2.search_batch(np.asarray(Q, np.float32))
    assert [[h["doc_id"] for h in g] for g in again] == [[h["doc_id"] for h in g] for g in got]
    # a small first batch stays flat whatever comes later; approximate=False is always flat
    small = _retriever(tmp_path / "s", approximate=True, nlist=8, nprobe=2)
    small.add_embeddings(X[:1000], normalize=False)
    small.add_embeddings(X[1000:], normalize=False)
    assert small.get_stats()["faiss_index_type"] == "IndexFlatIP"
    exact = _retriever(tmp_path / "e", nlist=8, nprobe=2)
    exact.add_embeddings(X, normalize=False)
    assert exact.get_stats()["faiss_index_type"] == "IndexFlatIP"
    # saving a flat index where an approximate one was saved before drops the stale lists
    exact.config.index_dir = r.config.index_dir
    exact.save_index()
    r3 = _retriever(tmp_path, approximate=True, nlist=8, nprobe=2)
    r3.load_index()
    assert r3.get_stats()["faiss_index_type"] == "IndexFlatIP"


def test_environment_switch_for_unmodified_callers(monkeypatch):
    from tristage_rag_b200.stage1_retriever import Stage1Config

    assert Stage1Config().approximate is False
    monkeypatch.setenv("TS_APPROXIMATE", "1")
    assert Stage1Config(top_k_candidates=5).approximate is True and Stage1Config(approximate=False).approximate is False


def test_load_index_imports_the_reference_ivf_file(sim, tmp_path):
    import pickle

    X, centers = clustered(600, 32, 5, seed=6)
    cent = ivf_train.train_centroids(X, 5)
    a = oivf.assign_lists(X, cent)
    os_dir = tmp_path / "idx"
    os_dir.mkdir()
    write_faiss_ivf(str(os_dir / "stage1_faiss.index"), X, cent, a, nprobe=2)
    docs = [f"d{i}" for i in range(600)]
    with open(os_dir / "stage1_index.pkl", "wb") as f:
        pickle.dump({"documents": docs, "doc_metadata": [{}] * 600, "config": {}, "bm25_index": None}, f)
    Q = np.asarray(centers[:2], np.float32)
    r = _retriever(tmp_path, approximate=True, top_k_candidates=10)
    r.load_index()
    assert r.get_stats()["faiss_index_type"] == "IndexIVFFlat"
    assert r.faiss_index.nlist == 5 and r.faiss_index.nprobe == 2 and r.faiss_index.ntotal == 600
    assert (r.faiss_index._ivf.assignments() == a).all()                 # the reference's lists, not re-derived ones
    got = r.search_batch(Q)
    Xr, Qr = flat_ip.round_to(X, "bf16"), flat_ip.round_to(flat_ip.normalize_rows(Q).astype(np.float32), "bf16")
    lists, _ = oivf.coarse_probe(flat_ip.normalize_rows(Q), cent, 2)
    rD, rI = oivf.ivf_search(Xr, Qr, a, lists, 10)
    for b in range(2):
        assert [h["doc_id"] for h in got[b]] == rI[b].tolist() and got[b][0]["document"] == docs[rI[b][0]]
    e = _retriever(tmp_path, approximate=False, top_k_candidates=10)       # same file, exact search over its vectors
    e.load_index()
    assert e.get_stats()["faiss_index_type"] == "IndexFlatIP" and e.faiss_index.ntotal == 600


def test_lists_shard_by_rows_like_the_exact_index(sim):
    """SURVEY §8e for the approximate mode: every "rank" holds a contiguous row range behind lists built from the
    SAME centroids; merging the per-shard [B, k] results (ts_topk_merge, what follows the all-gather) gives the
    single-index result, because a global list is the union of the shards' local lists."""
    import ctypes as C

    X, centers = clustered(1800, 32, 7, seed=21)
    Q = np.asarray(centers[:4] + 0.02, np.float32)
    idx, iv, cent = build(X, 7, "bf16")
    k, nprobe, G = 20, 2, 3
    D, I = iv.search_host(Q, k, nprobe)
    S, Id = np.empty((G, 4, k), np.float32), np.empty((G, 4, k), np.int64)
    keep = []
    for r, rows in enumerate(np.array_split(np.arange(1800), G)):
        part = _lib.Index(32, "bf16", "ip", 0)
        part.add(X[rows])
        part.set_id_base(int(rows[0]))
        piv = _lib.IVF(part, 7)
        piv.set_centroids(cent)
        S[r], Id[r] = piv.search_host(Q, k, nprobe)
        assert (piv.assignments() == iv.assignments()[rows]).all()
        keep.append((part, piv))
    out_s, out_i = np.empty((4, k), np.float32), np.empty((4, k), np.int64)
    p = lambda a: C.c_void_p(a.ctypes.data)      # noqa: E731
    _lib.check(sim.ts_topk_merge(0, p(S), p(Id), G, 4, k, p(out_s), p(out_i), None))
    assert np.array_equal(out_i, I) and np.array_equal(out_s, D)


def test_batches_scan_list_by_list_with_identical_results(sim, monkeypatch):
    """Batches: the (query, probe) pairs are sorted by list so the CTAs of one list are neighbours in launch
    order (L2 reuse on the device); the result must not depend on that order, nor on falling back to the
    (segment, probe, query) grid when there are more pairs than the order kernel sorts."""
    X, centers = clustered(1400, 24, 7, seed=31)
    idx, iv, cent = build(X, 7, "bf16")
    rng = np.random.default_rng(1)
    Q = flat_ip.normalize_rows(centers[rng.integers(0, 7, size=40)]
                               + 0.4 * rng.standard_normal((40, 24)).astype(np.float32) / np.sqrt(24)).astype(np.float32)
    D, I = iv.search_host(Q, 10, 3)                       # 120 pairs: ordered by list
    monkeypatch.setenv("TS_IVF_NOORDER", "1")
    D0, I0 = iv.search_host(Q, 10, 3)
    monkeypatch.delenv("TS_IVF_NOORDER")
    assert np.array_equal(I, I0) and np.array_equal(D, D0)
    one = [iv.search_host(Q[b:b + 1], 10, 3) for b in range(5)]          # batch 1 never builds a work list
    assert all(np.array_equal(one[b][1][0], I[b]) and np.array_equal(one[b][0][0], D[b]) for b in range(5))
    monkeypatch.setenv("TS_IVF_MAXPAIRS", "100")          # more pairs than the order kernel takes (4096 in production):
    D1, I1 = iv.search_host(Q, 10, 3)                     # the (segment, probe, query) grid, unordered
    assert np.array_equal(I1, I) and np.array_equal(D1, D)


def test_faiss_readers_never_crash_on_damaged_files(tmp_path):
    """Seeded random truncations, byte flips and length-field edits of valid flat / IVF index files: the readers
    either return a self-consistent result or raise FaissFormatError -- never another exception."""
    import struct

    from tristage_rag_b200.faiss_io import FaissFormatError, read_faiss_flat, read_faiss_ivf

    X, _ = clustered(120, 8, 3, seed=4)
    cent = ivf_train.train_centroids(X, 4)
    a = oivf.assign_lists(X, cent)
    ivf_blob = write_faiss_ivf(str(tmp_path / "v.index"), X, cent, a, nprobe=2)
    flat_blob = b"IxFI" + struct.pack("<iqqqBi", 8, 120, 1 << 20, 1 << 20, 1, 0) + struct.pack("<Q", 960) + X.tobytes()
    rng = np.random.default_rng(0)
    outcomes = {"ok": 0, "refused": 0}
    for blob, reader in ((ivf_blob, read_faiss_ivf), (flat_blob, read_faiss_flat)):
        for trial in range(150):
            b = bytearray(blob)
            kind = trial % 3
            if kind == 0:
                b = b[: int(rng.integers(0, len(b)))]
            elif kind == 1:
                for _ in range(int(rng.integers(1, 4))):
                    b[int(rng.integers(0, min(len(b), 200)))] ^= 1 << int(rng.integers(0, 8))     # header area
            else:
                at = int(rng.integers(0, max(1, min(len(b), 220) - 8)))
                b[at:at + 8] = struct.pack("<Q", [0, 1, 2 ** 31, 2 ** 63 - 1, 2 ** 64 - 1, 121, 7][int(rng.integers(0, 7))])
            p = str(tmp_path / "damaged")
            open(p, "wb").write(bytes(b))
            try:
                reader(p)
                outcomes["ok"] += 1
            except FaissFormatError:
                outcomes["refused"] += 1
    assert outcomes["refused"] > 100 and outcomes["ok"] > 0, outcomes      # flips inside float payloads and ignored fields are legitimate
