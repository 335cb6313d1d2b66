"""GPU: device-side BM25 search and rank fusion (ts_bm25_*, ts_hybrid_fuse_host) against the host
path, which tests/test_host.py pins bit for bit to the reference's BM25Index / fusion code.  The same
cases run on the SIMT emulator in tests/test_cudasim.py."""
import types

import numpy as np
import pytest

from test_cudasim import _zipf_corpus

from oracle import fakes, flat_ip
from tristage_rag_b200 import _lib
from tristage_rag_b200 import stage1_retriever as s1

pytestmark = pytest.mark.gpu


def test_device_bm25_equals_host_search(cuda_device):
    docs, vocab = _zipf_corpus(20000, vocab_size=2000, seed=1)
    bm = s1.BM25Index()
    bm.fit(docs)
    dev = s1.DeviceBM25(bm, cuda_device)
    queries = ["w0 w1 w2", "w7 w7 w250", "w1999", "w5 unknownword w5 w0", " ".join(vocab[:40]), "", "nothing known",
               "w0 " * 12] + [" ".join(vocab[i: i + 3]) for i in range(0, 60, 3)]
    for top_k in (1, 10, 300, 1024):
        got = dev.search_batch(queries, top_k)
        for q, g in zip(queries, got):
            assert g == bm.search(q, top_k), (q, top_k)
    assert dev._dev.launches == 8
    tiny = s1.BM25Index()
    tiny.fit(["a b", "b c", "zz"])
    assert s1.DeviceBM25(tiny, cuda_device).search_batch(["b", "q"], 10) == [tiny.search("b", 10), tiny.search("q", 10)]


@pytest.mark.parametrize("method", ["rrf", "weighted"])
def test_device_fusion_equals_reference_arithmetic(cuda_device, method):
    rng = np.random.default_rng(3)
    host = types.SimpleNamespace(config=types.SimpleNamespace(rrf_k=60, dense_weight=0.7, bm25_weight=0.3))
    fuse = s1.Stage1Retriever._reciprocal_rank_fusion if method == "rrf" else s1.Stage1Retriever._weighted_fusion
    B, k1, k2, top_k = 16, 500, 300, 500
    dense_ids = np.full((B, k1), -1, np.int64)
    dense_sc = np.full((B, k1), flat_ip.LOWEST_F32, np.float32)
    bm_ids = np.full((B, k2), -1, np.int64)
    bm_sc = np.zeros((B, k2), np.float64)
    want = []
    for b in range(B):
        nd = int(rng.integers(1, k1 + 1)) if b else k1
        nb = int(rng.integers(1, k2 + 1)) if b else k2
        universe = 700 if b % 2 == 0 else 100000
        d_ids = rng.choice(universe, size=nd, replace=False)
        d_sc = np.sort(rng.random(nd).astype(np.float32))[::-1] + np.float32(0.01)
        m_ids = rng.choice(universe, size=nb, replace=False)
        m_sc = np.sort(rng.random(nb) * 9)[::-1] + 0.5
        dense_ids[b, :nd], dense_sc[b, :nd] = d_ids, d_sc
        bm_ids[b, :nb], bm_sc[b, :nb] = m_ids, m_sc
        want.append(fuse(host, [(int(i), float(s)) for i, s in zip(d_ids, d_sc)],
                         [(int(i), float(s)) for i, s in zip(m_ids, m_sc)])[:top_k])
    ids, scores, n = _lib.hybrid_fuse(method, 60, 0.7, 0.3, dense_ids, dense_sc, bm_ids, bm_sc, top_k, cuda_device)
    for b in range(B):
        assert [(int(ids[b, r]), float(scores[b, r])) for r in range(int(n[b]))] == want[b], b


def test_search_batch_hybrid_on_device_equals_host_path(cuda_device, tmp_path):
    docs, _ = _zipf_corpus(3000, vocab_size=400, seed=5)
    queries = ["w0 w1", "w3 w3 w40", "", "w399 w2 w7", "unknown"] + [f"w{i} w{2 * i}" for i in range(1, 28)]
    out = {}
    for on_device in (False, True):
        for fusion in ("rrf", "weighted"):
            cfg = s1.Stage1Config(device="cpu", cache_dir=str(tmp_path / "m"), index_dir=str(tmp_path / "i"),
                                  top_k_candidates=100, enable_bm25=True, bm25_top_k=300, fusion_method=fusion,
                                  hybrid_on_device=on_device, storage_dtype="fp32", gpu_index=cuda_device)
            r = s1.Stage1Retriever(cfg, model=fakes.FakeSentenceEncoder(96))
            r.add_documents(docs)
            qs = queries if fusion == "rrf" else [q for q in queries if q not in ("", "unknown")]
            out[(on_device, fusion)] = r.search_batch(qs, 100)
            assert (r._device_bm25 is not None) == on_device
    for fusion in ("rrf", "weighted"):
        assert out[(True, fusion)] == out[(False, fusion)]
