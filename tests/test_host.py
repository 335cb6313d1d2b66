"""CPU: host-side logic of the drop-in classes (everything around the two GPU
calls), against the golden fixtures generated from the unmodified reference.
The dense index is replaced by the oracle's IndexFlatIP here so the Python
around it can be checked without a GPU; the GPU tests repeat the same cases
through libtristage."""
import inspect
import json
import os
import tempfile

import numpy as np
import pytest

from oracle import fakes, flat_ip
from tristage_rag_b200 import stage1_retriever as s1
from tristage_rag_b200 import stage2_rescorer as s2

REF = "/root/reference"


def _golden(golden_dir):
    with open(os.path.join(golden_dir, "pipeline_c1.json")) as f:
        return json.load(f)


def _retriever(case, docs, tmp):
    cfg = s1.Stage1Config(device="cpu", cache_dir=os.path.join(tmp, "m"), index_dir=os.path.join(tmp, "i"),
                          top_k_candidates=case["s1_topk"], batch_size=16, enable_bm25=case["enable_bm25"],
                          bm25_top_k=case["bm25_topk"], fusion_method=case["fusion"], use_fp16=False)
    r = s1.Stage1Retriever(cfg, model=fakes.FakeSentenceEncoder(768))
    # oracle stand-in for the GPU index (host-logic test only)
    r._create_faiss_index = lambda emb: (setattr(r, "faiss_index", flat_ip.IndexFlatIP(emb.shape[1])),
                                         r.faiss_index.add(emb))
    r.add_documents(list(docs))
    return r


def test_stage1_host_logic_matches_reference_golden(golden_dir):
    g = _golden(golden_dir)
    for case in g["cases"]:
        docs = g[case["docs"]]
        with tempfile.TemporaryDirectory() as tmp:
            r = _retriever(case, docs, tmp)
            stats = r.get_stats()
            for key, val in case["stats"].items():
                assert stats[key] == val, (case["name"], key)
            for qc in case["queries"]:
                res = r.search(qc["query"], case["s1_topk"])
                assert [x["doc_id"] for x in res] == [x["doc_id"] for x in qc["stage1"]], case["name"]
                np.testing.assert_allclose([x["score"] for x in res], [x["score"] for x in qc["stage1"]], rtol=1e-5)
                for x in res:
                    assert set(x) == {"doc_id", "document", "score", "stage1_score", "metadata", "stage"}
                    assert type(x["doc_id"]) is int and type(x["score"]) is float and x["stage"] == "stage1"
                json.dumps(res)                      # MCP server json-dumps stage outputs


def test_stage1_error_and_edge_behaviour():
    with tempfile.TemporaryDirectory() as tmp:
        cfg = s1.Stage1Config(cache_dir=os.path.join(tmp, "m"), index_dir=os.path.join(tmp, "i"), enable_bm25=False)
        r = s1.Stage1Retriever(cfg, model=fakes.FakeSentenceEncoder(32))
        with pytest.raises(ValueError, match=r"No documents indexed\. Call add_documents\(\) first\."):
            r.search("anything")
        r.add_documents([])                          # no-op
        assert r.faiss_index is None and r.documents == []
        assert os.path.isdir(cfg.cache_dir) and os.path.isdir(cfg.index_dir)
        r.load_index(os.path.join(tmp, "missing.pkl"))   # warning + return
        assert r.embedding_dim == 32


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_bm25_matches_reference_including_refit_quirk():
    import sys
    import types

    st = types.ModuleType("sentence_transformers")
    st.SentenceTransformer = object
    st.CrossEncoder = object
    sys.modules.setdefault("sentence_transformers", st)
    sys.modules.setdefault("faiss", fakes.fake_faiss_module())
    sys.path.insert(0, REF)
    from src.stage1_retriever import BM25Index as RefBM25

    docs = ["the cat sat on the mat", "dogs and cats and dogs", "quantum computing is fast", "the the the",
            "", "Cats! cats? CATS."]
    more = ["a new document about cats", "computing with quantum dogs"]
    ours, ref = s1.BM25Index(), RefBM25()
    for batch in (docs, docs + more):               # second fit reproduces the stale-state quirk
        ours.fit(list(batch))
        ref.fit(list(batch))
        assert ours.avg_doc_len == ref.avg_doc_len and ours.idf == ref.idf
        for q in ["cats", "the cat", "quantum dogs dogs", "nothing matches", ""]:
            assert ours.search(q, 4) == ref.search(q, 4)
            assert ours.search(q, 100) == ref.search(q, 100)
            for i in range(len(batch)):
                assert ours.score(q, i) == ref.score(q, i)
    # a larger random corpus: the vectorised CSR search must stay bit-identical to the reference's
    # per-document Python loop (scores as Python floats, stable order, zero-score padding)
    rng = np.random.default_rng(0)
    vocab = [f"w{i}" for i in range(300)]
    zipf = 1.0 / np.arange(1, 301)
    zipf /= zipf.sum()
    big = [" ".join(rng.choice(vocab, size=int(rng.integers(0, 60)), p=zipf)) for _ in range(700)]
    ours, ref = s1.BM25Index(), RefBM25()
    for batch in (big[:500], big):
        ours.fit(list(batch))
        ref.fit(list(batch))
        assert ours.idf == ref.idf
        for q in ["w0 w1 w2", "w7 w7 w250", "w299", "w5 unknownword w5 w0", " ".join(vocab[:40])]:
            for k in (1, 10, 300, 2000):
                assert ours.search(q, k) == ref.search(q, k), (q, k)


def test_drop_in_surface_matches_the_contract():
    """names, signatures and dataclass fields the callers use."""
    for f in ["model_name", "device", "cache_dir", "index_dir", "top_k_candidates", "batch_size", "max_text_length",
              "enable_bm25", "bm25_top_k", "fusion_method", "rrf_k", "dense_weight", "bm25_weight", "use_fp16",
              "nlist", "nprobe"]:
        assert f in s1.Stage1Config.__dataclass_fields__
    c = s1.Stage1Config()
    assert (c.top_k_candidates, c.bm25_top_k, c.rrf_k, c.fusion_method) == (500, 300, 60, "rrf")
    for f in ["model_name", "device", "cache_dir", "max_seq_length", "batch_size", "top_k_candidates", "use_fp16",
              "pooling_method", "normalize_embeddings", "scoring_method", "use_gpu_if_available"]:
        assert f in s2.Stage2Config.__dataclass_fields__
    c2 = s2.Stage2Config()
    assert (c2.max_seq_length, c2.top_k_candidates, c2.scoring_method) == (192, 100, "maxsim")
    sig = inspect.signature(s1.Stage1Retriever.search)
    assert list(sig.parameters) == ["self", "query", "top_k"] and sig.parameters["top_k"].default is None
    sig = inspect.signature(s1.Stage1Retriever.add_documents)
    assert list(sig.parameters) == ["self", "documents", "metadata"]
    for m in ["save_index", "load_index", "get_stats", "_normalize_embeddings", "_encode_batch",
              "_reciprocal_rank_fusion", "_weighted_fusion"]:
        assert callable(getattr(s1.Stage1Retriever, m))
    sig = inspect.signature(s2.ColBERTScorer.rescore_candidates)
    assert list(sig.parameters) == ["self", "query", "candidates"]
    for m in ["encode_query", "encode_documents_batch", "compute_similarity_matrix", "get_model_info",
              "clear_gpu_memory", "_maxsim_score", "_colbert_score", "encode_single_document"]:
        assert callable(getattr(s2.ColBERTScorer, m))


def test_stage2_empty_and_encode_failure_semantics():
    tok = fakes.FakeTokenizer()
    sc = s2.ColBERTScorer(s2.Stage2Config(device="cpu"), tokenizer=tok, model=fakes.FakeTokenModel(tok, 32))
    assert sc.rescore_candidates("q", []) == []
    cands = [{"doc_id": 0, "document": "a b c", "score": 1.0}]

    def boom(_docs):
        raise RuntimeError("encoder down")

    sc.encode_documents_batch = boom
    assert sc.rescore_candidates("q", cands) is cands       # reference :260-263: returned unchanged
    info = sc.get_model_info()
    assert info["embedding_dim"] == 32 and info["scoring_method"] == "maxsim"


def test_faiss_flat_reader_accepts_only_self_consistent_files(tmp_path):
    """tristage_rag_b200/faiss_io.py: the layout is restated from FAISS's public writer (FAISS is not installable
    here, so the file below is produced by this test, not by FAISS); the reader must refuse anything whose
    fields do not add up instead of guessing."""
    import struct

    from tristage_rag_b200.faiss_io import FaissFormatError, read_faiss_flat

    rng = np.random.default_rng(0)
    x = rng.standard_normal((57, 24)).astype(np.float32)

    def blob(fourcc=b"IxFI", d=24, n=57, metric=0, n_floats=None, data=None):
        data = x.tobytes() if data is None else data
        return (fourcc + struct.pack("<iqqqBi", d, n, 1 << 20, 1 << 20, 1, metric) +
                struct.pack("<Q", n * d if n_floats is None else n_floats) + data)

    def write(name, b):
        p = str(tmp_path / name)
        open(p, "wb").write(b)
        return p

    got, metric = read_faiss_flat(write("ok", blob()))
    assert metric == "ip" and got.dtype == np.float32 and (got == x).all()
    assert read_faiss_flat(write("l2", blob(b"IxF2", metric=1)))[1] == "l2"
    for name, b in (("ivf", blob(b"IwFl")), ("other", blob(b"IxPQ")), ("short", blob()[:-4]), ("long", blob() + b"\\0"),
                    ("count", blob(n_floats=57 * 24 - 1)), ("dim", blob(d=25)), ("mix", blob(b"IxFI", metric=1)),
                    ("tiny", b"IxFI")):
        with pytest.raises(FaissFormatError):
            read_faiss_flat(write(name, b))
