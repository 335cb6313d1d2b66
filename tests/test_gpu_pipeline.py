"""GPU: BASELINE config #1 -- test_docs.json / demo docs through the drop-in
Stage1Retriever -> ColBERTScorer, against results recorded from the UNMODIFIED
reference classes (tests/golden/pipeline_c1.json, oracle/gen_golden.py)."""
import json
import os
import tempfile

import numpy as np
import pytest

from oracle import fakes
from tristage_rag_b200 import ColBERTScorer, Stage1Config, Stage1Retriever, Stage2Config

pytestmark = pytest.mark.gpu


def _golden(golden_dir):
    with open(os.path.join(golden_dir, "pipeline_c1.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("storage", ["fp32", "bf16"])
def test_config1_matches_reference(cuda_device, golden_dir, storage):
    g = _golden(golden_dir)
    tol = 1e-5 if storage == "fp32" else 4e-3
    for case in g["cases"]:
        docs = g[case["docs"]]
        with tempfile.TemporaryDirectory() as tmp:
            c1 = Stage1Config(device="cpu", cache_dir=os.path.join(tmp, "m"), index_dir=os.path.join(tmp, "i"),
                              top_k_candidates=case["s1_topk"], batch_size=16, enable_bm25=case["enable_bm25"],
                              bm25_top_k=case["bm25_topk"], fusion_method=case["fusion"], use_fp16=False,
                              storage_dtype=storage, gpu_index=cuda_device)
            c2 = Stage2Config(device="cpu", cache_dir=os.path.join(tmp, "m"), max_seq_length=192, batch_size=8,
                              top_k_candidates=case["s2_topk"], use_fp16=False, scoring_method=case["scoring"],
                              storage_dtype=storage, gpu_index=cuda_device)
            r1 = Stage1Retriever(c1, model=fakes.FakeSentenceEncoder(768))
            tok = fakes.FakeTokenizer()
            r2 = ColBERTScorer(c2, tokenizer=tok, model=fakes.FakeTokenModel(tok, 128))
            r1.add_documents(list(docs))
            stats = r1.get_stats()
            for key, val in case["stats"].items():
                assert stats[key] == val, (case["name"], key)
            for qc in case["queries"]:
                st1 = r1.search(qc["query"], case["s1_topk"])
                if storage == "fp32" or not case["enable_bm25"]:
                    # dense ranks feed RRF: with bf16 storage only near-ties may swap
                    assert [x["doc_id"] for x in st1] == [x["doc_id"] for x in qc["stage1"]], (case["name"], qc["query"])
                    np.testing.assert_allclose([x["score"] for x in st1], [x["score"] for x in qc["stage1"]],
                                               rtol=tol, atol=tol)
                assert {x["doc_id"] for x in st1} == {x["doc_id"] for x in qc["stage1"]}
                json.dumps(st1)
                st2 = r2.rescore_candidates(qc["query"], st1)
                got = {x["doc_id"]: x["stage2_score"] for x in st2}
                ref = {x["doc_id"]: x["stage2_score"] for x in qc["stage2"]}
                assert [x["doc_id"] for x in st2] == [x["doc_id"] for x in qc["stage2"]], (case["name"], qc["query"])
                for k_, v in ref.items():
                    assert got[k_] == pytest.approx(v, rel=tol, abs=tol)
                for x in st2:
                    assert x["stage"] == "stage2" and type(x["stage2_score"]) is float
                json.dumps(st2)
            # second query round hits the resident token store (no re-encoding)
            n_before = r2._store.ndocs
            r2.rescore_candidates(case["queries"][0]["query"], r1.search(case["queries"][0]["query"]))
            assert r2._store.ndocs == n_before


def test_save_load_and_batch_search(cuda_device, golden_dir):
    g = _golden(golden_dir)
    docs = g["demo_docs"]
    with tempfile.TemporaryDirectory() as tmp:
        cfg = Stage1Config(cache_dir=os.path.join(tmp, "m"), index_dir=os.path.join(tmp, "i"), enable_bm25=True,
                           top_k_candidates=5, gpu_index=cuda_device)
        r = Stage1Retriever(cfg, model=fakes.FakeSentenceEncoder(768))
        r.add_documents(docs[:4])
        r.add_documents(docs[4:], metadata=[{"i": i} for i in range(4, len(docs))])
        a = r.search("What is machine learning?")
        batch = r.search_batch(g["demo_queries"], 5)
        assert [x["doc_id"] for x in batch[0]] == [x["doc_id"] for x in a]
        pkl = os.path.join(tmp, "pipeline_index.pkl")
        r.save_index(pkl)
        assert os.path.exists(os.path.join(cfg.index_dir, "stage1_faiss.index"))
        r2 = Stage1Retriever(cfg, model=fakes.FakeSentenceEncoder(768))
        r2.load_index(pkl)
        b = r2.search("What is machine learning?")
        assert [(x["doc_id"], x["score"]) for x in b] == [(x["doc_id"], x["score"]) for x in a]
        assert r2.doc_metadata[5] == {"i": 5}
