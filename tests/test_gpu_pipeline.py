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


@pytest.mark.parametrize("storage", ["fp32", "bf16"])
def test_batched_pipeline_matches_reference_batch_search(cuda_device, golden_dir, storage):
    """BatchedPipeline (one Stage-1 scan + one Stage-2 launch per batch) over the drop-in stages
    against the UNMODIFIED reference RetrievalPipeline.batch_search (pipeline_batch.json)."""
    from tristage_rag_b200.pipeline import BatchedPipeline

    with open(os.path.join(golden_dir, "pipeline_batch.json")) as f:
        g = json.load(f)
    tol = 1e-5 if storage == "fp32" else 4e-3
    for case in g["cases"]:
        docs, cfg = g[case["docs"]], case["config"]
        with tempfile.TemporaryDirectory() as tmp:
            c1 = Stage1Config(device="cpu", cache_dir=os.path.join(tmp, "m"), index_dir=os.path.join(tmp, "i"),
                              top_k_candidates=cfg["stage1_top_k"], batch_size=16,
                              enable_bm25=cfg["stage1_enable_bm25"], bm25_top_k=cfg.get("stage1_bm25_top_k", 300),
                              fusion_method=cfg.get("stage1_fusion_method", "rrf"), use_fp16=False,
                              storage_dtype=storage, gpu_index=cuda_device)
            c2 = Stage2Config(device="cpu", cache_dir=os.path.join(tmp, "m"), max_seq_length=192, batch_size=8,
                              top_k_candidates=cfg["stage2_top_k"], use_fp16=False, storage_dtype=storage,
                              gpu_index=cuda_device)
            r1 = Stage1Retriever(c1, model=fakes.FakeSentenceEncoder(768))
            tok = fakes.FakeTokenizer()
            r2 = ColBERTScorer(c2, tokenizer=tok, model=fakes.FakeTokenModel(tok, 128))
            r1.add_documents(list(docs))
            bp = BatchedPipeline(stage1=r1, stage2=r2, stage3=fakes.FakeReranker(cfg["stage3_top_k"]),
                                 stage1_top_k=cfg["stage1_top_k"], final_top_k=cfg["stage3_top_k"],
                                 save_intermediate_results=cfg["save_intermediate_results"])
            l1 = r1.faiss_index._index.launches
            out = bp.batch_search(g["queries"])
            s1_launches = r1.faiss_index._index.launches - l1
            if storage == "bf16":     # convert + threshold pre-pass + scan + select
                assert s1_launches <= 4, f"Stage 1 must be one batched search, saw {s1_launches} launches"
            assert bp.performance_stats["total_queries"] == case["total_queries"]
            for got, ref in zip(out, case["results"]):
                assert got["query"] == ref["query"] and sorted(got.keys()) == ref["keys"]
                assert sorted(got["timing"].keys()) == ref["timing_keys"]
                if storage == "fp32":
                    assert [x["doc_id"] for x in got["stage1_results"]] == ref["stage1_ids"]
                    assert [x["doc_id"] for x in got["stage2_results"]] == ref["stage2_ids"]
                    assert [x["doc_id"] for x in got["results"]] == [x["doc_id"] for x in ref["results"]]
                elif ref["results"] and ref["results"][0]["stage3_score"] > 0:
                    # bf16 storage: near-ties may swap further down, a clear winner must not
                    assert got["results"][0]["doc_id"] == ref["results"][0]["doc_id"]
                ref_by_id = {x["doc_id"]: x for x in ref["results"]}
                for x in got["results"]:
                    assert x["stage"] == "stage3" and type(x["stage2_score"]) is float
                    if x["doc_id"] in ref_by_id:
                        assert x["stage2_score"] == pytest.approx(ref_by_id[x["doc_id"]]["stage2_score"], rel=tol, abs=tol)
                        assert x["stage3_score"] == pytest.approx(ref_by_id[x["doc_id"]]["stage3_score"])
                json.dumps(got)
            # the batched answer equals the per-query path of the same classes, in fewer launches
            l1 = r1.faiss_index._index.launches
            for got, q in zip(out, g["queries"]):
                seq = r2.rescore_candidates(q, r1.search(q, cfg["stage1_top_k"]))
                seq = fakes.FakeReranker(cfg["stage3_top_k"]).rerank(q, seq)[: cfg["stage3_top_k"]]
                assert [x["doc_id"] for x in got["results"]] == [x["doc_id"] for x in seq]
            assert s1_launches < r1.faiss_index._index.launches - l1


@pytest.mark.parametrize("method", ["maxsim", "colbert"])
def test_compute_similarity_matrix_equals_the_per_pair_scores(cuda_device, method):
    """ColBERTScorer.compute_similarity_matrix (src/stage2_rescorer.py:307-320): the reference returns
    np.array([_maxsim_score / _colbert_score(query, doc) for doc in documents]) without sorting.  Here it is one
    launch over the resident token store; every entry must equal the oracle's score of the SAME encoder outputs,
    the per-pair functions of the class, and what rescore_candidates attaches to the documents."""
    from oracle import flat_ip, maxsim

    docs = [f"document number {i} talks about topic {i % 7} and item {i * 3 % 11} in some detail" for i in range(40)]
    docs += ["", "short", "a much longer text " * 30]
    tok = fakes.FakeTokenizer()
    sc = ColBERTScorer(Stage2Config(device="cpu", top_k_candidates=len(docs), scoring_method=method, gpu_index=cuda_device),
                       tokenizer=tok, model=fakes.FakeTokenModel(tok, 128))
    query = "topic 3 item 5 in detail"
    got = sc.compute_similarity_matrix(query, docs)
    assert isinstance(got, np.ndarray) and got.shape == (len(docs),)
    q = sc.encode_query(query)[0].float().numpy()
    mode = 0 if method == "maxsim" else 1
    nr = lambda x: flat_ip.round_to(maxsim.l2_normalize_tokens(x), "bf16")   # noqa: E731
    ref = np.array([maxsim.score(nr(q), nr(sc.encode_single_document(d)[0].float().numpy()), mode, normalize=False) for d in docs])
    np.testing.assert_allclose(got, ref, rtol=1e-3, atol=2e-4)
    pair = sc._maxsim_score if method == "maxsim" else sc._colbert_score
    for i in (0, 17, len(docs) - 1):
        assert float(pair(sc.encode_query(query), sc.encode_single_document(docs[i]))) == pytest.approx(got[i], rel=1e-5, abs=1e-6)
    res = sc.rescore_candidates(query, [{"doc_id": i, "document": d, "score": 0.0} for i, d in enumerate(docs)])
    by_id = {r["doc_id"]: r["stage2_score"] for r in res}
    assert len(by_id) == len(docs) and all(by_id[i] == pytest.approx(got[i], rel=1e-6, abs=1e-7) for i in range(len(docs)))
    assert [r["doc_id"] for r in res] == sorted(range(len(docs)), key=lambda i: -got[i])     # stable descending
