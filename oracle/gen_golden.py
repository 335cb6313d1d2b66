#!/usr/bin/env python
"""Generate tests/golden/*.json from the UNMODIFIED reference code.

Run in the authoring container only (needs /root/reference):

    python oracle/gen_golden.py

What is imported from the reference (read-only, never copied):
  * ``src.stage2_rescorer.ColBERTScorer._maxsim_score`` / ``_colbert_score``
    (stage2_rescorer.py:167-201) -- called as unbound functions on seeded
    tensors -> ``stage2_reference.json``.
  * ``src.stage1_retriever.Stage1Retriever`` / ``BM25Index`` and
    ``src.stage2_rescorer.ColBERTScorer`` as whole classes, constructed with
    the fake encoders of ``oracle/fakes.py`` and the restated ``faiss`` module
    (FAISS is not installable here) on BASELINE config #1 inputs
    (non_mcp/test_docs.json, non_mcp/pipeline_config.yaml values, the
    mcp/demo.py documents and queries) -> ``pipeline_c1.json``.
  * ``Stage1Retriever._normalize_embeddings`` (stage1_retriever.py:285-288)
    on a seeded matrix -> ``stage1_normalize.json``.
  * ``src.retrieval_pipeline.RetrievalPipeline.batch_search`` (:426-448) over
    those classes with ``FakeReranker`` as Stage 3 -> ``pipeline_batch.json``.

Unused third-party imports of the reference (``sentence_transformers``) are
stubbed in ``sys.modules``; nothing in the reference is modified.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"

from oracle import fakes, flat_ip  # noqa: E402

DEMO_DOCS = None
DEMO_QUERIES = None


def import_reference():
    st = types.ModuleType("sentence_transformers")
    st.SentenceTransformer = object
    st.CrossEncoder = object
    sys.modules["sentence_transformers"] = st
    sys.modules["faiss"] = fakes.fake_faiss_module()
    sys.path.insert(0, REF)
    import src.stage1_retriever as s1
    import src.stage2_rescorer as s2

    return s1, s2


def demo_fixture_inputs():
    """Borrowed INPUTS (not code): the 5 test docs and the demo docs/queries."""
    with open(os.path.join(REF, "non_mcp", "test_docs.json")) as f:
        test_docs = json.load(f)
    # mcp/demo.py:21-32 and :45-49 are string literals; read them by executing
    # nothing -- parse the literals out of the source text.
    import ast

    src = open(os.path.join(REF, "mcp", "demo.py")).read()
    tree = ast.parse(src)
    lists = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and isinstance(node.value, ast.List):
            name = getattr(node.targets[0], "id", None)
            if name in ("documents", "queries"):
                lists[name] = [ast.literal_eval(e) for e in node.value.elts]
    return test_docs, lists["documents"], lists["queries"]


def gen_stage2(s2):
    cases = []
    # the SURVEY.md Appendix-B probe (torch CPU generator, seed 0)
    torch.manual_seed(0)
    q = torch.randn(1, 32, 128)
    d = torch.randn(1, 180, 128)
    cases.append(dict(kind="torch_seed0", Lq=32, Ld=180, H=128,
                      maxsim=float(s2.ColBERTScorer._maxsim_score(None, q, d)),
                      colbert=float(s2.ColBERTScorer._colbert_score(None, q, d))))
    # numpy-seeded cases (PCG64 standard_normal is stable across platforms)
    shapes = [(32, 180, 128), (32, 16, 128), (32, 98, 128), (7, 33, 64), (1, 2, 8),
              (32, 192, 128), (5, 2, 128), (64, 100, 96), (32, 1, 128), (3, 17, 32)]
    for i, (Lq, Ld, H) in enumerate(shapes):
        rng = np.random.default_rng(1000 + i)
        qn = rng.standard_normal((Lq, H)).astype(np.float32) * np.float32(1 + i)
        dn = rng.standard_normal((Ld, H)).astype(np.float32) * np.float32(0.5 + i)
        qt = torch.from_numpy(qn).unsqueeze(0)          # [1, Lq, H]  (stage2_rescorer.py:165)
        dt = torch.from_numpy(dn)                       # [Ld, H]     (stage2_rescorer.py:230)
        if Ld == 1:
            # squeeze(0) would collapse a 1-token doc -> give [1,1,H]
            dt = dt.unsqueeze(0)
        cases.append(dict(kind="numpy", seed=1000 + i, Lq=Lq, Ld=Ld, H=H,
                          q_scale=float(1 + i), d_scale=float(0.5 + i),
                          maxsim=float(s2.ColBERTScorer._maxsim_score(None, qt, dt)),
                          colbert=float(s2.ColBERTScorer._colbert_score(None, qt, dt))))
    with open(os.path.join(GOLD, "stage2_reference.json"), "w") as f:
        json.dump(dict(source="reference src/stage2_rescorer.py:167-201, torch %s" % torch.__version__,
                       cases=cases), f, indent=1)
    print("stage2_reference.json:", len(cases), "cases")


def gen_normalize(s1):
    rng = np.random.default_rng(7)
    x = (rng.standard_normal((6, 12)) * 3).astype(np.float32)
    x[4] = 0.0                                           # zero row: eps keeps it finite
    y = s1.Stage1Retriever._normalize_embeddings(None, x)
    with open(os.path.join(GOLD, "stage1_normalize.json"), "w") as f:
        json.dump(dict(source="reference src/stage1_retriever.py:285-288", seed=7, shape=[6, 12],
                       scale=3.0, zero_row=4, dtype=str(y.dtype),
                       y=[[float(v) for v in r] for r in y]), f, indent=1)
    print("stage1_normalize.json")


def run_pipeline_case(s1, s2, docs, queries, enable_bm25, fusion, s1_topk, bm25_topk, s2_topk, scoring):
    """Unmodified reference Stage1Retriever.search -> ColBERTScorer.rescore_candidates."""
    enc = fakes.FakeSentenceEncoder(768)
    tok = fakes.FakeTokenizer()
    tokmodel = fakes.FakeTokenModel(tok, 128)

    def _load_s1(self):
        self.model = enc
        self.embedding_dim = enc.get_sentence_embedding_dimension()

    def _load_s2(self):
        self.tokenizer, self.model, self.use_amp = tok, tokmodel, False

    # encoders are OUT of the hot path (BASELINE.json north_star); only the
    # weight-loading hooks are replaced, the scoring code is the reference's
    s1.Stage1Retriever._load_model = _load_s1
    s2.ColBERTScorer._load_model = _load_s2
    tmp = tempfile.mkdtemp()
    c1 = s1.Stage1Config(device="cpu", cache_dir=os.path.join(tmp, "m"), index_dir=os.path.join(tmp, "i"),
                         top_k_candidates=s1_topk, batch_size=16, enable_bm25=enable_bm25,
                         bm25_top_k=bm25_topk, fusion_method=fusion, use_fp16=False)
    c2 = s2.Stage2Config(device="cpu", cache_dir=os.path.join(tmp, "m"), max_seq_length=192, batch_size=8,
                         top_k_candidates=s2_topk, use_fp16=False, scoring_method=scoring)
    r1 = s1.Stage1Retriever(c1)
    r2 = s2.ColBERTScorer(c2)
    r1.add_documents(list(docs))
    out = []
    for q in queries:
        st1 = r1.search(q, s1_topk)
        st2 = r2.rescore_candidates(q, st1)
        out.append(dict(query=q,
                        stage1=[dict(doc_id=r["doc_id"], score=r["score"], stage1_score=r["stage1_score"]) for r in st1],
                        stage2=[dict(doc_id=r["doc_id"], stage2_score=r["stage2_score"], score=r["score"]) for r in st2]))
    stats = r1.get_stats()
    stats.pop("config")
    return dict(enable_bm25=enable_bm25, fusion=fusion, s1_topk=s1_topk, bm25_topk=bm25_topk,
                s2_topk=s2_topk, scoring=scoring, n_docs=len(docs), stats=stats, queries=out)


def gen_pipeline(s1, s2):
    test_docs, demo_docs, demo_queries = demo_fixture_inputs()
    cases = []
    # BASELINE config #1: pipeline_config.yaml values (S1 top_k 50, bm25 on / rrf,
    # bm25_top_k 100, S2 top_k 20, maxsim) on test_docs.json + the demo queries
    cases.append(dict(name="c1_test_docs_rrf", docs="test_docs",
                      **run_pipeline_case(s1, s2, test_docs, demo_queries, True, "rrf", 50, 100, 20, "maxsim")))
    cases.append(dict(name="c1_test_docs_dense_only", docs="test_docs",
                      **run_pipeline_case(s1, s2, test_docs, demo_queries, False, "rrf", 50, 100, 20, "maxsim")))
    cases.append(dict(name="demo_docs_weighted_colbert", docs="demo_docs",
                      **run_pipeline_case(s1, s2, demo_docs, demo_queries, True, "weighted", 8, 5, 3, "colbert")))
    cases.append(dict(name="demo_docs_dense_only_k3", docs="demo_docs",
                      **run_pipeline_case(s1, s2, demo_docs, demo_queries, False, "rrf", 3, 5, 2, "maxsim")))
    with open(os.path.join(GOLD, "pipeline_c1.json"), "w") as f:
        json.dump(dict(source="unmodified reference Stage1Retriever.search + ColBERTScorer.rescore_candidates "
                              "with oracle/fakes.py encoders and the restated faiss module",
                       test_docs=test_docs, demo_docs=demo_docs, demo_queries=demo_queries, cases=cases),
                  f, indent=1)
    print("pipeline_c1.json:", len(cases), "cases")


def gen_pipeline_batch(s1, s2):
    """Unmodified reference RetrievalPipeline.batch_search (src/retrieval_pipeline.py:426-448, the
    sequential loop over :323-424) with the reference Stage-1/Stage-2 classes, the fake encoders
    and FakeReranker as Stage 3 -> pipeline_batch.json.  Pins what tristage_rag_b200/pipeline.py's
    BatchedPipeline must return."""
    import src.retrieval_pipeline as rp

    test_docs, demo_docs, demo_queries = demo_fixture_inputs()
    enc = fakes.FakeSentenceEncoder(768)
    tok = fakes.FakeTokenizer()
    tokmodel = fakes.FakeTokenModel(tok, 128)

    def _load_s1(self):
        self.model = enc
        self.embedding_dim = enc.get_sentence_embedding_dimension()

    def _load_s2(self):
        self.tokenizer, self.model, self.use_amp = tok, tokmodel, False

    s1.Stage1Retriever._load_model = _load_s1
    s2.ColBERTScorer._load_model = _load_s2
    cases = []
    queries = list(demo_queries) + ["", "zzzz qqqq unknown words only", demo_queries[0]]
    for name, docs, kw in (
            ("demo_rrf_keep3", demo_docs, dict(stage1_top_k=8, stage1_enable_bm25=True, stage1_bm25_top_k=5,
                                               stage1_fusion_method="rrf", stage2_top_k=5, stage3_top_k=3,
                                               save_intermediate_results=True)),
            ("test_docs_dense_keep2", test_docs, dict(stage1_top_k=50, stage1_enable_bm25=False, stage2_top_k=20,
                                                      stage3_top_k=2, save_intermediate_results=False))):
        tmp = tempfile.mkdtemp()
        cfg = rp.PipelineConfig(device="cpu", cache_dir=os.path.join(tmp, "m"), index_dir=os.path.join(tmp, "i"),
                                log_file=os.path.join(tmp, "log.txt"), stage1_use_fp16=False, stage2_use_fp16=False,
                                auto_cleanup=False, **kw)
        pipe = rp.RetrievalPipeline(config=cfg)
        pipe.stage1 = s1.Stage1Retriever(s1.Stage1Config(
            device="cpu", cache_dir=cfg.cache_dir, index_dir=cfg.index_dir, top_k_candidates=cfg.stage1_top_k,
            batch_size=16, enable_bm25=cfg.stage1_enable_bm25, bm25_top_k=cfg.stage1_bm25_top_k,
            fusion_method=cfg.stage1_fusion_method, use_fp16=False))
        pipe.stage2 = s2.ColBERTScorer(s2.Stage2Config(
            device="cpu", cache_dir=cfg.cache_dir, max_seq_length=192, batch_size=8,
            top_k_candidates=cfg.stage2_top_k, use_fp16=False, scoring_method=cfg.stage2_scoring_method))
        pipe.stage3 = fakes.FakeReranker(top_k_final=cfg.stage3_top_k)
        pipe.add_documents(list(docs))
        res = pipe.batch_search(queries)
        slim = lambda rows, keys: [{k: r[k] for k in keys if k in r} for r in rows]   # noqa: E731
        cases.append(dict(
            name=name, docs="demo_docs" if docs is demo_docs else "test_docs", config=kw,
            total_queries=pipe.performance_stats["total_queries"],
            results=[dict(query=r["query"], keys=sorted(r.keys()), timing_keys=sorted(r["timing"].keys()),
                          results=slim(r["results"], ("doc_id", "score", "stage1_score", "stage2_score", "stage3_score", "stage")),
                          stage1_ids=[x["doc_id"] for x in r["stage1_results"]],
                          stage2_ids=[x["doc_id"] for x in r["stage2_results"]]) for r in res]))
    with open(os.path.join(GOLD, "pipeline_batch.json"), "w") as f:
        json.dump(dict(source="unmodified reference RetrievalPipeline.batch_search with reference Stage-1/2 classes, "
                              "oracle/fakes.py encoders + FakeReranker, restated faiss module",
                       test_docs=test_docs, demo_docs=demo_docs, queries=queries, cases=cases), f, indent=1)
    print("pipeline_batch.json:", len(cases), "cases")


def gen_flat_ip():
    """Oracle-generated regression vectors for the restated IndexFlatIP (FAISS is
    absent: these pin the oracle against itself and document the semantics)."""
    rng = np.random.default_rng(42)
    x = flat_ip.normalize_rows(rng.standard_normal((40, 16)).astype(np.float32)).astype(np.float32)
    q = flat_ip.normalize_rows(rng.standard_normal((3, 16)).astype(np.float32)).astype(np.float32)
    x[7] = x[3]                                          # exact duplicate -> tie broken by id asc
    idx = flat_ip.IndexFlatIP(16)
    idx.add(x)
    D, I = idx.search(q, 5)
    D2, I2 = idx.search(q, 50)                           # k > ntotal -> -1 padding
    with open(os.path.join(GOLD, "stage1_flat_ip.json"), "w") as f:
        json.dump(dict(source="oracle/flat_ip.py (restated faiss.IndexFlatIP; FAISS not installable)",
                       seed=42, n=40, d=16, nq=3, dup=[7, 3],
                       k5=dict(D=D.tolist(), I=I.tolist()),
                       k50_valid=int((I2[0] >= 0).sum()), k50_I0=I2[0].tolist(),
                       k50_pad_score=float(D2[0, -1])), f, indent=1)
    print("stage1_flat_ip.json")


def gen_ivf():
    """Oracle-generated regression vectors for the restated IndexIVFFlat (FAISS is absent: these pin
    oracle/ivf.py against itself and document the semantics -- assignment, probe order, -1 padding)."""
    from oracle import ivf as oivf

    rng = np.random.default_rng(7)
    centers = flat_ip.normalize_rows(rng.standard_normal((4, 12)).astype(np.float32))
    x = flat_ip.normalize_rows(centers[rng.integers(0, 4, size=60)] + 0.25 * rng.standard_normal((60, 12)).astype(np.float32))
    x = x.astype(np.float32)
    q = flat_ip.normalize_rows(centers[:2] + 0.05).astype(np.float32)
    cent = oivf.kmeans_ip(x, 5, niter=10, seed=1234)
    assign = oivf.assign_lists(x, cent)
    lists, lscores = oivf.coarse_probe(q, cent, 2)
    D, I = oivf.ivf_search(x, q, assign, lists, 8)
    Dp, Ip = oivf.ivf_search(x, q, assign, lists[:, :1], 40)          # one list, k larger than the list
    with open(os.path.join(GOLD, "stage1_ivf.json"), "w") as f:
        json.dump(dict(source="oracle/ivf.py (restated faiss.IndexIVFFlat over IndexFlatIP; FAISS not installable)",
                       seed=7, n=60, d=12, nlist=5, nprobe=2, x=x.tolist(), q=q.tolist(),
                       centroids=cent.tolist(), assign=assign.tolist(), lists=lists.tolist(),
                       lscores=lscores.tolist(), k8=dict(D=D.tolist(), I=I.tolist()),
                       one_list_k40_valid=[int((Ip[b] >= 0).sum()) for b in range(2)], one_list_k40_I0=Ip[0].tolist()),
                  f, indent=1)
    print("stage1_ivf.json")


def main():
    os.makedirs(GOLD, exist_ok=True)
    s1, s2 = import_reference()
    gen_stage2(s2)
    gen_normalize(s1)
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    try:
        gen_pipeline(s1, s2)
        gen_pipeline_batch(s1, s2)
    finally:
        os.chdir(cwd)
    gen_flat_ip()
    gen_ivf()


if __name__ == "__main__":
    main()
