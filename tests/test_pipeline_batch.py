"""CPU: the orchestration of tristage_rag_b200/pipeline.py::BatchedPipeline against what the
UNMODIFIED reference RetrievalPipeline.batch_search returned (tests/golden/pipeline_batch.json,
oracle/gen_golden.py).  Stage 1/2 are replayed from the golden here (no GPU); the GPU test
(tests/test_gpu_pipeline.py) runs the same comparison over the real drop-in stages."""
import json
import os
import types

import pytest

from oracle import fakes
from tristage_rag_b200.pipeline import BatchedPipeline


def _golden(golden_dir):
    with open(os.path.join(golden_dir, "pipeline_batch.json")) as f:
        return json.load(f)


class ReplayStage1:
    """search_batch answering from the recorded Stage-1 id lists; counts its calls."""

    def __init__(self, docs, by_query):
        self.docs, self.by_query, self.calls = docs, by_query, 0

    def search_batch(self, queries, top_k=None):
        self.calls += 1
        return [[{"doc_id": i, "document": self.docs[i], "score": 1.0 / (1 + r), "stage1_score": 1.0 / (1 + r),
                  "metadata": {}, "stage": "stage1"} for r, i in enumerate(self.by_query[q][:top_k])] for q in queries]


class ReplayStage2:
    def __init__(self, by_query):
        self.by_query, self.calls = by_query, 0

    def rescore_candidates_batch(self, queries, candidates):
        self.calls += 1
        assert all(candidates), "queries without candidates must not reach Stage 2"
        out = []
        for q, cands in zip(queries, candidates):
            have = {c["doc_id"]: c for c in cands}
            rows = []
            for r, i in enumerate(self.by_query[q]):
                u = have[i].copy()
                u["stage2_score"], u["stage"] = 1.0 - 0.01 * r, "stage2"
                rows.append(u)
            out.append(rows)
        return out


def test_batched_pipeline_returns_what_the_reference_loop_returns(golden_dir):
    g = _golden(golden_dir)
    case = next(c for c in g["cases"] if c["name"] == "demo_rrf_keep3")
    docs, queries, cfg = g[case["docs"]], g["queries"], case["config"]
    s1 = ReplayStage1(docs, {r["query"]: r["stage1_ids"] for r in case["results"]})
    s2 = ReplayStage2({r["query"]: r["stage2_ids"] for r in case["results"]})
    stats = {"total_queries": 0, "avg_stage1_time": 0.0, "avg_stage2_time": 0.0, "avg_stage3_time": 0.0,
             "avg_total_time": 0.0, "stage_time_history": []}
    ref_pipe = types.SimpleNamespace(                       # the attributes of the reference orchestrator
        stage1=s1, stage2=s2, stage3=fakes.FakeReranker(cfg["stage3_top_k"]), performance_stats=stats,
        config=types.SimpleNamespace(stage1_top_k=cfg["stage1_top_k"], stage3_top_k=cfg["stage3_top_k"],
                                     save_intermediate_results=cfg["save_intermediate_results"], enable_timing=True,
                                     auto_cleanup=False))
    out = BatchedPipeline(ref_pipe).batch_search(queries)
    assert (s1.calls, s2.calls) == (1, 1), "one Stage-1 and one Stage-2 call for the whole batch"
    assert len(out) == len(case["results"]) and stats["total_queries"] == case["total_queries"]
    assert len(stats["stage_time_history"]) == len(queries)
    for got, ref in zip(out, case["results"]):
        assert got["query"] == ref["query"] and sorted(got.keys()) == ref["keys"]
        assert sorted(got["timing"].keys()) == ref["timing_keys"]
        assert [x["doc_id"] for x in got["stage1_results"]] == ref["stage1_ids"]
        assert [x["doc_id"] for x in got["stage2_results"]] == ref["stage2_ids"]
        assert [x["doc_id"] for x in got["results"]] == [x["doc_id"] for x in ref["results"]]
        assert [x["stage3_score"] for x in got["results"]] == pytest.approx([x["stage3_score"] for x in ref["results"]])
        assert all(x["stage"] == "stage3" for x in got["results"])
        json.dumps(got)


def test_empty_stages_short_circuit_like_the_reference():
    docs = ["a b", "c d"]
    s1 = ReplayStage1(docs, {"hit": [0, 1], "miss": [], "lost": [1]})
    s2 = ReplayStage2({"hit": [1, 0], "lost": []})
    bp = BatchedPipeline(stage1=s1, stage2=s2, stage3=None, stage1_top_k=5, final_top_k=1,
                         save_intermediate_results=True, enable_timing=False)
    assert bp.batch_search([]) == []
    out = bp.batch_search(["hit", "miss", "lost"])
    assert [x["doc_id"] for x in out[0]["results"]] == [1] and out[0]["timing"] == {}
    assert out[1]["results"] == [] and out[1]["stage1_results"] == [] and out[1]["stage2_results"] == []
    assert out[2]["results"] == [] and [x["doc_id"] for x in out[2]["stage1_results"]] == [1] and out[2]["stage2_results"] == []
    assert bp.search("hit")["results"][0]["stage2_score"] == 1.0
    assert bp.performance_stats["total_queries"] == 0            # timing disabled: no stats, like the reference
    with pytest.raises(TypeError):
        BatchedPipeline(stage1=object(), stage2=s2)
    with pytest.raises(ValueError):
        BatchedPipeline(types.SimpleNamespace(stage1=None, stage2=None))
