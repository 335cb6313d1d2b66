"""Batched orchestration of Stage 1 -> Stage 2 (-> Stage 3): the ``batch_search`` the reference
runs as a sequential loop (``src/retrieval_pipeline.py:426-448``, one ``search`` per query, so
Stage 1 and Stage 2 only ever see batch 1), re-shaped so that a batch of queries costs ONE
Stage-1 scan launch and ONE Stage-2 scoring launch.

``BatchedPipeline`` wraps any object with the reference orchestrator's attributes -- the
reference's own ``RetrievalPipeline`` after ``initialize_stages()`` (``.stage1``, ``.stage2``,
``.stage3``, ``.config.stage1_top_k / stage3_top_k / save_intermediate_results /
enable_timing``, ``.performance_stats``) or the stages passed directly -- and returns, per
query, the dict ``RetrievalPipeline.search`` builds (``:323-424``): ``query``, ``results``,
``stage1_results``, ``stage2_results``, ``timing``, ``performance_stats``.  Results are
identical to the sequential loop; what differs is the launch count.

Only orchestration lives here: Stage 1 must offer ``search_batch`` and Stage 2
``rescore_candidates_batch`` (the drop-in classes of this package do); Stage 3 (the
cross-encoder, outside the hot path) is called per query through its own ``rerank``.
"""
from __future__ import annotations

import time
from typing import Any, Dict, List, Optional


class BatchedPipeline:
    def __init__(self, pipeline=None, stage1=None, stage2=None, stage3=None, stage1_top_k: Optional[int] = None,
                 final_top_k: Optional[int] = None, save_intermediate_results: Optional[bool] = None,
                 enable_timing: Optional[bool] = None):
        cfg = getattr(pipeline, "config", None)
        self.pipeline = pipeline
        self.stage1 = stage1 if stage1 is not None else getattr(pipeline, "stage1", None)
        self.stage2 = stage2 if stage2 is not None else getattr(pipeline, "stage2", None)
        self.stage3 = stage3 if stage3 is not None else getattr(pipeline, "stage3", None)
        if self.stage1 is None or self.stage2 is None:
            raise ValueError("BatchedPipeline needs stage1 and stage2 (call initialize_stages() on the pipeline first)")
        for obj, meth in ((self.stage1, "search_batch"), (self.stage2, "rescore_candidates_batch")):
            if not hasattr(obj, meth):
                raise TypeError(f"{type(obj).__name__} has no {meth}(): use the tristage_rag_b200 drop-in classes")

        def pick(explicit, name, default):
            return explicit if explicit is not None else getattr(cfg, name, default)

        self.stage1_top_k = pick(stage1_top_k, "stage1_top_k", 500)
        self.final_top_k = pick(final_top_k, "stage3_top_k", 20)
        self.save_intermediate_results = pick(save_intermediate_results, "save_intermediate_results", False)
        self.enable_timing = pick(enable_timing, "enable_timing", True)
        own = {"total_queries": 0, "avg_stage1_time": 0.0, "avg_stage2_time": 0.0, "avg_stage3_time": 0.0,
               "avg_total_time": 0.0, "stage_time_history": []}
        self.performance_stats = getattr(pipeline, "performance_stats", own)

    # -- bookkeeping with the reference's arithmetic (src/retrieval_pipeline.py:542-606) ------
    def _timing(self, total, s1, s2, s3) -> Dict[str, float]:
        if not self.enable_timing:
            return {}
        return {"stage1_time": s1 or 0.0, "stage2_time": s2 or 0.0, "stage3_time": s3 or 0.0,
                "total_time": total or 0.0}

    def _update_stats(self, s1, s2, s3, total) -> None:
        st = self.performance_stats
        st["total_queries"] += 1
        a = 1.0 / st["total_queries"]
        for key, v in (("avg_stage1_time", s1), ("avg_stage2_time", s2), ("avg_stage3_time", s3),
                       ("avg_total_time", total)):
            st[key] = (1 - a) * st[key] + a * v
        st["stage_time_history"].append({"stage1": s1, "stage2": s2, "stage3": s3, "total": total})
        if len(st["stage_time_history"]) > 100:
            st["stage_time_history"] = st["stage_time_history"][-100:]

    # -- the call ----------------------------------------------------------------------------
    def batch_search(self, queries: List[str], top_k: Optional[int] = None) -> List[Dict[str, Any]]:
        """``[RetrievalPipeline.search(q, top_k) for q in queries]`` with one launch per stage.

        Per-query stage times are the batch's wall time divided by the batch size (the stages
        no longer run per query); Stage 3 is timed per query like the reference does."""
        queries = list(queries)
        if not queries:
            return []
        top_k = top_k or self.final_top_k
        n = len(queries)
        now = time.time if self.enable_timing else (lambda: None)
        t_all = now()

        t0 = now()
        stage1 = self.stage1.search_batch(queries, self.stage1_top_k)              # ONE Stage-1 scan
        s1_each = (time.time() - t0) / n if t0 else None

        t0 = now()
        live = [b for b in range(n) if stage1[b]]
        rescored = self.stage2.rescore_candidates_batch([queries[b] for b in live],
                                                        [stage1[b] for b in live]) if live else []
        stage2: List[List[Dict[str, Any]]] = [[] for _ in range(n)]
        for b, r in zip(live, rescored):                                          # ONE Stage-2 launch
            stage2[b] = r
        s2_each = (time.time() - t0) / n if t0 else None

        out = []
        for b, q in enumerate(queries):
            if not stage1[b]:                                                     # reference :364-372
                out.append({"query": q, "results": [], "stage1_results": [], "stage2_results": [],
                            "timing": self._timing((time.time() - t_all) / n if t_all else None, s1_each, None, None),
                            "performance_stats": self.performance_stats})
                continue
            if not stage2[b]:                                                     # reference :382-390
                out.append({"query": q, "results": [], "stage1_results": stage1[b], "stage2_results": [],
                            "timing": self._timing((time.time() - t_all) / n if t_all else None, s1_each, s2_each, None),
                            "performance_stats": self.performance_stats})
                continue
            t0 = now()
            final = self.stage3.rerank(q, stage2[b]) if self.stage3 is not None else stage2[b]
            s3 = time.time() - t0 if t0 else None
            final = final[:top_k]
            total = (s1_each + s2_each + s3) if t0 else None
            if self.enable_timing:
                self._update_stats(s1_each, s2_each, s3, total)
            out.append({"query": q, "results": final,
                        "stage1_results": stage1[b] if self.save_intermediate_results else [],
                        "stage2_results": stage2[b] if self.save_intermediate_results else [],
                        "timing": self._timing(total, s1_each, s2_each, s3),
                        "performance_stats": self.performance_stats.copy()})
        cleanup = getattr(self.pipeline, "_cleanup_memory", None)
        if cleanup is not None and getattr(getattr(self.pipeline, "config", None), "auto_cleanup", False):
            cleanup()                                                             # once per batch (reference: per query)
        return out

    def search(self, query: str, top_k: Optional[int] = None) -> Dict[str, Any]:
        return self.batch_search([query], top_k)[0]
