"""CPU oracle for the TriStage-RAG candidate-scoring hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker or as the timed CPU baseline -- never as a fallback for the CUDA path.

What it restates (reference paths are relative to /root/reference):

* Stage 1 (``src/stage1_retriever.py``): row normalisation ``x/(|x|+1e-8)``
  (:285-288) and ``faiss.IndexFlatIP.add/search`` as called at :276-277,:313,
  :380.  FAISS itself is a third-party dependency that is NOT vendored in the
  reference (``requirements.txt:10`` ``faiss-cpu>=1.7.0``, unpinned, no lock
  file) and is not installable here, so its published semantics are restated
  in ``flat_ip.py``: fp32 inner product, k results per query in descending
  score order, int64 labels, ``-1`` labels (score = lowest float) when
  ``k > ntotal``.
* Stage 1, approximate mode (``src/stage1_retriever.py:262-273``): ``faiss.IndexIVFFlat`` over an
  ``IndexFlatIP`` quantizer, restated in ``ivf.py`` (k-means training, list assignment, probing, list scan).
* Stage 2 (``src/stage2_rescorer.py``): ``_maxsim_score`` (:167-183) and
  ``_colbert_score`` (:185-201), plus the stable descending sort + truncate of
  ``rescore_candidates`` (:294-297).

Pinning status: the reference ships no tests and no golden vectors (SURVEY.md
§4, §8c) -- "parity unpinned" by the reference's own tests.  The Stage-2
restatement and the Stage-1 Python around the index ARE pinned against the
reference's own code, imported from /root/reference in the authoring container
by ``oracle/gen_golden.py`` (fixtures under ``tests/golden/``).  The
``IndexFlatIP`` arithmetic itself can only be pinned against its published
semantics (FAISS is absent): "parity unpinned" for that one call.
"""
