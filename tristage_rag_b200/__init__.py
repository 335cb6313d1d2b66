"""tristage_rag_b200 -- B200-native candidate-scoring hot path of TriStage-RAG.

Only what the path needs: the CUDA kernels + C ABI (``csrc/``, ``lib/``), the
ctypes binding (``_lib``), the drop-in ``Stage1Retriever`` / ``ColBERTScorer``
classes mirroring ``/root/reference/src/stage1_retriever.py`` and
``src/stage2_rescorer.py``, and the row-sharded multi-GPU wrappers (``dist``).
"""
from ._lib import IVF, Index, TokStore, TristageError, rank_desc, topk_merge  # noqa: F401
from .stage1_retriever import BM25Index, IndexFlatIP, IndexIVFFlat, Stage1Config, Stage1Retriever  # noqa: F401
from .stage2_rescorer import ColBERTScorer, Stage2Config  # noqa: F401

__all__ = ["Index", "IVF", "TokStore", "TristageError", "rank_desc", "topk_merge", "BM25Index", "IndexFlatIP", "IndexIVFFlat",
           "Stage1Config", "Stage1Retriever", "ColBERTScorer", "Stage2Config"]
