"""Drop-in ``Stage1Config`` / ``BM25Index`` / ``Stage1Retriever`` for
``/root/reference/src/stage1_retriever.py`` with the dense scoring on a B200.

Same names, fields, methods, return shapes and error behaviour as the
reference class that ``src/retrieval_pipeline.py:244-256,316,358`` and
``non_mcp/main.py:167-177,224,254`` drive.  What changes underneath:

* ``faiss.IndexFlatIP.add/search`` (:270,277,313,380) -> ``libtristage.so``
  (``ts_index_add`` / ``ts_index_search_host``): hand-written sm_100a kernels,
  corpus resident in HBM as bf16 (``storage_dtype``), fp32 accumulate, exact
  top-k fused into the scan.  No CPU fallback: without the library or a B200
  the constructor's first device call raises.
* search is exact by default.  The reference silently switches to an approximate
  ``IndexIVFFlat(nlist=100, nprobe=10)`` when the first batch has more than
  1000 rows (:262-273); BASELINE.json pins exact search, so that switch is
  opt-in here: ``Stage1Config.approximate=True`` reproduces the reference's rule
  with inverted lists over the same resident rows (``IndexIVFFlat`` below,
  ``csrc/ivf.cu``); ``get_stats()['faiss_index_type']`` names the index in use.
* the encoder (SentenceTransformer, :137-254) stays the reference's PyTorch
  model and is outside the hot path; pass ``model=`` to inject one (tests use
  ``oracle/fakes.py``), otherwise it is loaded like the reference does.
* BM25 / RRF / weighted fusion (:35-112, :326-366) are lexical Python work,
  not part of the accelerated path; they are re-implemented here with the
  same arithmetic so hybrid results match (an inverted index replaces the
  O(N) per-query Python scan, scores are bit-identical).

Extra (not in the reference): ``search_batch`` and ``add_embeddings`` expose the
batched regime the reference's batch-1 API cannot reach; with
``Stage1Config.hybrid_on_device`` the BM25 search and the rank fusion of a batch
also run as GPU kernels (``DeviceBM25``, ``ts_bm25_*`` / ``ts_hybrid_fuse_host``),
with results identical to the host path.
"""
from __future__ import annotations

import logging
import math
import os
import pickle
import re
from collections import defaultdict
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from . import _lib


@dataclass
class Stage1Config:
    """Field-for-field the reference dataclass (src/stage1_retriever.py:16-33),
    plus ``storage_dtype`` / ``gpu_index`` for the B200 index."""
    model_name: str = "google/embeddinggemma-300m"
    device: str = "auto"
    cache_dir: str = "./models"
    index_dir: str = "./faiss_index"
    top_k_candidates: int = 500
    batch_size: int = 32
    max_text_length: int = 512
    enable_bm25: bool = True
    bm25_top_k: int = 300
    fusion_method: str = "rrf"
    rrf_k: int = 60
    dense_weight: float = 0.7
    bm25_weight: float = 0.3
    use_fp16: bool = True
    nlist: int = 100
    nprobe: int = 10
    storage_dtype: str = "bf16"   # HBM corpus dtype: bf16 | fp16 | fp32
    gpu_index: int = 0
    hybrid_on_device: bool = False   # search_batch: BM25 search + RRF/weighted fusion as GPU kernels (ts_bm25_*)
    # the reference's index rule: IVF (nlist, nprobe) when the first batch has > 1000 rows.  The reference's callers
    # build this config with explicit keywords (src/retrieval_pipeline.py:244-255) and cannot pass new fields, so
    # the default can also be flipped from the environment of an unmodified deployment: TS_APPROXIMATE=1
    approximate: bool = field(default_factory=lambda: os.environ.get("TS_APPROXIMATE", "0") not in ("", "0"))
    # one process per GPU (torchrun): stripe the corpus rows over the ranks of the default (or the given) process
    # group.  Every rank makes the SAME add_documents / search calls; a rank encodes and keeps only its share of each
    # added batch, searches it, and the [B, k] lists are all-gathered and merged (dist.gather_topk + ts_topk_merge),
    # so every rank returns the identical, exact result.  TS_SHARDED=1 flips the default for an unmodified deployment.
    sharded: bool = field(default_factory=lambda: os.environ.get("TS_SHARDED", "0") not in ("", "0"))
    process_group: Any = None


class BM25Index:
    """BM25 with the reference's arithmetic (src/stage1_retriever.py:35-112).

    Reference behaviours kept on purpose:
    * ``fit`` appends to ``doc_freqs`` / ``doc_lens`` without clearing them, so a
      second ``fit`` (every later ``add_documents`` call, :316-322) leaves stale
      rows that skew ``avg_doc_len`` and the document frequencies;
    * a query token that occurs twice is counted twice (:93-99);
    * ``search`` ranks EVERY document, zero scores included, with a stable
      descending sort (:105-112).
    """

    def __init__(self, k1: float = 1.2, b: float = 0.75):
        self.k1 = k1
        self.b = b
        self.doc_freqs: List[Dict[str, int]] = []
        self.idf: Dict[str, float] = {}
        self.doc_lens: List[int] = []
        self.avg_doc_len = 0
        self.corpus_size = 0
        self.vocabulary = set()
        self.documents: List[str] = []
        self._postings = None          # CSR inverted index, built lazily by search()

    def tokenize(self, text: str) -> List[str]:
        return re.sub(r"[^a-z0-9\s]", " ", text.lower()).split()

    def fit(self, documents: List[str]):
        self.documents = documents
        self.corpus_size = len(documents)
        for doc in documents:
            tokens = self.tokenize(doc)
            self.vocabulary.update(tokens)
            tf: Dict[str, int] = defaultdict(int)
            for t in tokens:
                tf[t] += 1
            self.doc_freqs.append(tf)
            self.doc_lens.append(len(tokens))
        self.avg_doc_len = sum(self.doc_lens) / self.corpus_size if self.corpus_size > 0 else 0
        df: Dict[str, int] = defaultdict(int)
        for tf in self.doc_freqs:              # one pass instead of |V| passes; same counts
            for t in tf:
                df[t] += 1
        for token in self.vocabulary:
            d = df.get(token, 0)
            self.idf[token] = math.log((self.corpus_size - d + 0.5) / (d + 0.5) + 1.0)
        self._postings = None

    def score(self, query: str, doc_idx: int) -> float:
        if doc_idx >= len(self.doc_freqs):
            return 0.0
        tf_map = self.doc_freqs[doc_idx]
        dl = self.doc_lens[doc_idx]
        s = 0.0
        for token in self.tokenize(query):
            if token in tf_map and token in self.idf:
                tf = tf_map[token]
                s += self.idf[token] * ((tf * (self.k1 + 1)) / (tf + self.k1 * (1 - self.b + self.b * dl / self.avg_doc_len)))
        return s

    def _build_postings(self):
        """Inverted index as CSR arrays: for term t, ``_docs[_off[t]:_off[t+1]]`` are the documents
        that contain it (ascending) and ``_w`` the per-posting score contribution
        ``idf * tf*(k1+1) / (tf + k1*(1 - b + b*dl/avgdl))`` -- the reference's expression (:93-99)
        evaluated once per posting in IEEE double, same operation order, so sums are bit-identical."""
        n = min(len(self.documents), len(self.doc_freqs))
        by_term: Dict[str, List[Tuple[int, int]]] = defaultdict(list)
        for idx in range(n):
            for t, tf in self.doc_freqs[idx].items():
                by_term[t].append((idx, tf))
        self._tid = {t: i for i, t in enumerate(by_term)}
        off = np.zeros(len(by_term) + 1, np.int64)
        docs, tfs, idfs = [], [], []
        for i, (t, plist) in enumerate(by_term.items()):
            off[i + 1] = off[i] + len(plist)
            docs.extend(d for d, _ in plist)
            tfs.extend(tf for _, tf in plist)
            idfs.extend([self.idf.get(t, 0.0)] * len(plist))
        self._off = off
        self._docs = np.asarray(docs, np.int64)
        tf = np.asarray(tfs, np.int64)
        idf = np.asarray(idfs, np.float64)
        dl = np.asarray(self.doc_lens[:n], np.int64)[self._docs] if len(docs) else np.zeros(0, np.int64)
        if len(docs) and self.avg_doc_len:
            self._w = idf * ((tf * (self.k1 + 1)) / (tf + self.k1 * (1 - self.b + self.b * dl / self.avg_doc_len)))
        else:
            self._w = np.zeros(len(docs), np.float64)
        self._postings = True

    def search(self, query: str, top_k: int = 10) -> List[Tuple[int, float]]:
        """The reference scores EVERY document in a Python loop and sorts all of them (:103-112);
        here only documents that share a token with the query are touched.  Ranking is the same
        stable descending sort: equal scores (all the zeros included) keep ascending index order."""
        if not getattr(self, "_postings", None) or getattr(self, "_tid", None) is None:
            self._build_postings()
        n = len(self.documents)
        scores = np.zeros(n, np.float64)
        for token in self.tokenize(query):     # same accumulation order as score(): query-token order
            t = self._tid.get(token)
            if t is None or token not in self.idf:
                continue
            sl = slice(self._off[t], self._off[t + 1])
            scores[self._docs[sl]] += self._w[sl]          # a document occurs once per term
        top_k = max(int(top_k), 0)
        if len(self._w) and (self._w <= 0).any():
            order = np.argsort(-scores, kind="stable")[:top_k]      # stale-fit quirk can make idf <= 0
        else:
            hit = np.flatnonzero(scores > 0)
            order = hit[np.argsort(-scores[hit], kind="stable")][:top_k]
            if len(order) < min(top_k, n):                          # pad with zero-score docs, ascending index
                zero = np.flatnonzero(scores[: top_k + len(hit)] <= 0)[: min(top_k, n) - len(order)]
                order = np.concatenate([order, zero])
        return [(int(i), float(scores[i])) for i in order]


class DeviceBM25:
    """``BM25Index.search`` for batches of queries on the GPU (``ts_bm25_*``): the fitted index's CSR
    postings and per-posting fp64 weights are uploaded once per fit; results are bit-identical to
    ``BM25Index.search`` (same fp64 sums in query-token order, same stable ranking)."""

    def __init__(self, bm25: BM25Index, device: int = 0):
        if not getattr(bm25, "_postings", None) or getattr(bm25, "_tid", None) is None:
            bm25._build_postings()
        self.bm25 = bm25
        self.n_docs = len(bm25.documents)
        self._dev = _lib.BM25(self.n_docs, bm25._off, bm25._docs, bm25._w, device)   # raises if a weight is <= 0

    def term_ids(self, query: str) -> List[int]:
        tid, idf = self.bm25._tid, self.bm25.idf
        return [tid[t] for t in self.bm25.tokenize(query) if t in tid and t in idf]

    def search_arrays(self, queries: List[str], top_k: int):
        """(scores [B, top_k] float64, ids [B, top_k] int64; -1 beyond the corpus)."""
        return self._dev.search([self.term_ids(q) for q in queries], top_k)

    def search_batch(self, queries: List[str], top_k: int = 10) -> List[List[Tuple[int, float]]]:
        scores, ids = self.search_arrays(queries, top_k)
        return [[(int(i), float(s)) for i, s in zip(ids[b], scores[b]) if i >= 0] for b in range(len(queries))]


class IndexFlatIP:
    """The ``faiss.IndexFlatIP`` surface the reference touches (``d``, ``ntotal``,
    ``add``, ``search``), backed by one ``ts_index`` shard on the GPU."""

    def __init__(self, d: int, storage_dtype: str = "bf16", device: int = 0, reserve_rows: int = 0):
        self.d = int(d)
        self.storage_dtype = storage_dtype
        self.device = device
        self._index = _lib.Index(self.d, storage_dtype, "ip", device, reserve_rows)
        self.is_trained = True

    @property
    def ntotal(self) -> int:
        return self._index.ntotal

    def add(self, x: np.ndarray) -> None:
        self._index.add(np.ascontiguousarray(x, dtype=np.float32), normalize=False)

    def search(self, q: np.ndarray, k: int, path: str = "auto"):
        if k > _lib.TS_MAX_K:
            raise ValueError(f"top_k={k} exceeds the fused top-k limit {_lib.TS_MAX_K}")
        return self._index.search_host(q, k, normalize_q=False, path=path)

    def save(self, path: str) -> None:
        self._index.save(path)
        if os.path.exists(path + ".ivf.npz"):      # lists of an earlier approximate index at the same place
            os.remove(path + ".ivf.npz")

    @classmethod
    def load(cls, path: str, storage_dtype: str = "bf16", device: int = 0) -> "IndexFlatIP":
        obj = cls.__new__(cls)
        obj._index = _lib.Index.load(path, device)
        obj.d, obj.storage_dtype, obj.device, obj.is_trained = obj._index.dim, storage_dtype, device, True
        return obj


class ShardedFlatIP:
    """``IndexFlatIP`` whose rows are striped over the ranks of a ``torch.distributed`` group: ``add_local`` appends
    this rank's share of a batch together with the GLOBAL row numbers of those rows, ``search`` runs the local exact
    top-k, maps local rows to global ids, all-gathers the [B, k] lists and merges them on the device
    (``ts_topk_merge``; tie rule score desc, id asc) -- the same result on every rank, equal to one flat index over
    all rows because the global top-k is a subset of the union of the local top-k lists."""

    def __init__(self, d: int, storage_dtype: str = "bf16", device: int = 0, group=None, merge_fn=None, local=None):
        self.d, self.storage_dtype, self.device, self.group = int(d), storage_dtype, device, group
        self.local = local if local is not None else IndexFlatIP(d, storage_dtype, device)
        self.merge_fn = merge_fn
        self._gids = np.zeros(0, np.int64)
        self.ntotal = 0                    # GLOBAL row count (what faiss's ntotal means to the callers)
        self.is_trained = True

    def add_local(self, x_local: np.ndarray, gids: np.ndarray, ntotal_after: int) -> None:
        if len(x_local):
            self.local.add(x_local)
            self._gids = np.concatenate([self._gids, np.asarray(gids, np.int64)])
        self.ntotal = int(ntotal_after)

    def search(self, q: np.ndarray, k: int, path: str = "auto"):
        import torch
        import torch.distributed as dist

        from . import dist as tdist

        q = np.ascontiguousarray(q, dtype=np.float32)
        B = q.shape[0]
        if self.local.ntotal > 0:
            D, I = self.local.search(q, k, path=path)
            I = np.where(I >= 0, self._gids[np.maximum(I, 0)], -1)
        else:                                   # this rank holds no rows (fewer rows than ranks)
            D = np.full((B, k), np.float32(-3.4028234663852886e38), np.float32)
            I = np.full((B, k), -1, np.int64)
        on_gpu = dist.get_backend(self.group) == "nccl"
        dev = torch.device("cuda", self.device) if on_gpu else torch.device("cpu")
        all_s, all_i = tdist.gather_topk(torch.from_numpy(D).to(dev), torch.from_numpy(I).to(dev), self.group)
        merge = self.merge_fn or (lambda s, i: _lib.topk_merge(s, i, device=self.device))
        ms, mi = merge(all_s, all_i)
        return ms.cpu().numpy(), mi.cpu().numpy()

    def save(self, path: str) -> None:
        raise NotImplementedError("a sharded Stage-1 index is saved per rank with dist.ShardedIndex.save")


class IndexIVFFlat:
    """The ``faiss.IndexIVFFlat`` surface the reference touches (``train``, ``add``, ``nprobe``, ``search``,
    ``ntotal``; reference :263-273,313,380), backed by one ``ts_index`` shard plus inverted lists of row numbers
    over it (``ts_ivf``).  The rows are stored once; ``exact_search`` scans all of them."""

    def __init__(self, d: int, nlist: int, storage_dtype: str = "bf16", device: int = 0, nprobe: int = 1):
        self.d, self.nlist, self.nprobe = int(d), int(nlist), int(nprobe)      # faiss default nprobe = 1
        self.storage_dtype, self.device = storage_dtype, device
        self._index = _lib.Index(self.d, storage_dtype, "ip", device)
        self._ivf = _lib.IVF(self._index, self.nlist)

    @property
    def is_trained(self) -> bool:
        return self._ivf.is_trained

    @property
    def ntotal(self) -> int:
        return self._index.ntotal

    def train(self, x: np.ndarray) -> None:
        from .ivf import train_centroids

        self._ivf.set_centroids(train_centroids(np.ascontiguousarray(x, dtype=np.float32), self.nlist))

    def add(self, x: np.ndarray) -> None:
        if not self.is_trained:
            raise RuntimeError("IndexIVFFlat.add before train")      # faiss asserts is_trained
        # the rows are appended now and put into their lists by the next search / save (ts_ivf_search syncs by itself):
        # ingest in many small batches does not rebuild the lists once per batch
        self._index.add(np.ascontiguousarray(x, dtype=np.float32), normalize=False)

    def search(self, q: np.ndarray, k: int, path: str = "auto"):
        if k > _lib.TS_MAX_K:
            raise ValueError(f"top_k={k} exceeds the fused top-k limit {_lib.TS_MAX_K}")
        return self._ivf.search_host(q, k, self.nprobe, normalize_q=False)

    def exact_search(self, q: np.ndarray, k: int, path: str = "auto"):
        return self._index.search_host(q, k, normalize_q=False, path=path)

    def save(self, path: str) -> None:
        self._ivf.sync()
        self._index.save(path)
        tmp = path + ".ivf.tmp.npz"
        np.savez(tmp, centroids=self._ivf.centroids(), assign=self._ivf.assignments(),
                 nlist=np.int64(self.nlist), nprobe=np.int64(self.nprobe))
        os.replace(tmp, path + ".ivf.npz")

    @classmethod
    def from_parts(cls, index: "_lib.Index", centroids: np.ndarray, assign: np.ndarray, nprobe: int,
                   storage_dtype: str, device: int) -> "IndexIVFFlat":
        obj = cls.__new__(cls)
        obj.d, obj.nlist, obj.nprobe = index.dim, int(centroids.shape[0]), int(nprobe)
        obj.storage_dtype, obj.device, obj._index = storage_dtype, device, index
        obj._ivf = _lib.IVF(index, obj.nlist)
        obj._ivf.set_centroids(centroids)
        obj._ivf.set_assignments(assign)
        return obj

    @classmethod
    def load(cls, path: str, storage_dtype: str = "bf16", device: int = 0) -> "IndexIVFFlat":
        with np.load(path + ".ivf.npz") as z:
            return cls.from_parts(_lib.Index.load(path, device), z["centroids"], z["assign"], int(z["nprobe"]),
                                  storage_dtype, device)


class Stage1Retriever:
    """Stage 1: dense embeddings + exact GPU top-k (+ optional BM25 fusion)."""

    def __init__(self, config: Stage1Config, model=None):
        self.config = config
        self.logger = logging.getLogger(__name__)
        self.model = model
        self.embedding_dim = None
        self.faiss_index = None
        self.bm25_index = None
        self.documents: List[str] = []
        self.doc_metadata: List[Dict[str, Any]] = []
        self._device_bm25 = None
        self._shard_merge_fn = None       # tests on a CPU-only box inject the G*k -> k merge
        os.makedirs(self.config.cache_dir, exist_ok=True)
        os.makedirs(self.config.index_dir, exist_ok=True)
        _lib.lib()                       # fail loudly now if the CUDA library is missing
        self._load_model()

    # -- encoder: outside the hot path (reference :137-254) ------------------
    def _load_model(self):
        if self.model is None:
            from sentence_transformers import SentenceTransformer  # not in this image: inject model=

            device = self.config.device
            if device == "auto":
                import torch

                device = "cuda" if torch.cuda.is_available() else "cpu"
            base = os.path.join(self.config.cache_dir, os.path.basename(self.config.model_name))
            legacy = os.path.join(self.config.cache_dir, self.config.model_name)
            source = base if os.path.isdir(base) else (legacy if os.path.isdir(legacy) else self.config.model_name)
            self.model = SentenceTransformer(source, device=device, cache_folder=self.config.cache_dir)
        if hasattr(self.model, "get_sentence_embedding_dimension"):
            self.embedding_dim = self.model.get_sentence_embedding_dimension()
        else:
            self.embedding_dim = self.model.encode("sample text", convert_to_numpy=True).shape[0]
        self.logger.info(f"Model loaded successfully. Embedding dimension: {self.embedding_dim}")

    def _encode_batch(self, texts: List[str]) -> np.ndarray:
        emb = self.model.encode(texts, batch_size=self.config.batch_size, convert_to_numpy=True,
                                show_progress_bar=False)
        return np.asarray(emb).astype(np.float32)

    def _normalize_embeddings(self, embeddings: np.ndarray) -> np.ndarray:
        """x / (|x| + 1e-8), fp32 numpy -- reference :285-288."""
        norms = np.linalg.norm(embeddings, axis=1, keepdims=True)
        return embeddings / (norms + 1e-8)

    def _shard(self):
        """(rank, world, group) when the corpus is striped over a process group, else None"""
        if not self.config.sharded:
            return None
        import torch.distributed as dist

        if not dist.is_initialized() or dist.get_world_size(self.config.process_group) < 2:
            return None
        g = self.config.process_group
        return dist.get_rank(g), dist.get_world_size(g), g

    def _add_sharded(self, texts: Optional[List[str]], embeddings: Optional[np.ndarray], n: int, base: int, shard):
        """this rank's share of a batch of n rows whose global row numbers start at `base`"""
        from .dist import shard_range

        rank, world, group = shard
        lo, hi = shard_range(n, rank, world)
        if embeddings is None:
            mine = self._normalize_embeddings(self._encode_batch(texts[lo:hi])).astype(np.float32) if hi > lo \
                else np.zeros((0, self.embedding_dim), np.float32)
        else:
            mine = embeddings[lo:hi]
        if self.faiss_index is None:
            self.faiss_index = ShardedFlatIP(self.embedding_dim if embeddings is None else embeddings.shape[1],
                                             self.config.storage_dtype, self.config.gpu_index, group, self._shard_merge_fn)
        self.faiss_index.add_local(mine, np.arange(base + lo, base + hi, dtype=np.int64), base + n)

    def _create_faiss_index(self, embeddings: np.ndarray):
        from .ivf import MIN_ROWS_FOR_IVF

        if self.config.approximate and len(embeddings) > MIN_ROWS_FOR_IVF:
            # the reference's rule (:262-273): IVF trained on the first batch, nprobe from the config
            self.faiss_index = IndexIVFFlat(embeddings.shape[1], self.config.nlist, self.config.storage_dtype,
                                            self.config.gpu_index)
            self.faiss_index.train(embeddings)
            self.faiss_index.add(embeddings)
            self.faiss_index.nprobe = self.config.nprobe
            self.logger.info(f"GPU IVF index created with {len(embeddings)} vectors")
            return
        self.faiss_index = IndexFlatIP(embeddings.shape[1], self.config.storage_dtype, self.config.gpu_index)
        self.faiss_index.add(embeddings)
        self.logger.info(f"GPU flat index created with {len(embeddings)} vectors")

    # -- ingest ---------------------------------------------------------------
    def add_documents(self, documents: List[str], metadata: Optional[List[Dict[str, Any]]] = None):
        if not documents:
            return
        self.documents.extend(documents)
        if metadata is None:
            metadata = [{}] * len(documents)   # one shared dict, like the reference (:302)
        self.doc_metadata.extend(metadata)
        shard = self._shard()
        if shard is not None:                 # every rank encodes and keeps only its share of the batch
            self._add_sharded(documents, None, len(documents), len(self.documents) - len(documents), shard)
        else:
            embeddings = self._normalize_embeddings(self._encode_batch(documents)).astype(np.float32)
            self._add_normalized(embeddings)
        self._refit_bm25()

    def add_embeddings(self, embeddings: np.ndarray, documents: Optional[List[str]] = None,
                       metadata: Optional[List[Dict[str, Any]]] = None, normalize: bool = True):
        """Ingest pre-computed embeddings (synthetic corpora, offline encoders)."""
        n = len(embeddings)
        docs = list(documents) if documents is not None else [f"doc-{len(self.documents) + i}" for i in range(n)]
        assert len(docs) == n
        self.documents.extend(docs)
        self.doc_metadata.extend(metadata if metadata is not None else [{}] * n)
        e = np.asarray(embeddings, dtype=np.float32)
        e = self._normalize_embeddings(e).astype(np.float32) if normalize else e
        shard = self._shard()
        if shard is not None:
            self._add_sharded(None, e, n, len(self.documents) - n, shard)
        else:
            self._add_normalized(e)
        if documents is not None:
            self._refit_bm25()

    def _add_normalized(self, embeddings: np.ndarray):
        if self.faiss_index is None:
            self._create_faiss_index(embeddings)
        else:
            self.faiss_index.add(embeddings)

    def _refit_bm25(self):
        if self.config.enable_bm25:
            if self.bm25_index is None:
                self.bm25_index = BM25Index()
            self.bm25_index.fit(self.documents)
        self._device_bm25 = None           # postings changed: re-upload lazily

    def _hybrid_batch_on_device(self, texts: List[str], D: np.ndarray, I: np.ndarray, top_k: int):
        """search_batch's BM25 + fusion step as GPU kernels.  Returns per-query fused (doc, score) lists, or
        None when this index / request has to take the host path (weights <= 0 after a stale refit, list
        sizes beyond the kernels' limits, a weighted fusion whose normaliser is 0 -- the reference raises
        ZeroDivisionError there and so does the host path)."""
        cfg = self.config
        k2 = cfg.bm25_top_k
        if k2 < 1 or k2 > _lib.TS_BM25_MAX_K or top_k + k2 > _lib.TS_FUSE_MAX or not self.bm25_index.documents:
            return None
        if getattr(self, "_device_bm25", None) is None:
            try:
                self._device_bm25 = DeviceBM25(self.bm25_index, cfg.gpu_index)
            except _lib.TristageError as e:
                if e.code != -4:                        # TS_ERR_UNSUPPORTED: weights <= 0
                    raise
                return None
        try:
            bs, bi = self._device_bm25.search_arrays(texts, k2)
        except _lib.TristageError as e:
            if e.code != -3:                            # TS_ERR_NOMEM: no room for the score accumulator -> host path
                raise
            self.logger.warning(f"device BM25 out of memory ({e}); this batch takes the host path")
            return None
        if cfg.fusion_method != "rrf":
            if (D[:, 0] == 0).any() or (bs[:, 0] == 0).any() or (D.max(axis=1) == 0).any():
                return None
        ids, scores, n = _lib.hybrid_fuse(cfg.fusion_method, cfg.rrf_k, cfg.dense_weight, cfg.bm25_weight, I, D, bi, bs,
                                          top_k, cfg.gpu_index)
        return [[(int(ids[b, r]), float(scores[b, r])) for r in range(int(n[b]))] for b in range(len(texts))]

    # -- fusion (reference :326-366) -----------------------------------------
    def _reciprocal_rank_fusion(self, dense_results, bm25_results):
        scores: Dict[int, float] = defaultdict(float)
        for rank, (doc_idx, _) in enumerate(dense_results):
            scores[doc_idx] += 1.0 / (self.config.rrf_k + rank + 1)
        for rank, (doc_idx, _) in enumerate(bm25_results):
            scores[doc_idx] += 1.0 / (self.config.rrf_k + rank + 1)
        fused = list(scores.items())
        fused.sort(key=lambda x: x[1], reverse=True)
        return fused

    def _weighted_fusion(self, dense_results, bm25_results):
        scores: Dict[int, float] = defaultdict(float)
        if dense_results:
            mx = max(s for _, s in dense_results)
            for doc_idx, s in dense_results:
                scores[doc_idx] += self.config.dense_weight * (s / mx)
        if bm25_results:
            mx = max(s for _, s in bm25_results)
            for doc_idx, s in bm25_results:
                scores[doc_idx] += self.config.bm25_weight * (s / mx)
        fused = list(scores.items())
        fused.sort(key=lambda x: x[1], reverse=True)
        return fused

    # -- search ---------------------------------------------------------------
    def _format(self, final_results) -> List[Dict[str, Any]]:
        out = []
        for doc_idx, score in final_results:
            if doc_idx < len(self.documents):
                out.append({"doc_id": doc_idx, "document": self.documents[doc_idx], "score": score,
                            "stage1_score": score, "metadata": self.doc_metadata[doc_idx], "stage": "stage1"})
        return out

    def _fuse(self, query: str, dense_results, top_k: int):
        bm25_results = []
        if self.config.enable_bm25 and self.bm25_index is not None:
            bm25_results = self.bm25_index.search(query, self.config.bm25_top_k)
        if self.config.enable_bm25 and bm25_results:
            fused = (self._reciprocal_rank_fusion(dense_results, bm25_results)
                     if self.config.fusion_method == "rrf" else self._weighted_fusion(dense_results, bm25_results))
            return fused[:top_k]
        return dense_results[:top_k]

    def search(self, query: str, top_k: Optional[int] = None) -> List[Dict[str, Any]]:
        if self.faiss_index is None:
            raise ValueError("No documents indexed. Call add_documents() first.")
        top_k = top_k or self.config.top_k_candidates
        if self.config.hybrid_on_device and self.config.enable_bm25 and self.bm25_index is not None:
            return self.search_batch([query], top_k)[0]      # device BM25 + fusion; identical results
        q = self._normalize_embeddings(self._encode_batch([query]))
        D, I = self.faiss_index.search(q, top_k)
        dense = [(int(i), float(s)) for i, s in zip(I[0], D[0]) if i >= 0]
        return self._format(self._fuse(query, dense, top_k))

    def search_batch(self, queries, top_k: Optional[int] = None, path: str = "auto") -> List[List[Dict[str, Any]]]:
        """Batched search: ``queries`` is a list of strings or an fp32 [B, d]
        embedding matrix (already encoded).  One GPU call for the whole batch."""
        if self.faiss_index is None:
            raise ValueError("No documents indexed. Call add_documents() first.")
        top_k = top_k or self.config.top_k_candidates
        texts = list(queries) if not isinstance(queries, np.ndarray) else None
        emb = self._encode_batch(texts) if texts is not None else np.asarray(queries, np.float32)
        D, I = self.faiss_index.search(self._normalize_embeddings(emb), top_k, path=path)
        if (texts is not None and self.config.hybrid_on_device and self.config.enable_bm25
                and self.bm25_index is not None):
            fused = self._hybrid_batch_on_device(texts, D, I, top_k)
            if fused is not None:
                return [self._format(f) for f in fused]
        results = []
        for b in range(len(D)):
            dense = [(int(i), float(s)) for i, s in zip(I[b], D[b]) if i >= 0]
            final = self._fuse(texts[b], dense, top_k) if texts is not None else dense[:top_k]
            results.append(self._format(final))
        return results

    # -- persistence (reference :421-465; same file names and quirks) ---------
    def save_index(self, index_path: Optional[str] = None):
        if index_path is None:
            index_path = os.path.join(self.config.index_dir, "stage1_index.pkl")
        data = {"documents": self.documents, "doc_metadata": self.doc_metadata,
                "config": self.config.__dict__, "bm25_index": self.bm25_index}
        faiss_path = os.path.join(self.config.index_dir, "stage1_faiss.index")   # fixed location, like :434
        if isinstance(self.faiss_index, ShardedFlatIP):
            raise NotImplementedError("sharded Stage 1: save the shards with tristage_rag_b200.dist.ShardedIndex.save")
        if self.faiss_index is not None:
            self.faiss_index.save(faiss_path)
        with open(index_path, "wb") as f:
            pickle.dump(data, f)
        self.logger.info(f"Stage 1 index saved to {index_path}")

    def load_index(self, index_path: Optional[str] = None):
        if index_path is None:
            index_path = os.path.join(self.config.index_dir, "stage1_index.pkl")
        if not os.path.exists(index_path):
            self.logger.warning(f"Index file not found: {index_path}")
            return
        with open(index_path, "rb") as f:
            data = pickle.load(f)
        self.documents = data["documents"]
        self.doc_metadata = data["doc_metadata"]
        self.bm25_index = data.get("bm25_index")
        faiss_path = os.path.join(self.config.index_dir, "stage1_faiss.index")
        if os.path.exists(faiss_path):
            with open(faiss_path, "rb") as f:
                magic = f.read(8)
            if magic == b"TSSHARD2":
                kind = IndexIVFFlat if os.path.exists(faiss_path + ".ivf.npz") else IndexFlatIP
                self.faiss_index = kind.load(faiss_path, self.config.storage_dtype, self.config.gpu_index)
            elif magic[:4] == b"IwFl":
                # the reference's approximate index (:264): import vectors, centroids and lists.  With
                # approximate=False only the vectors are kept and every search is exact.
                from .faiss_io import read_faiss_ivf

                parts = read_faiss_ivf(faiss_path)
                self.faiss_index = None
                if self.config.approximate:
                    base = _lib.Index(parts["vectors"].shape[1], self.config.storage_dtype, "ip", self.config.gpu_index)
                    base.add(parts["vectors"], normalize=False)
                    self.faiss_index = IndexIVFFlat.from_parts(base, parts["centroids"], parts["assign"], parts["nprobe"],
                                                               self.config.storage_dtype, self.config.gpu_index)
                else:
                    self.faiss_index = IndexFlatIP(parts["vectors"].shape[1], self.config.storage_dtype,
                                                   self.config.gpu_index)
                    self.faiss_index.add(parts["vectors"])
            else:
                # a file the reference itself wrote with faiss.write_index (:436): import the vectors
                from .faiss_io import read_faiss_flat

                x, metric = read_faiss_flat(faiss_path)
                if metric != "ip":
                    raise ValueError(f"{faiss_path}: the reference uses inner-product indexes, this one is {metric}")
                self.faiss_index = None
                self._create_faiss_index(np.ascontiguousarray(x, np.float32))     # already normalised at ingest (:307)
        self._device_bm25 = None
        self.logger.info(f"Stage 1 index loaded from {index_path}")

    def get_stats(self) -> Dict[str, Any]:
        return {
            "total_documents": len(self.documents),
            "embedding_dimension": self.embedding_dim,
            "faiss_index_type": type(self.faiss_index).__name__ if self.faiss_index else None,
            "bm25_enabled": self.config.enable_bm25,
            "bm25_vocabulary_size": len(self.bm25_index.vocabulary) if self.bm25_index else 0,
            "config": self.config.__dict__,
        }
