"""Drop-in ``Stage2Config`` / ``ColBERTScorer`` for
``/root/reference/src/stage2_rescorer.py`` with MaxSim scoring on a B200.

Same names, fields, methods and return shapes as the reference class driven by
``src/retrieval_pipeline.py:260-270,375`` and ``non_mcp/main.py:181-190,271``.
What changes underneath:

* the per-candidate loop over ``_maxsim_score`` / ``_colbert_score``
  (:167-201, :268-273) -> ONE ``ts_maxsim`` launch over all candidates
  (hand-written sm_100a kernel: TMA gather of ragged docs, tcgen05 contraction,
  fused row-max + mean/softmax epilogue);
* the reference re-encodes every candidate document on every query
  (:255-259); here token embeddings are kept in a GPU token store
  (``ts_tokstore``), filled the first time a document is seen -- or ahead of
  time with ``index_documents`` / ``add_token_embeddings`` -- so a query only
  pays for its own encoding;
* the final stable descending sort + truncate (:294-297) runs on the device
  (``ts_rank_desc``).

The HF tokenizer/model (:54-165, :207-242) stay the reference's PyTorch
encoder, outside the hot path; pass ``tokenizer=`` / ``model=`` to inject them
(tests use ``oracle/fakes.py``).  No CPU fallback for the scoring.
"""
from __future__ import annotations

import logging
import os
import threading
from dataclasses import dataclass, field
from typing import Any, Dict, Hashable, List, Optional

import numpy as np
import torch

from . import _lib


@dataclass
class Stage2Config:
    """Field-for-field the reference dataclass (src/stage2_rescorer.py:15-27),
    plus ``storage_dtype`` / ``gpu_index`` for the B200 token store."""
    model_name: str = "lightonai/GTE-ModernColBERT-v1"
    device: str = "auto"
    cache_dir: str = "./models"
    max_seq_length: int = 192
    batch_size: int = 16
    top_k_candidates: int = 100
    use_fp16: bool = True
    pooling_method: str = "cls"
    normalize_embeddings: bool = True
    scoring_method: str = "maxsim"   # "maxsim" or "colbert"
    use_gpu_if_available: bool = True
    storage_dtype: str = "bf16"      # HBM token-store dtype: bf16 | fp16 | fp32
    gpu_index: int = 0
    # one process per GPU (torchrun): every rank makes the same calls; a document's token embeddings live on ONE
    # rank (doc_id % world, or a hash of the text), which encodes and scores it; the [B, C] score matrices are
    # summed with one all-reduce, so every rank returns the identical result.  TS_SHARDED=1 flips the default.
    sharded: bool = field(default_factory=lambda: os.environ.get("TS_SHARDED", "0") not in ("", "0"))
    process_group: Any = None


class ColBERTScorer:
    """ColBERT-style late-interaction rescoring with GPU-resident token embeddings."""

    def __init__(self, config: Stage2Config, tokenizer=None, model=None):
        self.config = config
        self.logger = logging.getLogger(__name__)
        self.model = model
        self.tokenizer = tokenizer
        self.device = self._get_device()
        self.use_amp = False
        _lib.lib()                        # fail loudly if the CUDA library is missing
        self._store: Optional[_lib.TokStore] = None
        self._slot: Dict[Hashable, int] = {}      # doc key -> position in the token store
        # "which keys are missing -> encode -> append to the store -> remember the slots" must be one step: two
        # threads (the reference's Flask UI is threaded) that both read the same ``store.ndocs`` would map
        # their keys onto each other's documents, and the wrong slots would stay cached
        self._index_lock = threading.RLock()
        self._truncation_warned = False
        if self.config.max_seq_length > _lib.TS_S2_MAX_LD:
            self.logger.warning(f"max_seq_length={self.config.max_seq_length} exceeds the token store's limit of "
                                f"{_lib.TS_S2_MAX_LD} tokens per document: longer documents are cut to "
                                f"{_lib.TS_S2_MAX_LD} tokens (the reference scores all of them)")
        self._load_model()

    # -- encoder side: outside the hot path -----------------------------------
    def _get_device(self) -> str:
        if self.config.device == "auto":
            return "cuda" if (torch.cuda.is_available() and self.config.use_gpu_if_available) else "cpu"
        return self.config.device

    def _load_model(self):
        if self.model is None or self.tokenizer is None:
            from transformers import AutoModel, AutoTokenizer

            base = os.path.join(self.config.cache_dir, os.path.basename(self.config.model_name))
            legacy = os.path.join(self.config.cache_dir, self.config.model_name)
            source = base if os.path.isdir(base) else (legacy if os.path.isdir(legacy) else self.config.model_name)
            self.tokenizer = AutoTokenizer.from_pretrained(source, cache_dir=self.config.cache_dir)
            self.model = AutoModel.from_pretrained(source, cache_dir=self.config.cache_dir)
            self.model.to(self.device)
            self.model.eval()
            self.use_amp = self.config.use_fp16 and self.device == "cuda"

    def _forward(self, encoded):
        with torch.no_grad():
            if self.use_amp:
                with torch.autocast("cuda"):
                    return self.model(**encoded)
            return self.model(**encoded)

    def _encode_single_text(self, text: str) -> torch.Tensor:
        """[1, L, H] hidden states of the un-padded text (reference :134-165)."""
        if not text or not text.strip():
            text = "empty"
        encoded = self.tokenizer(text, truncation=True, max_length=self.config.max_seq_length,
                                 return_tensors="pt", padding=False)
        encoded = {k: v.to(self.device) for k, v in encoded.items()}
        out = self._forward(encoded)
        n = int(encoded["attention_mask"].sum().item())
        return out.last_hidden_state[:, :n, :]

    def encode_query(self, query: str) -> torch.Tensor:
        return self._encode_single_text(query)

    def encode_single_document(self, document: str) -> torch.Tensor:
        return self._encode_single_text(document)

    def encode_documents_batch(self, documents: List[str]) -> List[torch.Tensor]:
        """List of [L_i, H] hidden states, padding removed (reference :207-242)."""
        all_emb = []
        for i in range(0, len(documents), self.config.batch_size):
            batch = [t if t and t.strip() else "empty" for t in documents[i:i + self.config.batch_size]]
            encoded = self.tokenizer(batch, truncation=True, padding=True,
                                     max_length=self.config.max_seq_length, return_tensors="pt")
            encoded = {k: v.to(self.device) for k, v in encoded.items()}
            out = self._forward(encoded)
            for j in range(len(batch)):
                n = int(encoded["attention_mask"][j].sum().item())
                all_emb.append(out.last_hidden_state[j, :n, :])
        return all_emb

    # -- token store -----------------------------------------------------------
    def _ensure_store(self, dim: int) -> _lib.TokStore:
        if self._store is None:
            self._store = _lib.TokStore(dim, self.config.storage_dtype, self.config.gpu_index)
        elif self._store.dim != dim:
            raise ValueError(f"token dim {dim} != store dim {self._store.dim}")
        return self._store

    def add_token_embeddings(self, keys: List[Hashable], token_embeddings: List[Any]) -> None:
        """Insert pre-computed per-document token matrices ([L_i, H] each, raw
        hidden states: they are L2-normalised on the device at ingest)."""
        mats = [np.asarray(t.detach().float().cpu().numpy() if isinstance(t, torch.Tensor) else t, np.float32)
                .reshape(-1, np.shape(t)[-1]) for t in token_embeddings]
        if not mats:
            return
        n_long = sum(1 for m in mats if m.shape[0] > _lib.TS_S2_MAX_LD)
        if n_long and not self._truncation_warned:
            self._truncation_warned = True
            self.logger.warning(f"{n_long} document(s) longer than {_lib.TS_S2_MAX_LD} tokens were cut to that length "
                                "(token-store limit); their scores differ from the reference's")
        mats = [m[: _lib.TS_S2_MAX_LD] for m in mats]
        with self._index_lock:
            store = self._ensure_store(mats[0].shape[1])
            base = store.ndocs
            store.add(np.concatenate(mats, axis=0), [m.shape[0] for m in mats], normalize=True)
            for i, key in enumerate(keys):
                self._slot[key] = base + i

    def _shard(self):
        if not self.config.sharded:
            return None
        import torch.distributed as dist

        if not dist.is_initialized() or dist.get_world_size(self.config.process_group) < 2:
            return None
        g = self.config.process_group
        return dist.get_rank(g), dist.get_world_size(g), g

    @staticmethod
    def _owner(key: Hashable, world: int) -> int:
        """rank that keeps the token embeddings of a document: by doc id when it is an int, else by its text"""
        import zlib

        ident = key[0] if isinstance(key, tuple) else key
        if isinstance(ident, (int, np.integer)) and not isinstance(ident, bool):
            return int(ident) % world
        text = key[1] if isinstance(key, tuple) and len(key) > 1 else str(key)
        return zlib.crc32(str(text).encode("utf-8", "replace")) % world

    def _sum_over_ranks(self, scores: np.ndarray, shard) -> np.ndarray:
        """every candidate was scored by exactly one rank (0.0 elsewhere): one all-reduce assembles the matrix"""
        import torch.distributed as dist

        _, _, group = shard
        on_gpu = dist.get_backend(group) == "nccl"
        t = torch.from_numpy(np.ascontiguousarray(scores, np.float32))
        t = t.to(torch.device("cuda", self.config.gpu_index)) if on_gpu else t
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return t.cpu().numpy()

    def _ensure_indexed(self, pairs) -> None:
        """``pairs`` = (key, text) of the documents a request needs: the ones the store does not hold yet are
        encoded and appended, atomically with respect to other threads doing the same.  Sharded: only the
        documents this rank owns."""
        shard = self._shard()
        if shard is not None:
            pairs = [(k, d) for k, d in pairs if self._owner(k, shard[1]) == shard[0]]
        with self._index_lock:
            seen, todo = set(), []
            for k, d in pairs:
                if k not in self._slot and k not in seen:
                    seen.add(k)
                    todo.append((k, d))
            if todo:
                self.add_token_embeddings([k for k, _ in todo], self.encode_documents_batch([d for _, d in todo]))

    def index_documents(self, documents: List[str], doc_ids: Optional[List[Hashable]] = None) -> None:
        """Encode documents once and keep their token embeddings on the GPU."""
        keys = [(i, d) for i, d in zip(doc_ids, documents)] if doc_ids is not None else [("text", d) for d in documents]
        self._ensure_indexed(zip(keys, documents))

    @staticmethod
    def _key(candidate: Dict[str, Any]) -> Hashable:
        # the text is part of the key: doc ids are reused after a clear_index
        return (candidate.get("doc_id", "text"), candidate["document"])

    def _mode(self) -> int:
        return _lib.TS_S2_MAXSIM if self.config.scoring_method == "maxsim" else _lib.TS_S2_COLBERT

    def _score_slots(self, query_embeddings: torch.Tensor, slots: List[int]) -> np.ndarray:
        """slots: position in THIS rank's token store, or -1 for a document another rank owns (scores 0.0 here)"""
        q = query_embeddings.detach().float().cpu().numpy().reshape(1, -1, query_embeddings.shape[-1])
        cand = np.asarray(slots, dtype=np.int64).reshape(1, -1)
        if self._store is None:                       # this rank owns none of the documents seen so far
            s = np.zeros(cand.shape[1], np.float32)
        else:
            s = self._store.maxsim_host(q, cand, mode=self._mode(), normalize_q=True)[0]
        shard = self._shard()
        return self._sum_over_ranks(s, shard) if shard is not None else s

    # -- the two scoring functions, one pair at a time (reference :167-201) ----
    def _pair_score(self, query_embeddings, doc_embeddings, mode: int) -> torch.Tensor:
        d = doc_embeddings
        d = d.reshape(-1, d.shape[-1])
        store = _lib.TokStore(d.shape[-1], self.config.storage_dtype, self.config.gpu_index)
        store.add(d.detach().float().cpu().numpy(), [d.shape[0]], normalize=True)
        q = query_embeddings.detach().float().cpu().numpy().reshape(1, -1, d.shape[-1])
        s = store.maxsim_host(q, np.zeros((1, 1), np.int64), mode=mode, normalize_q=True)
        return torch.tensor(float(s[0, 0]))

    def _maxsim_score(self, query_embeddings: torch.Tensor, doc_embeddings: torch.Tensor) -> torch.Tensor:
        return self._pair_score(query_embeddings, doc_embeddings, _lib.TS_S2_MAXSIM)

    def _colbert_score(self, query_embeddings: torch.Tensor, doc_embeddings: torch.Tensor) -> torch.Tensor:
        return self._pair_score(query_embeddings, doc_embeddings, _lib.TS_S2_COLBERT)

    # -- the call the pipeline makes -------------------------------------------
    def rescore_candidates(self, query: str, candidates: List[Dict[str, Any]]) -> List[Dict[str, Any]]:
        if not candidates:
            return []
        query_embeddings = self.encode_query(query)
        try:
            keys = [self._key(c) for c in candidates]
            self._ensure_indexed((k, c["document"]) for k, c in zip(keys, candidates))
        except Exception as e:                         # reference :260-263
            self.logger.error(f"Error encoding documents: {e}")
            return candidates
        scores = self._score_slots(query_embeddings, [self._slot.get(k, -1) for k in keys])
        scored = []
        for c, s in zip(candidates, scores):
            u = c.copy()
            u["stage2_score"] = float(s)
            u["stage"] = "stage2"
            scored.append(u)
        # stable descending sort + truncate on the device (reference :294-297)
        top_k = min(self.config.top_k_candidates, len(scored), _lib.TS_MAX_K)
        dev = torch.device("cuda", self.config.gpu_index)
        _, pos = _lib.rank_desc(torch.from_numpy(scores.reshape(1, -1)).to(dev), top_k, device=self.config.gpu_index)
        order = [int(p) for p in pos[0].cpu().tolist() if p >= 0]
        if self.config.top_k_candidates > _lib.TS_MAX_K and len(scored) > _lib.TS_MAX_K:
            taken = set(order)
            rest = sorted((i for i in range(len(scored)) if i not in taken),
                          key=lambda i: scored[i]["stage2_score"], reverse=True)
            order += rest[: self.config.top_k_candidates - len(order)]
        return [scored[i] for i in order]

    def rescore_candidates_batch(self, queries: List[str],
                                 candidates: List[List[Dict[str, Any]]]) -> List[List[Dict[str, Any]]]:
        """``rescore_candidates`` for a batch of queries with ONE scoring launch (the reference
        loops over queries, src/retrieval_pipeline.py:444-448).  Per-query results are identical
        to calling ``rescore_candidates`` query by query."""
        assert len(queries) == len(candidates)
        out: List[List[Dict[str, Any]]] = [[] for _ in queries]
        live = [b for b, c in enumerate(candidates) if c]
        if not live:
            return out
        q_embs = [self.encode_query(queries[b]) for b in live]
        try:
            self._ensure_indexed((self._key(c), c["document"]) for b in live for c in candidates[b])
        except Exception as e:                         # reference :260-263, per batch
            self.logger.error(f"Error encoding documents: {e}")
            return [list(c) for c in candidates]
        H = q_embs[0].shape[-1]
        lq = [int(q.reshape(-1, H).shape[0]) for q in q_embs]
        if max(lq) > _lib.TS_S2_MAX_LQ:                # rare: very long queries take the per-query path
            for b in live:
                out[b] = self.rescore_candidates(queries[b], candidates[b])
            return out
        Cmax = max(len(candidates[b]) for b in live)
        qbuf = np.zeros((len(live), max(lq), H), np.float32)
        cand = np.full((len(live), Cmax), -1, np.int64)
        n_cand = np.zeros(len(live), np.int32)
        for r, b in enumerate(live):
            qbuf[r, : lq[r]] = q_embs[r].detach().float().cpu().numpy().reshape(-1, H)
            slots = [self._slot.get(self._key(c), -1) for c in candidates[b]]
            cand[r, : len(slots)] = slots
            n_cand[r] = len(slots)
        if self._store is None:
            scores = np.zeros(cand.shape, np.float32)
        else:
            scores = self._store.maxsim_host(qbuf, cand, q_len=np.asarray(lq, np.int32), n_cand=n_cand,
                                             mode=self._mode(), normalize_q=True)
        shard = self._shard()
        if shard is not None:
            scores = self._sum_over_ranks(scores, shard)
        top_k = min(self.config.top_k_candidates, Cmax, _lib.TS_MAX_K)
        dev = torch.device("cuda", self.config.gpu_index)
        _, pos = _lib.rank_desc(torch.from_numpy(scores).to(dev), top_k,
                                n_cand=torch.from_numpy(n_cand).to(dev), device=self.config.gpu_index)
        pos = pos.cpu().numpy()
        for r, b in enumerate(live):
            if self.config.top_k_candidates > _lib.TS_MAX_K and len(candidates[b]) > _lib.TS_MAX_K:
                out[b] = self.rescore_candidates(queries[b], candidates[b])
                continue
            res = []
            for p_ in pos[r]:
                if p_ < 0:
                    break
                u = candidates[b][int(p_)].copy()
                u["stage2_score"] = float(scores[r, int(p_)])
                u["stage"] = "stage2"
                res.append(u)
            out[b] = res
        return out

    def compute_similarity_matrix(self, query: str, documents: List[str]) -> np.ndarray:
        query_embeddings = self.encode_query(query)
        self.index_documents(documents)
        slots = [self._slot.get(("text", d), -1) for d in documents]
        return np.array([float(s) for s in self._score_slots(query_embeddings, slots)])

    def get_model_info(self) -> Dict[str, Any]:
        return {
            "model_name": self.config.model_name,
            "device": self.device,
            "max_seq_length": self.config.max_seq_length,
            "use_fp16": self.use_amp,
            "pooling_method": self.config.pooling_method,
            "scoring_method": self.config.scoring_method,
            "batch_size": self.config.batch_size,
            "embedding_dim": self.model.config.hidden_size if self.model else None,
            # limits of the GPU token store / kernels the reference does not have
            "max_doc_tokens": _lib.TS_S2_MAX_LD,
            "max_top_k_on_device": _lib.TS_MAX_K,
        }

    def clear_gpu_memory(self):
        """Best-effort cache release; the token store is index state and stays."""
        if torch.cuda.is_available():
            torch.cuda.empty_cache()
