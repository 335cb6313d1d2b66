"""Deterministic stand-ins for what is absent in this image (test infrastructure).

* ``fake_faiss_module()`` -- a module object exposing ``IndexFlatIP``,
  ``IndexIVFFlat``, ``METRIC_INNER_PRODUCT``, ``write_index``/``read_index``
  backed by the ``oracle.flat_ip`` restatement, so the UNMODIFIED reference
  ``Stage1Retriever`` runs end to end on the CPU (gen_golden.py).
* ``FakeSentenceEncoder`` -- hashed bag-of-words sentence embeddings with the
  ``SentenceTransformer.encode`` surface the reference calls
  (src/stage1_retriever.py:236-248, benchmark/tristage_mteb_model.py:187).
* ``FakeTokenizer`` / ``FakeTokenModel`` -- the ``AutoTokenizer``/``AutoModel``
  surface ``ColBERTScorer`` calls (src/stage2_rescorer.py:105-111,140-165,
  215-231): per-token hashed embeddings as ``last_hidden_state``.

* ``FakeReranker`` -- the ``rerank`` surface of the Stage-3 cross-encoder wrapper
  (src/stage3_reranker.py:230-264) with a word-overlap score.

No model weights and no network exist here; these make BASELINE config #1
(non_mcp/test_docs.json through the pipeline) runnable and reproducible.
"""
from __future__ import annotations

import hashlib
import pickle
import re
import types

import numpy as np

from . import flat_ip


def _seed(token: str) -> int:
    return int.from_bytes(hashlib.sha256(token.encode("utf-8")).digest()[:8], "little")


def _token_vec(token: str, dim: int) -> np.ndarray:
    return np.random.default_rng(_seed(token)).standard_normal(dim).astype(np.float32)


def _words(text: str) -> list[str]:
    return re.sub(r"[^a-z0-9\s]", " ", text.lower()).split()


class FakeSentenceEncoder:
    def __init__(self, dim: int = 768):
        self.dim = dim
        self._cache: dict[str, np.ndarray] = {}

    def get_sentence_embedding_dimension(self) -> int:
        return self.dim

    def _vec(self, tok: str) -> np.ndarray:
        v = self._cache.get(tok)
        if v is None:
            v = self._cache[tok] = _token_vec(tok, self.dim)
        return v

    def encode(self, texts, batch_size: int = 32, convert_to_numpy: bool = True,
               show_progress_bar: bool = False, normalize_embeddings: bool = False, **_):
        single = isinstance(texts, str)
        if single:
            texts = [texts]
        out = np.zeros((len(texts), self.dim), np.float32)
        for i, t in enumerate(texts):
            toks = _words(t) or ["empty"]
            for w in toks:
                out[i] += self._vec(w)
            out[i] += 0.05 * self._vec("§bias§")
        if normalize_embeddings:
            out /= np.linalg.norm(out, axis=1, keepdims=True) + 1e-12
        return out[0] if single else out


class _Batch(dict):
    def to(self, device):
        return self


class FakeTokenizer:
    """Whitespace tokenizer with [CLS]/[SEP] ids; pads with id 0."""

    def __init__(self):
        self.vocab: dict[str, int] = {"[PAD]": 0, "[CLS]": 1, "[SEP]": 2}
        self.inv: list[str] = ["[PAD]", "[CLS]", "[SEP]"]

    def _id(self, w: str) -> int:
        i = self.vocab.get(w)
        if i is None:
            i = self.vocab[w] = len(self.inv)
            self.inv.append(w)
        return i

    def __call__(self, texts, truncation=True, padding=False, max_length=192, return_tensors="pt", **_):
        import torch

        if isinstance(texts, str):
            texts = [texts]
        rows = []
        for t in texts:
            ids = [1] + [self._id(w) for w in _words(t)][: max(0, max_length - 2)] + [2]
            rows.append(ids[:max_length])
        L = max(len(r) for r in rows)
        ids = torch.zeros((len(rows), L), dtype=torch.long)
        mask = torch.zeros((len(rows), L), dtype=torch.long)
        for i, r in enumerate(rows):
            ids[i, : len(r)] = torch.tensor(r)
            mask[i, : len(r)] = 1
        return _Batch(input_ids=ids, attention_mask=mask)


class FakeTokenModel:
    """``AutoModel`` stand-in: last_hidden_state[b, t] = hashed vector of token t
    plus a small position term (so repeated words differ slightly)."""

    def __init__(self, tokenizer: FakeTokenizer, hidden_size: int = 128):
        self.tok = tokenizer
        self.config = types.SimpleNamespace(hidden_size=hidden_size)
        self._cache: dict[int, np.ndarray] = {}

    def to(self, device):
        return self

    def eval(self):
        return self

    def _vec(self, tid: int) -> np.ndarray:
        v = self._cache.get(tid)
        if v is None:
            v = self._cache[tid] = _token_vec("tok:" + self.tok.inv[tid], self.config.hidden_size)
        return v

    def __call__(self, input_ids=None, attention_mask=None, **_):
        import torch

        B, L = input_ids.shape
        H = self.config.hidden_size
        out = np.zeros((B, L, H), np.float32)
        for b in range(B):
            for t in range(L):
                out[b, t] = self._vec(int(input_ids[b, t])) + 0.1 * _token_vec(f"pos:{t}", H)
        return types.SimpleNamespace(last_hidden_state=torch.from_numpy(out))


class FakeReranker:
    """Stage-3 stand-in with the ``rerank`` flow of the reference's cross-encoder wrapper
    (src/stage3_reranker.py:230-264: copy, ``stage3_score``, ``stage: "stage3"``, stable sort
    descending, truncate).  Stage 3 is outside the hot path; the score is a deterministic word
    overlap so the orchestration around it can be pinned."""

    def __init__(self, top_k_final: int = 20):
        self.config = types.SimpleNamespace(top_k_final=top_k_final, batch_size=32)

    def predict(self, query: str, documents) -> list:
        qw = set(_words(query))
        return [len(qw & set(_words(d))) / (1.0 + len(_words(d))) ** 0.5 for d in documents]

    def rerank(self, query: str, candidates):
        if not candidates:
            return []
        scores = self.predict(query, [c["document"] for c in candidates])
        out = []
        for c, s in zip(candidates, scores):
            u = c.copy()
            u["stage3_score"] = s
            u["stage"] = "stage3"
            out.append(u)
        out.sort(key=lambda x: x["stage3_score"], reverse=True)
        return out[: self.config.top_k_final]


class _IVFFlatAsExact(flat_ip.IndexFlatIP):
    """The reference switches to IndexIVFFlat for first batches > 1000 rows
    (src/stage1_retriever.py:262-273).  IVF is approximate and needs k-means;
    BASELINE.json pins EXACT search, so the stand-in is the exact scan."""

    def __init__(self, quantizer, d, nlist, metric):
        super().__init__(d)
        self.nlist, self.nprobe = nlist, 1

    def train(self, x):
        pass


def fake_faiss_module():
    m = types.ModuleType("faiss")
    m.IndexFlatIP = flat_ip.IndexFlatIP
    m.IndexIVFFlat = _IVFFlatAsExact
    m.METRIC_INNER_PRODUCT = 0

    def write_index(index, path):
        with open(path, "wb") as f:
            pickle.dump({"d": index.d, "x": index.xb}, f)

    def read_index(path):
        with open(path, "rb") as f:
            blob = pickle.load(f)
        idx = flat_ip.IndexFlatIP(blob["d"])
        if len(blob["x"]):
            idx.add(blob["x"])
        return idx

    m.write_index, m.read_index = write_index, read_index
    return m
