"""Stage-2 oracle: late-interaction MaxSim / "colbert" scoring on the CPU.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Restates
``ColBERTScorer._maxsim_score`` (src/stage2_rescorer.py:167-183),
``_colbert_score`` (:185-201) and the ordering of ``rescore_candidates``
(:294-297).  Pinned against the reference's own functions by
``oracle/gen_golden.py`` -> ``tests/golden/stage2_reference.json``.
"""
from __future__ import annotations

import numpy as np

MODE_MAXSIM = 0
MODE_COLBERT = 1


def l2_normalize_tokens(x: np.ndarray, eps: float = 1e-12) -> np.ndarray:
    """``torch.nn.functional.normalize(x, p=2, dim=-1)``: x / max(|x|, eps)
    (src/stage2_rescorer.py:173-174,188-189)."""
    x = np.asarray(x, dtype=np.float32)
    n = np.sqrt((x.astype(np.float32) ** 2).sum(axis=-1, keepdims=True, dtype=np.float32))
    return x / np.maximum(n, np.float32(eps))


def sim_matrix(q_tok: np.ndarray, d_tok: np.ndarray, normalize: bool = True) -> np.ndarray:
    """Cosine similarity matrix [Lq, Ld] (src/stage2_rescorer.py:173-177)."""
    q = np.asarray(q_tok, np.float32).reshape(-1, q_tok.shape[-1])
    d = np.asarray(d_tok, np.float32).reshape(-1, d_tok.shape[-1])
    if normalize:
        q, d = l2_normalize_tokens(q), l2_normalize_tokens(d)
    return q @ d.T


def maxsim_score(q_tok, d_tok, normalize: bool = True) -> float:
    """mean_i max_j cos(q_i, d_j) -- MEAN over query tokens, not the sum of
    canonical ColBERT (src/stage2_rescorer.py:180-183)."""
    m = sim_matrix(q_tok, d_tok, normalize).max(axis=-1)
    return float(m.mean(dtype=np.float32))


def colbert_score(q_tok, d_tok, normalize: bool = True) -> float:
    """sum_i softmax(m)_i * m_i with m_i = max_j cos(q_i, d_j)
    (src/stage2_rescorer.py:195-199)."""
    m = sim_matrix(q_tok, d_tok, normalize).max(axis=-1).astype(np.float32)
    e = np.exp(m - m.max())
    w = e / e.sum(dtype=np.float32)
    return float((m * w).sum(dtype=np.float32))


def score(q_tok, d_tok, mode: int = MODE_MAXSIM, normalize: bool = True) -> float:
    return maxsim_score(q_tok, d_tok, normalize) if mode == MODE_MAXSIM else colbert_score(q_tok, d_tok, normalize)


def score_candidates(q_tok, doc_tok_list, mode: int = MODE_MAXSIM, normalize: bool = True) -> np.ndarray:
    """One score per candidate, the per-candidate loop of
    ``rescore_candidates`` (src/stage2_rescorer.py:268-273)."""
    return np.array([score(q_tok, d, mode, normalize) for d in doc_tok_list], dtype=np.float32)


def rescore_order(scores, top_k: int) -> np.ndarray:
    """Indices after the reference's stable descending sort + truncate
    (src/stage2_rescorer.py:294-297): equal scores keep incoming order."""
    scores = np.asarray(scores)
    order = sorted(range(len(scores)), key=lambda i: scores[i], reverse=True)
    return np.array(order[:top_k], dtype=np.int64)
