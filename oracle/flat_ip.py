"""Stage-1 oracle: exact inner-product top-k (restated ``faiss.IndexFlatIP``).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  "parity unpinned" for the
FAISS arithmetic itself (faiss-cpu>=1.7.0 is an un-vendored, unpinned pip
dependency of the reference, ``requirements.txt:10``); the call sites this
follows are ``src/stage1_retriever.py:276-277`` (``IndexFlatIP(d)``, ``add``),
``:313`` (``add``) and ``:380`` (``search``).

Tie rule: FAISS's order among exactly-equal scores is implementation defined
(heap for k<100, reservoir for k>=100).  The oracle is deterministic instead:
score descending, then id ascending.  The parity checker (``check_topk``)
treats swaps inside a relative-score band as ties, so the rule only matters
for exact duplicates.
"""
from __future__ import annotations

import numpy as np

# FAISS fills unused result slots with label -1 and the "neutral" element of
# its max-similarity heap, std::numeric_limits<float>::lowest().
LOWEST_F32 = np.float32(-3.4028234663852886e38)


def normalize_rows(x: np.ndarray) -> np.ndarray:
    """``Stage1Retriever._normalize_embeddings`` (src/stage1_retriever.py:285-288).

    eps is ADDED to the norm (not a clamp); dtype follows numpy promotion, i.e.
    fp32 in -> fp32 out.
    """
    norms = np.linalg.norm(x, axis=1, keepdims=True)
    return x / (norms + 1e-8)


def topk_desc(scores: np.ndarray, k: int, ids: np.ndarray | None = None):
    """Per-row top-k of ``scores[B, N]``: score desc, id asc.  Pads with
    (LOWEST_F32, -1) when k > N.  Returns (D[B,k] f32, I[B,k] i64)."""
    scores = np.asarray(scores, dtype=np.float32)
    B, N = scores.shape
    D = np.full((B, k), LOWEST_F32, dtype=np.float32)
    I = np.full((B, k), -1, dtype=np.int64)
    kk = min(k, N)
    if kk == 0:
        return D, I
    for b in range(B):
        s = scores[b]
        if kk < N:
            # everything >= the kk-th largest value, then an exact ordered cut
            kth = np.partition(s, N - kk)[N - kk]
            cand = np.nonzero(s >= kth)[0]
        else:
            cand = np.arange(N)
        cid = cand if ids is None else ids[cand]
        order = np.lexsort((cid, -s[cand].astype(np.float64)))[:kk]
        D[b, :kk] = s[cand][order]
        I[b, :kk] = cid[order]
    return D, I


class IndexFlatIP:
    """Restated ``faiss.IndexFlatIP``: flat fp32 storage, exact IP search."""

    def __init__(self, d: int):
        self.d = int(d)
        self.ntotal = 0
        self._chunks: list[np.ndarray] = []
        self._x: np.ndarray | None = None
        self.is_trained = True

    def add(self, x: np.ndarray) -> None:
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d, (x.shape, self.d)
        self._chunks.append(x.copy())
        self._x = None
        self.ntotal += x.shape[0]

    @property
    def xb(self) -> np.ndarray:
        if self._x is None:
            self._x = (np.concatenate(self._chunks, axis=0) if self._chunks
                       else np.zeros((0, self.d), np.float32))
            self._chunks = [self._x]
        return self._x

    def search(self, q: np.ndarray, k: int, block: int = 262144):
        """D[B,k] fp32 descending, I[B,k] int64, ``-1`` beyond ntotal."""
        q = np.ascontiguousarray(q, dtype=np.float32)
        assert q.ndim == 2 and q.shape[1] == self.d
        x = self.xb
        B = q.shape[0]
        if self.ntotal <= block:
            return topk_desc(q @ x.T, k)
        # blocked scan so a [B, N] score matrix is never materialised
        return search_streamed(((s, x[s:min(s + block, self.ntotal)]) for s in range(0, self.ntotal, block)), q, k)


def search_streamed(blocks, q: np.ndarray, k: int):
    """The same exact search over a corpus that arrives as ``(first_row, rows[n, d])`` blocks (a corpus too
    large for host memory is fetched back from the device shard by shard): per-block exact top-k, merged
    with the deterministic order (score desc, id asc).  Returns (D[B,k], I[B,k]) padded like FAISS."""
    q = np.ascontiguousarray(q, dtype=np.float32)
    B = q.shape[0]
    bestD = np.full((B, 0), LOWEST_F32, np.float32)
    bestI = np.full((B, 0), -1, np.int64)
    for s, xb in blocks:
        xb = np.asarray(xb, dtype=np.float32)
        D, I = topk_desc(q @ xb.T, min(k, xb.shape[0]))
        I = I + s
        catD = np.concatenate([bestD, D], axis=1)
        catI = np.concatenate([bestI, I], axis=1)
        newD = np.empty((B, min(k, catD.shape[1])), np.float32)
        newI = np.empty_like(newD, dtype=np.int64)
        for b in range(B):
            o = np.lexsort((catI[b], -catD[b].astype(np.float64)))[: newD.shape[1]]
            newD[b], newI[b] = catD[b][o], catI[b][o]
        bestD, bestI = newD, newI
    D = np.full((B, k), LOWEST_F32, np.float32)
    I = np.full((B, k), -1, np.int64)
    D[:, : bestD.shape[1]] = bestD
    I[:, : bestI.shape[1]] = bestI
    return D, I


def round_to(x: np.ndarray, dtype: str) -> np.ndarray:
    """Round fp32 values to the GPU storage dtype and back to fp32, so the
    oracle scores the SAME stored values the kernels see."""
    import torch

    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    if dtype in ("bf16", "bfloat16"):
        return t.to(torch.bfloat16).to(torch.float32).numpy()
    if dtype in ("fp16", "f16", "float16"):
        return t.to(torch.float16).to(torch.float32).numpy()
    if dtype in ("fp32", "f32", "float32"):
        return t.numpy().copy()
    if dtype == "tf32":
        # what the tensor path sees of fp32 storage (kind::tf32): 10 mantissa bits, round to nearest even
        u = t.numpy().view(np.uint32).astype(np.uint64)
        r = ((u + 0xFFF + ((u >> 13) & 1)) & 0xFFFFE000).astype(np.uint32)
        return r.view(np.float32).copy()
    raise ValueError(dtype)


def check_topk(got_D, got_I, ref_scores_of, ref_D, ref_I, rel=1e-3):
    """Parity rule of BASELINE.json north_star / SURVEY.md §8d.

    ids must equal the oracle's except for swaps inside a near-tie band of
    ``rel`` relative score; scores must agree within ``rel`` relative.

    ``ref_scores_of(b, ids) -> fp32 oracle scores`` of arbitrary ids for query b
    (needed to judge ids we returned that the oracle did not).
    Returns a list of human-readable violations (empty = pass).
    """
    bad = []
    got_D = np.asarray(got_D)
    got_I = np.asarray(got_I)
    B, k = ref_I.shape
    if got_I.shape != ref_I.shape:
        return [f"shape {got_I.shape} != {ref_I.shape}"]
    for b in range(B):
        valid = ref_I[b] >= 0
        nv = int(valid.sum())
        if (got_I[b, nv:] != -1).any():
            bad.append(f"q{b}: expected -1 padding beyond {nv}")
        g_ids, r_ids = got_I[b, :nv], ref_I[b, :nv]
        if nv == 0:
            continue
        if len(set(g_ids.tolist())) != nv:
            bad.append(f"q{b}: duplicate ids returned")
        if (g_ids < 0).any():
            bad.append(f"q{b}: negative id among valid slots")
            continue
        kth = float(ref_D[b, nv - 1])
        band = rel * max(abs(kth), 1e-30)
        # 1. scores we report match the oracle's score of the same id
        o = ref_scores_of(b, g_ids)
        err = np.abs(got_D[b, :nv].astype(np.float64) - o) / np.maximum(np.abs(o), 1e-30)
        if (err > rel).any():
            j = int(err.argmax())
            bad.append(f"q{b}: score of id {g_ids[j]} off by {err[j]:.2e} rel "
                       f"(got {got_D[b, j]}, oracle {o[j]})")
        # 2. every id we return that the oracle does not must sit in the
        #    boundary band; every oracle id we miss likewise
        extra = np.setdiff1d(g_ids, r_ids)
        missing = np.setdiff1d(r_ids, g_ids)
        if len(extra):
            so = ref_scores_of(b, extra)
            if (so < kth - band).any():
                bad.append(f"q{b}: returned ids {extra[so < kth - band][:5]} below the k-th band")
        if len(missing):
            sm = ref_scores_of(b, missing)
            if (sm > kth + band).any():
                bad.append(f"q{b}: missed ids {missing[sm > kth + band][:5]} above the k-th band")
        # 3. order: descending by OUR scores, and consistent with the oracle's
        #    scores up to the tie band
        if (np.diff(got_D[b, :nv]) > 0).any():
            bad.append(f"q{b}: scores not descending")
        inv = o[:-1] - o[1:]
        tol = rel * np.maximum(np.abs(o[:-1]), 1e-30)
        if (inv < -tol).any():
            j = int((inv + tol).argmin())
            bad.append(f"q{b}: order inversion beyond tie band at rank {j}")
    return bad
