"""Stage-1 oracle, approximate mode: restated ``faiss.IndexIVFFlat`` over an ``IndexFlatIP`` coarse
quantizer, the index the reference builds when its first batch has more than 1000 rows
(``src/stage1_retriever.py:262-273``: ``IndexIVFFlat(quantizer, d, nlist, METRIC_INNER_PRODUCT)``,
``train``, ``add``, ``nprobe = config.nprobe``; later batches ``add`` at ``:313``; ``search`` at ``:380``).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  "parity unpinned": FAISS is an un-vendored,
unpinned dependency (``requirements.txt:10``) that cannot be installed here, so what follows restates
its published algorithm and cannot be checked against its output:

* ``train``: k-means (``faiss::Clustering`` as ``Level1Quantizer::train_q1`` runs it: 10 iterations,
  at most 256 training points per centroid, points assigned with the QUANTIZER -- i.e. by largest
  inner product -- centroids = plain means, an empty cluster is re-seeded by splitting a populated
  one with a +-1/1024 perturbation).  The random choices use numpy's generator, not FAISS's, so the
  centroids are a k-means solution of the same kind, not FAISS's centroids; a deployment that needs
  the reference's own lists imports them from its index file (``tristage_rag_b200/faiss_io.py``).
* ``add``: each row goes to the list of the centroid with the largest inner product.
* ``search``: the ``nprobe`` lists with the largest <q, centroid>, an exact inner-product scan of
  their rows, top-k by descending score, ``-1`` / lowest-float padding when fewer than k rows were
  scanned.  Tie rule as in ``flat_ip.py``: score descending, then id ascending (lists: lowest list
  number first).
"""
from __future__ import annotations

import numpy as np

from .flat_ip import LOWEST_F32, topk_desc

EPS_SPLIT = 1.0 / 1024.0


def assign_lists(x: np.ndarray, centroids: np.ndarray) -> np.ndarray:
    """List of every row: argmax_c <x, centroid_c> in float64 (ties -> lowest list)."""
    s = np.asarray(x, np.float64) @ np.asarray(centroids, np.float64).T
    return np.argmax(s, axis=1).astype(np.int32)


def assign_margin(x: np.ndarray, centroids: np.ndarray) -> np.ndarray:
    """Best minus second-best centroid score per row (how safe each assignment is against rounding)."""
    s = np.asarray(x, np.float64) @ np.asarray(centroids, np.float64).T
    if s.shape[1] < 2:
        return np.full(s.shape[0], np.inf)
    part = np.partition(s, s.shape[1] - 2, axis=1)
    return part[:, -1] - part[:, -2]


def coarse_probe(q32: np.ndarray, centroids: np.ndarray, nprobe: int):
    """quantizer.search(q, nprobe): ([B, nprobe] list numbers, [B, nprobe] fp32 scores), best first."""
    s = (np.asarray(q32, np.float64) @ np.asarray(centroids, np.float64).T).astype(np.float32)
    D, I = topk_desc(s, min(nprobe, centroids.shape[0]))
    return I.astype(np.int32), D


def ivf_search(x_stored: np.ndarray, q_stored: np.ndarray, assign: np.ndarray, probes: np.ndarray, k: int):
    """Exact top-k of every query over the rows of its probed lists.  ``x_stored`` / ``q_stored`` are
    the values the kernels see (already rounded to the storage dtype, fp32 containers)."""
    B = q_stored.shape[0]
    D = np.full((B, k), LOWEST_F32, np.float32)
    I = np.full((B, k), -1, np.int64)
    for b in range(B):
        lists = [int(l) for l in probes[b] if l >= 0]
        rows = np.nonzero(np.isin(assign, lists))[0]
        if rows.size == 0:
            continue
        s = (x_stored[rows].astype(np.float32) @ q_stored[b].astype(np.float32))[None, :]
        d, i = topk_desc(s, min(k, rows.size), ids=rows.astype(np.int64))
        D[b, : d.shape[1]], I[b, : i.shape[1]] = d[0], i[0]
    return D, I


def kmeans_ip(x: np.ndarray, nlist: int, niter: int = 10, seed: int = 1234, max_points_per_centroid: int = 256):
    """The training loop described in the module docstring, written the slow, obvious way."""
    x = np.asarray(x, np.float32)
    n, d = x.shape
    assert n >= nlist >= 1
    rng = np.random.default_rng(seed)
    if n > nlist * max_points_per_centroid:
        x = x[np.sort(rng.permutation(n)[: nlist * max_points_per_centroid])]
        n = x.shape[0]
    cent = x[np.sort(rng.permutation(n)[:nlist])].astype(np.float32).copy()
    for _ in range(niter):
        a = assign_lists(x, cent)
        cnt = np.bincount(a, minlength=nlist).astype(np.float64)
        new = np.zeros((nlist, d), np.float64)
        for i in range(n):
            new[a[i]] += x[i]
        for c in range(nlist):
            if cnt[c] > 0:
                new[c] /= cnt[c]
        for c in range(nlist):                     # re-seed empty clusters from the largest one
            if cnt[c] == 0:
                big = int(np.argmax(cnt))
                new[c] = new[big]
                sign = np.where(np.arange(d) % 2 == 0, 1.0, -1.0)
                new[c] *= 1.0 + sign * EPS_SPLIT
                new[big] *= 1.0 - sign * EPS_SPLIT
                cnt[c] = cnt[big] / 2
                cnt[big] -= cnt[c]
        cent = new.astype(np.float32)
    return cent
