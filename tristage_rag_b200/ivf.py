"""Training of the coarse quantizer for the approximate Stage-1 mode (``_lib.IVF`` / ``csrc/ivf.cu``).

Stands in for ``self.faiss_index.train(embeddings)`` (``/root/reference/src/stage1_retriever.py:267``).
FAISS runs this k-means on the host for the reference and so do we: the work is bounded by the
training-set cap (at most 256 points per centroid, 25 600 x dim at the reference's ``nlist = 100``) and
happens once per index; the per-row work of ``add`` and ``search`` is on the device.

The algorithm is the one ``faiss::Clustering`` runs for an ``IndexIVFFlat`` with an inner-product
quantizer: random initial centroids taken from the data, ``niter`` (10) rounds of
"assign every point to the centroid with the largest inner product, move each centroid to the mean of its
points", empty clusters re-seeded by splitting the largest cluster with a +-1/1024 perturbation.  The
random choices come from numpy's generator, so the centroids are not bit-identical to FAISS's; a
deployment that needs the reference's own lists imports them from its index file
(``faiss_io.read_faiss_ivf``).
"""
from __future__ import annotations

import numpy as np

NITER = 10
MAX_POINTS_PER_CENTROID = 256
MIN_ROWS_FOR_IVF = 1000          # the reference switches index type above this many first-batch rows (:262)
_EPS_SPLIT = 1.0 / 1024.0


def train_centroids(x: np.ndarray, nlist: int, niter: int = NITER, seed: int = 1234,
                    max_points_per_centroid: int = MAX_POINTS_PER_CENTROID) -> np.ndarray:
    """x [n, dim] fp32 (n >= nlist) -> centroids [nlist, dim] fp32."""
    x = np.ascontiguousarray(x, np.float32)
    n, d = x.shape
    if n < nlist:
        raise ValueError(f"cannot train {nlist} lists from {n} rows")
    rng = np.random.default_rng(seed)
    if n > nlist * max_points_per_centroid:
        x = x[np.sort(rng.permutation(n)[: nlist * max_points_per_centroid])]
        n = x.shape[0]
    cent = x[np.sort(rng.permutation(n)[:nlist])].copy()
    x64 = x.astype(np.float64)
    sign = np.where(np.arange(d) % 2 == 0, 1.0, -1.0)
    for _ in range(niter):
        a = np.argmax(x64 @ cent.astype(np.float64).T, axis=1)
        cnt = np.bincount(a, minlength=nlist).astype(np.float64)
        new = np.zeros((nlist, d), np.float64)
        order = np.argsort(a, kind="stable")                       # members of a cluster in row order, clusters back to back
        nz = cnt > 0
        starts = np.concatenate([[0], np.cumsum(cnt.astype(np.int64))])[:-1][nz]
        new[nz] = np.add.reduceat(x64[order], starts, axis=0)      # row-by-row sums, the order of the obvious loop
        new[nz] /= cnt[nz, None]
        for c in np.nonzero(~nz)[0]:
            big = int(np.argmax(cnt))
            new[c] = new[big] * (1.0 + sign * _EPS_SPLIT)
            new[big] = new[big] * (1.0 - sign * _EPS_SPLIT)
            cnt[c] = cnt[big] / 2
            cnt[big] -= cnt[c]
        cent = new.astype(np.float32)
    return cent
