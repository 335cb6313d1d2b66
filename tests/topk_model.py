"""Executable model (pure Python) of the exact top-k selection fused into the
Stage-1 tensor-path scan (tristage_rag_b200/csrc/s1_umma.cu + topk_select.cu).

It restates, for ONE query, what the epilogue threads of all CTAs do with the
scores of their tiles -- chunk-max fast path, per-(slice) candidate list with
overflow prune, the J best appended scores, publication of the J-th best,
the shared bound min_c pub[c] (optionally stale), the final filter of the
select kernel -- so the ALGORITHM can be property-tested on the CPU against the
oracle on inputs the GPU suites do not reach (adversarial orders, mass ties).
Test infrastructure only.
"""
from __future__ import annotations

import math
import random

import numpy as np

TILE, CHUNK = 256, 32


def nextbelow(x: float) -> float:
    return float(np.nextafter(np.float32(x), np.float32(-np.inf)))


def key(score: float, idx: int):
    """sort key: score descending, id ascending (device: ord(score)<<32 | ~idx)"""
    return (-float(score), idx)


def model_topk(scores: np.ndarray, k: int, n_slices: int, cap: int | None = None, stale: int = 0, seed: int = 0):
    """scores: fp32 [N] of one query.  Returns the list of (score, id) the device would output.
    stale: the shared bound seen by a slice may lag by up to `stale` tile rounds."""
    scores = np.asarray(scores, np.float32)
    N = len(scores)
    rng = random.Random(seed)
    cap = cap or (256 if k <= 128 else 1024)
    n_tiles = (N + TILE - 1) // TILE
    n_slices = max(1, min(n_slices, n_tiles))
    j = (k + n_slices - 1) // n_slices
    J = j if j <= 8 else 0
    NEG = -math.inf
    pub = [NEG] * n_slices
    # pre-pass: J-th best of the first tile of every slice
    if J:
        for s in range(n_slices):
            v = sorted(scores[s * TILE:(s + 1) * TILE].tolist(), reverse=True)
            pub[s] = v[J - 1] if len(v) >= J else NEG
    history = [min(pub) if J else NEG]                 # shared bound after each tile round
    st = [dict(tau=nextbelow(history[0]) if J else NEG, tj=[], pub_last=pub[s], lst=[]) for s in range(n_slices)]
    rounds = (n_tiles + n_slices - 1) // n_slices
    stats = dict(appends=0, prunes=0)
    for it in range(rounds):
        for s in range(n_slices):
            t = s + it * n_slices
            if t >= n_tiles:
                continue
            S = st[s]
            base = t * TILE
            for c0 in range(base, min(base + TILE, N), CHUNK):
                vals = scores[c0:min(c0 + CHUNK, N)]
                if float(vals.max()) > S["tau"]:
                    thr = S["tau"]
                    for off, v in enumerate(vals.tolist()):
                        if v > thr:
                            S["lst"].append(key(v, c0 + off))
                            stats["appends"] += 1
                            if J and (len(S["tj"]) < J or v > S["tj"][J - 1]):
                                S["tj"] = sorted(S["tj"] + [v], reverse=True)[:8]
                if len(S["lst"]) > cap - 32:
                    S["lst"] = sorted(S["lst"])[:k]
                    stats["prunes"] += 1
                    if len(S["lst"]) >= k:
                        S["tau"] = max(S["tau"], -S["lst"][k - 1][0])
            if J:
                tjJ = S["tj"][J - 1] if len(S["tj"]) >= J else NEG
                if tjJ > S["pub_last"]:
                    pub[s] = S["pub_last"] = tjJ
                seen = history[max(0, len(history) - 1 - rng.randint(0, stale))]
                S["tau"] = max(S["tau"], nextbelow(seen))
        history.append(min(pub) if J else NEG)
    # select kernel: final filter with min_c pub[c], then sort
    allk = [e for S in st for e in S["lst"]]
    if J:
        m = min(pub)
        allk = [e for e in allk if -e[0] >= m]
    out = sorted(set(allk))[:k]
    return [(-a, b) for a, b in out], stats
