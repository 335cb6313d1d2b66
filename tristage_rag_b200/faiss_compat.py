"""The part of the ``faiss`` module surface the reference touches, backed by libtristage.

The reference's Stage 1 binds to exactly these names (``/root/reference/src/stage1_retriever.py``):
``faiss.IndexFlatIP(d)`` (:263,276), ``faiss.IndexIVFFlat(quantizer, d, nlist, faiss.METRIC_INNER_PRODUCT)``
(:264), ``index.train / add / search / nprobe / ntotal`` (:267,270,273,277,313,380), ``faiss.write_index``
(:436) and ``faiss.read_index`` (:463).  Installing this module under the name ``faiss`` lets the reference's
OWN ``src/stage1_retriever.py`` run unmodified on the B200 kernels -- the thinnest drop-in there is:

    import sys, tristage_rag_b200.faiss_compat as faiss_compat
    sys.modules["faiss"] = faiss_compat            # before `import src.stage1_retriever`

(``INTEGRATION.md`` way C; ``tools/faiss_shim_check.py`` runs it against the reference's own output.)
The corpus storage dtype comes from ``TS_STORAGE_DTYPE`` (bf16 | fp16 | fp32, default bf16), the GPU from
``TS_GPU_INDEX``.  Indexes are exact unless the caller itself builds an ``IndexIVFFlat`` -- which the
reference does for first batches of more than 1000 rows.
"""
from __future__ import annotations

import os

from . import stage1_retriever as _s1

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1


def _storage() -> str:
    return os.environ.get("TS_STORAGE_DTYPE", "bf16")


def _gpu() -> int:
    return int(os.environ.get("TS_GPU_INDEX", "0"))


class IndexFlatIP(_s1.IndexFlatIP):
    def __init__(self, d: int):
        super().__init__(int(d), _storage(), _gpu())


class IndexIVFFlat(_s1.IndexIVFFlat):
    def __init__(self, quantizer, d: int, nlist: int, metric: int = METRIC_INNER_PRODUCT):
        if metric != METRIC_INNER_PRODUCT:
            raise ValueError("only METRIC_INNER_PRODUCT is supported (the metric the reference uses)")
        if not isinstance(quantizer, _s1.IndexFlatIP) or quantizer.d != int(d):
            raise ValueError("the coarse quantizer must be an IndexFlatIP of the same dimension (what the reference passes)")
        super().__init__(int(d), int(nlist), _storage(), _gpu())
        self.quantizer = quantizer                 # kept for callers that read it back; centroids live in the ts_ivf


def write_index(index, path: str) -> None:
    index.save(path)


def read_index(path: str):
    """A shard file written by ``write_index`` above, or a file FAISS itself wrote (flat or IVF-flat, imported)."""
    import numpy as np

    from . import _lib
    from .faiss_io import faiss_fourcc, read_faiss_flat, read_faiss_ivf

    with open(path, "rb") as f:
        magic = f.read(8)
    if magic == b"TSSHARD2":
        kind = _s1.IndexIVFFlat if os.path.exists(path + ".ivf.npz") else _s1.IndexFlatIP
        return kind.load(path, _storage(), _gpu())
    if faiss_fourcc(path) == b"IwFl":
        parts = read_faiss_ivf(path)
        base = _lib.Index(parts["vectors"].shape[1], _storage(), "ip", _gpu())
        base.add(parts["vectors"], normalize=False)
        return _s1.IndexIVFFlat.from_parts(base, parts["centroids"], parts["assign"], parts["nprobe"], _storage(), _gpu())
    x, metric = read_faiss_flat(path)
    if metric != "ip":
        raise ValueError(f"{path}: the reference uses inner-product indexes, this one is {metric}")
    idx = IndexFlatIP(x.shape[1])
    idx.add(np.ascontiguousarray(x, np.float32))
    return idx
