"""Migration aid: read the flat FAISS index file the reference writes with ``faiss.write_index``
(``/root/reference/src/stage1_retriever.py:436``, ``stage1_faiss.index``) so an existing deployment can
move its vectors into a ``ts_index`` shard without re-encoding the corpus.

FAISS is not installable in this environment, so the layout below is restated from FAISS's public
``impl/index_write.cpp`` (``write_index`` -> ``IndexFlat`` branch, ``write_index_header``,
``WRITEXBVECTOR``) and could not be checked against a file written by FAISS here.  The reader therefore
accepts a file ONLY if every field is self-consistent (fourcc, dimensions, element count, exact file
size); anything else raises, it never guesses.

    uint32  fourcc            "IxFI" (IndexFlatIP) | "IxF2" (IndexFlatL2) | "IxFl" (IndexFlat)
    int32   d
    int64   ntotal
    int64   dummy, dummy      (1 << 20)
    uint8   is_trained
    int32   metric_type       0 = inner product, 1 = L2      (+ float32 metric_arg when metric_type > 1)
    uint64  n_floats          == ntotal * d   (the codes vector, written in units of 4 bytes)
    float32 xb[ntotal * d]

Only flat indexes are supported: the reference's IVF branch (``IndexIVFFlat``, fourcc "IwFl") is an
approximate index whose lists cannot be turned back into the insertion order the doc ids rely on.
"""
from __future__ import annotations

import os
import struct

import numpy as np

FLAT_FOURCC = {b"IxFI": "ip", b"IxF2": "l2", b"IxFl": None}


class FaissFormatError(ValueError):
    pass


def read_faiss_flat(path: str):
    """-> (vectors float32 [ntotal, d], metric "ip" | "l2").  Raises FaissFormatError unless the file is a
    self-consistent flat FAISS index."""
    size = os.path.getsize(path)
    with open(path, "rb") as f:
        head = f.read(4 + 4 + 8 + 8 + 8 + 1 + 4)
        if len(head) < 37:
            raise FaissFormatError(f"{path}: too short for a FAISS index header")
        fourcc = head[:4]
        if fourcc == b"IwFl":
            raise FaissFormatError(f"{path}: IndexIVFFlat (approximate) files cannot be migrated; rebuild from the documents")
        if fourcc not in FLAT_FOURCC:
            raise FaissFormatError(f"{path}: fourcc {fourcc!r} is not a flat FAISS index")
        d, ntotal, dummy0, dummy1, trained, metric = struct.unpack("<iqqqBi", head[4:])
        if d <= 0 or d > 65536 or ntotal < 0 or trained not in (0, 1) or metric not in (0, 1):
            raise FaissFormatError(f"{path}: implausible header (d={d}, ntotal={ntotal}, metric_type={metric})")
        (n_floats,) = struct.unpack("<Q", f.read(8))
        if n_floats != ntotal * d or size != 37 + 8 + n_floats * 4:
            raise FaissFormatError(f"{path}: vector block ({n_floats} floats, file {size} bytes) does not match "
                                   f"ntotal={ntotal} x d={d}")
        x = np.fromfile(f, dtype="<f4", count=n_floats).reshape(ntotal, d)
    kind = FLAT_FOURCC[fourcc] or ("ip" if metric == 0 else "l2")
    if (kind == "ip") != (metric == 0):
        raise FaissFormatError(f"{path}: fourcc {fourcc!r} and metric_type {metric} disagree")
    return x, kind
