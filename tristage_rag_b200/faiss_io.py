"""Migration aid: read the FAISS index file the reference writes with ``faiss.write_index``
(``/root/reference/src/stage1_retriever.py:436``, ``stage1_faiss.index``) so an existing deployment can
move its vectors -- and, for the approximate index, its centroids and lists -- into a ``ts_index`` shard
without re-encoding the corpus.

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

The reference's other index type (``IndexIVFFlat`` over an ``IndexFlatIP`` quantizer, built when the first
batch has more than 1000 rows, ``:262-273``) is read by ``read_faiss_ivf``: ``write_index`` -> ``IndexIVFFlat``
branch, ``write_ivf_header``, ``write_direct_map``, ``write_InvertedLists`` (``ArrayInvertedLists``):

    uint32  fourcc            "IwFl"
    int32 d | int64 ntotal | int64 dummy x 2 | uint8 is_trained | int32 metric_type      (index header)
    uint64  nlist
    uint64  nprobe
    <flat index>              the coarse quantizer: a complete "IxFI" index as above, ntotal == nlist
    uint8   direct_map_type   0 = none
    uint64  n, int64[n]       direct map array (empty for type 0)
    uint32  fourcc            "ilar"
    uint64  nlist
    uint64  code_size         == 4 * d
    uint32  fourcc            "full": uint64 nlist, uint64 size[nlist]
                              "sprs": uint64 2m, then m (list, size) uint64 pairs for the non-empty lists
    per non-empty list, in list order: float32 codes[size * d], int64 ids[size]

The lists store the row ids, which for the reference are the insertion positions (``add`` without ids), so the
vectors can be put back in insertion order and every row's list is known: the importer returns both.
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
            raise FaissFormatError(f"{path}: this is an IndexIVFFlat file; use read_faiss_ivf")
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


def faiss_fourcc(path: str) -> bytes:
    with open(path, "rb") as f:
        return f.read(4)


class _Cursor:
    def __init__(self, path: str, buf: memoryview):
        self.path, self.buf, self.pos = path, buf, 0

    def take(self, fmt: str):
        n = struct.calcsize(fmt)
        if self.pos + n > len(self.buf):
            raise FaissFormatError(f"{self.path}: truncated at byte {self.pos}")
        out = struct.unpack_from(fmt, self.buf, self.pos)
        self.pos += n
        return out

    def array(self, dtype: str, count: int) -> np.ndarray:
        n = count * np.dtype(dtype).itemsize
        if count < 0 or self.pos + n > len(self.buf):
            raise FaissFormatError(f"{self.path}: truncated at byte {self.pos} (wanted {count} x {dtype})")
        out = np.frombuffer(self.buf, dtype=dtype, count=count, offset=self.pos)
        self.pos += n
        return out


def _index_header(cur: _Cursor):
    d, ntotal, _d0, _d1, trained, metric = cur.take("<iqqqBi")
    if d <= 0 or d > 65536 or ntotal < 0 or trained not in (0, 1) or metric not in (0, 1):
        raise FaissFormatError(f"{cur.path}: implausible header (d={d}, ntotal={ntotal}, metric_type={metric})")
    return d, ntotal, trained, metric


def read_faiss_ivf(path: str):
    """-> dict(vectors float32 [ntotal, d] in insertion order, assign int32 [ntotal], centroids float32
    [nlist, d], nlist, nprobe, metric).  Raises FaissFormatError unless the file is a self-consistent
    IndexIVFFlat over a flat inner-product quantizer whose ids are the insertion positions."""
    with open(path, "rb") as f:
        raw = f.read()
    cur = _Cursor(path, memoryview(raw))
    (fourcc,) = cur.take("<4s")
    if fourcc != b"IwFl":
        raise FaissFormatError(f"{path}: fourcc {fourcc!r} is not an IndexIVFFlat file")
    d, ntotal, trained, metric = _index_header(cur)
    nlist, nprobe = cur.take("<QQ")
    if not trained or nlist < 1 or nlist > (1 << 24) or nprobe < 1:
        raise FaissFormatError(f"{path}: implausible IVF header (trained={trained}, nlist={nlist}, nprobe={nprobe})")
    (qcc,) = cur.take("<4s")
    if qcc not in FLAT_FOURCC:
        raise FaissFormatError(f"{path}: coarse quantizer {qcc!r} is not a flat index")
    qd, qn, _qt, qmetric = _index_header(cur)
    (n_floats,) = cur.take("<Q")
    if qd != d or qn != nlist or n_floats != nlist * d:
        raise FaissFormatError(f"{path}: quantizer ({qn} x {qd}, {n_floats} floats) does not match nlist={nlist}, d={d}")
    centroids = cur.array("<f4", nlist * d).reshape(nlist, d).copy()
    (dm_type,) = cur.take("<B")
    (dm_n,) = cur.take("<Q")
    if dm_type not in (0, 1) or dm_n not in (0, ntotal):
        raise FaissFormatError(f"{path}: unsupported direct map (type {dm_type}, {dm_n} entries)")
    cur.array("<i8", dm_n)
    (ilcc,) = cur.take("<4s")
    if ilcc != b"ilar":
        raise FaissFormatError(f"{path}: inverted lists {ilcc!r} are not ArrayInvertedLists")
    il_nlist, code_size = cur.take("<QQ")
    if il_nlist != nlist or code_size != 4 * d:
        raise FaissFormatError(f"{path}: inverted lists (nlist={il_nlist}, code_size={code_size}) do not match the header")
    (kind,) = cur.take("<4s")
    sizes = np.zeros(nlist, np.int64)
    (nsz,) = cur.take("<Q")
    if kind == b"full":
        if nsz != nlist:
            raise FaissFormatError(f"{path}: {nsz} list sizes for {nlist} lists")
        sizes[:] = cur.array("<u8", nlist)
    elif kind == b"sprs":
        if nsz % 2:
            raise FaissFormatError(f"{path}: odd sparse size table")
        pairs = cur.array("<u8", nsz).reshape(-1, 2)
        if len(pairs) and (pairs[:, 0].max() >= nlist or len(np.unique(pairs[:, 0])) != len(pairs)):
            raise FaissFormatError(f"{path}: bad sparse size table")
        sizes[pairs[:, 0].astype(np.int64)] = pairs[:, 1]
    else:
        raise FaissFormatError(f"{path}: list size table {kind!r} is neither full nor sparse")
    if int(sizes.sum()) != ntotal:
        raise FaissFormatError(f"{path}: lists hold {int(sizes.sum())} rows, header says {ntotal}")
    vectors = np.empty((ntotal, d), np.float32)
    assign = np.full(ntotal, -1, np.int32)
    for l in range(nlist):
        n = int(sizes[l])
        if n == 0:
            continue
        codes = cur.array("<f4", n * d).reshape(n, d)
        ids = cur.array("<i8", n)
        if ids.min() < 0 or ids.max() >= ntotal or (assign[ids] != -1).any():
            raise FaissFormatError(f"{path}: list {l} holds ids that are not insertion positions")
        vectors[ids] = codes
        assign[ids] = l
    if cur.pos != len(raw) or (assign < 0).any():
        raise FaissFormatError(f"{path}: {len(raw) - cur.pos} trailing bytes / unassigned rows")
    kind_q = FLAT_FOURCC[qcc] or ("ip" if qmetric == 0 else "l2")
    if metric != 0 or kind_q != "ip":
        raise FaissFormatError(f"{path}: the reference builds inner-product IVF indexes; this one is not")
    return {"vectors": vectors, "assign": assign, "centroids": centroids, "nlist": int(nlist), "nprobe": int(nprobe),
            "metric": "ip"}
