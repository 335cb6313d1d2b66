"""ctypes binding of ``lib/libtristage.so`` (the C ABI in ``include/tristage.h``).

There is no CPU fallback anywhere in this package: if the shared library is
missing, or no sm_100 device is usable, the calls below raise.  The library is
built in-tree by ``tristage_rag_b200/csrc/Makefile`` (``build()`` here, or
``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libtristage.so")
CSRC = os.path.join(_HERE, "csrc")

TS_F32, TS_BF16, TS_F16 = 0, 1, 2
TS_METRIC_IP, TS_METRIC_COSINE = 0, 1
TS_PATH_AUTO, TS_PATH_STREAM, TS_PATH_UMMA = 0, 1, 2
TS_FLAG_NORMALIZE_Q = 1
TS_S2_MAXSIM, TS_S2_COLBERT = 0, 1
TS_S2_FORCE_SIMT = 0x100          # mode bit: take the CUDA-core Stage-2 kernel
TS_MAX_K = 512
TS_BM25_MAX_K = 1024
TS_FUSE_MAX = 2048
TS_S2_MAX_LQ = 128
TS_S2_MAX_LD = 256
TS_ERR_EMPTY = -6

DTYPES = {"f32": TS_F32, "fp32": TS_F32, "float32": TS_F32,
          "bf16": TS_BF16, "bfloat16": TS_BF16,
          "f16": TS_F16, "fp16": TS_F16, "float16": TS_F16}
PATHS = {"auto": TS_PATH_AUTO, "stream": TS_PATH_STREAM, "umma": TS_PATH_UMMA}

DTYPE_NAMES = {TS_F32: "fp32", TS_BF16: "bf16", TS_F16: "fp16"}
TS_FILE_INDEX, TS_FILE_TOKSTORE = 1, 2


class FileInfo(C.Structure):
    """``ts_file_info`` (include/tristage.h): the header of one shard file."""
    _fields_ = [("kind", C.c_int32), ("version", C.c_int32), ("dim", C.c_int32), ("ld", C.c_int32),
                ("dtype", C.c_int32), ("metric", C.c_int32),
                ("n", C.c_int64), ("nrows", C.c_int64), ("ntokens", C.c_int64), ("id_base", C.c_int64),
                ("table_offset", C.c_uint64), ("table_bytes", C.c_uint64),
                ("payload_offset", C.c_uint64), ("payload_bytes", C.c_uint64),
                ("table_hash", C.c_uint64), ("payload_hash", C.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


# every symbol include/tristage.h declares: name -> (restype, argtypes)
_vp, _i, _i64, _u = C.c_void_p, C.c_int, C.c_int64, C.c_uint
SYMBOLS = {
    "ts_abi_version": (_i, []),
    "ts_last_error": (C.c_char_p, []),
    "ts_device_count": (_i, []),
    "ts_index_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i64]),
    "ts_index_destroy": (_i, [_vp]),
    "ts_index_add": (_i, [_vp, _vp, _i64, _i, _i, _i, _vp]),
    "ts_index_ntotal": (_i64, [_vp]),
    "ts_index_dim": (_i, [_vp]),
    "ts_index_reset": (_i, [_vp]),
    "ts_index_set_id_base": (_i, [_vp, _i64]),
    "ts_index_search": (_i, [_vp, _vp, _i, _i, _i, _u, _i, _vp, _vp, _vp]),
    "ts_index_search_host": (_i, [_vp, _vp, _i, _i, _i, _u, _i, _vp, _vp, _vp]),
    "ts_topk_merge": (_i, [_i, _vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "ts_topk_merge_packed": (_i, [_i, _vp, _i64, _i64, _i, _i, _i, _vp, _vp, _vp]),
    "ts_exchange_push": (_i, [_i, _vp, _i64, _vp, _i, _i, _i64, _i64, _i, _u, _vp]),
    "ts_exchange_wait_merge": (_i, [_i, _vp, _i, _i, _i, _i64, _i64, _i64, _i, _u, _vp, _vp, _vp]),
    "ts_exchange_wait_sum": (_i, [_i, _vp, _i, _i64, _i64, _i64, _i, _u, _vp, _vp]),
    "ts_maxsim_scatter": (_i, [_vp, _vp, _i, _vp, _i, _i, _vp, _vp, _i, _i, _u, _vp, _i, _i, _i64, _i64, _u, _vp]),
    "ts_exchange_wait_take": (_i, [_i, _vp, _vp, _i, _u, _i64, _vp, _vp]),
    "ts_index_debug_timeline": (_i, [_vp, _vp]),
    "ts_index_save": (_i, [_vp, C.c_char_p]),
    "ts_index_load": (_i, [C.POINTER(_vp), _i, C.c_char_p]),
    "ts_index_append_file": (_i, [_vp, C.c_char_p, _i64, _i64, _vp]),
    "ts_index_dtype": (_i, [_vp]),
    "ts_index_metric": (_i, [_vp]),
    "ts_index_get_rows": (_i, [_vp, _i64, _i64, _vp]),
    "ts_index_launch_count": (_i64, [_vp]),
    "ts_index_set_profiling": (_i, [_vp, _i]),
    "ts_index_scan_time": (_i, [_vp, C.POINTER(C.c_float), C.POINTER(_i)]),
    "ts_tokstore_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i64, _i64]),
    "ts_tokstore_destroy": (_i, [_vp]),
    "ts_tokstore_add": (_i, [_vp, _vp, _i, _i, _vp, _i, _i, _vp]),
    "ts_tokstore_ndocs": (_i64, [_vp]),
    "ts_tokstore_ntokens": (_i64, [_vp]),
    "ts_tokstore_reset": (_i, [_vp]),
    "ts_tokstore_set_id_base": (_i, [_vp, _i64]),
    "ts_tokstore_launch_count": (_i64, [_vp]),
    "ts_tokstore_save": (_i, [_vp, C.c_char_p]),
    "ts_tokstore_load": (_i, [C.POINTER(_vp), _i, C.c_char_p]),
    "ts_tokstore_append_file": (_i, [_vp, C.c_char_p, _i64, _i64, _vp]),
    "ts_tokstore_dim": (_i, [_vp]),
    "ts_exchange_buffer_bytes": (C.c_int64, [_i, _i, _i]),
    "ts_exchange_create": (_i, [C.POINTER(_vp), _i, _i, _i, _vp, _i, _i]),
    "ts_exchange_destroy": (_i, [_vp]),
    "ts_index_search_push": (_i, [_vp, _vp, _vp, _i, _i, _i, C.c_uint, _i, _vp]),
    "ts_exchange_merge": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "ts_index_search_sharded": (_i, [_vp, _vp, _vp, _i, _i, _i, C.c_uint, _i, _vp, _vp, _vp]),
    "ts_index_search_sharded_host": (_i, [_vp, _vp, _vp, _i, _i, _i, C.c_uint, _i, _vp, _vp, _vp]),
    "ts_tokstore_dtype": (_i, [_vp]),
    "ts_tokstore_layout": (_i, [_vp]),
    "ts_tokstore_set_profiling": (_i, [_vp, _i]),
    "ts_tokstore_scan_time": (_i, [_vp, C.POINTER(C.c_float), C.POINTER(_i)]),
    "ts_maxsim": (_i, [_vp, _vp, _i, _vp, _i, _i, _vp, _vp, _i, _i, _u, _vp, _vp]),
    "ts_maxsim_host": (_i, [_vp, _vp, _i, _vp, _i, _i, _vp, _vp, _i, _i, _u, _vp, _vp]),
    "ts_rank_desc": (_i, [_i, _vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "ts_bm25_create": (_i, [C.POINTER(_vp), _i, _i64, _i64, _vp, _vp, _vp]),
    "ts_bm25_destroy": (_i, [_vp]),
    "ts_bm25_ndocs": (_i64, [_vp]),
    "ts_bm25_launch_count": (_i64, [_vp]),
    "ts_bm25_search_host": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "ts_hybrid_fuse_host": (_i, [_i, _i, _i, C.c_double, C.c_double, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i,
                                 _vp, _vp, _vp, _vp]),
    "ts_ivf_create": (_i, [C.POINTER(_vp), _vp, _i]),
    "ts_ivf_destroy": (_i, [_vp]),
    "ts_ivf_nlist": (_i, [_vp]),
    "ts_ivf_is_trained": (_i, [_vp]),
    "ts_ivf_nassigned": (_i64, [_vp]),
    "ts_ivf_launch_count": (_i64, [_vp]),
    "ts_ivf_set_centroids": (_i, [_vp, _vp, _vp]),
    "ts_ivf_get_centroids": (_i, [_vp, _vp]),
    "ts_ivf_sync": (_i, [_vp, _vp]),
    "ts_ivf_set_assignments": (_i, [_vp, _vp, _i64, _vp]),
    "ts_ivf_get_assignments": (_i, [_vp, _vp, _i64]),
    "ts_ivf_list_sizes": (_i, [_vp, _vp]),
    "ts_ivf_coarse_host": (_i, [_vp, _vp, _i, _i, _u, _vp, _vp, _vp]),
    "ts_ivf_search": (_i, [_vp, _vp, _i, _i, _i, _i, _u, _vp, _vp, _vp]),
    "ts_ivf_search_host": (_i, [_vp, _vp, _i, _i, _i, _i, _u, _vp, _vp, _vp]),
    "ts_file_probe": (_i, [C.c_char_p, C.POINTER(FileInfo)]),
    "ts_file_verify": (_i, [C.c_char_p]),
    "ts_file_write_index_host": (_i, [C.c_char_p, _i, _i, _i, _i64, _i64, _vp, _vp]),
    "ts_file_write_tokstore_host": (_i, [C.c_char_p, _i, _i, _i64, _i64, _vp, _vp]),
}

_lib = None
_lock = threading.Lock()


class TristageError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libtristage error {code}: {msg}")
        self.code = code
        self.msg = msg


def build(verbose: bool = False) -> str:
    """Compile libtristage.so for sm_100a (nvcc cross-compiles without a GPU)."""
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    cmd = ["make", "-C", CSRC, "-j", str(min(8, os.cpu_count() or 1))]
    out = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout)
    if out.returncode != 0:
        raise RuntimeError("building libtristage.so failed")
    return LIB_PATH


def lib():
    """The loaded library; raises (never falls back) when it is missing."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(
                    f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(or `make -C tristage_rag_b200/csrc`). There is no CPU fallback.")
            L = C.CDLL(LIB_PATH)
            for name, (res, args) in SYMBOLS.items():
                fn = getattr(L, name)          # AttributeError if the ABI lost a symbol
                fn.restype, fn.argtypes = res, args
            if L.ts_abi_version() != 1:
                raise ImportError("libtristage ABI version mismatch")
            _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise TristageError(rc, (lib().ts_last_error() or b"").decode("utf-8", "replace"))


def _stream_ptr(device_index: int):
    import torch

    return C.c_void_p(torch.cuda.current_stream(device_index).cuda_stream)


def _torch_dtype(code: int):
    import torch

    return {TS_F32: torch.float32, TS_BF16: torch.bfloat16, TS_F16: torch.float16}[code]


def _code_of_torch(dt) -> int:
    import torch

    return {torch.float32: TS_F32, torch.bfloat16: TS_BF16, torch.float16: TS_F16}[dt]


def _locked(fn):
    """Serialise calls on one handle: the library keeps per-handle scratch (include/tristage.h: "a handle is
    not re-entrant"); the reference's callers are single threaded except the Flask dev server."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *a, **kw):
        with self._lock:
            return fn(self, *a, **kw)
    return wrapper


class Index:
    """One Stage-1 corpus shard on one GPU (``ts_index``)."""

    def __init__(self, dim: int, dtype: str = "bf16", metric: str = "ip", device: int = 0,
                 reserve_rows: int = 0, _handle=None):
        self.dim, self.device = int(dim), int(device)
        self.dtype = DTYPES[dtype]
        self.metric = TS_METRIC_COSINE if metric in ("cosine", "cos") else TS_METRIC_IP
        self._lock = threading.RLock()     # include/tristage.h: one call at a time per handle
        if _handle is not None:
            self._h = _handle
        else:
            h = C.c_void_p()
            check(lib().ts_index_create(C.byref(h), self.device, self.dim, self.dtype, self.metric, int(reserve_rows)))
            self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h is not None and _lib is not None:
            _lib.ts_index_destroy(h)

    @property
    def ntotal(self) -> int:
        return int(lib().ts_index_ntotal(self._h))

    @property
    def launches(self) -> int:
        return int(lib().ts_index_launch_count(self._h))

    def set_id_base(self, base: int) -> None:
        check(lib().ts_index_set_id_base(self._h, int(base)))

    def set_profiling(self, on: bool) -> None:
        check(lib().ts_index_set_profiling(self._h, int(on)))

    def scan_time_ms(self):
        """(mean ms, count) of the scan kernels recorded since the last call."""
        ms, n = C.c_float(), C.c_int()
        check(lib().ts_index_scan_time(self._h, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    def debug_timeline(self):
        """TS_DBG_TIMELINE=1: the 16 globaltimer slots (ns) of the last search step (ts_index_debug_timeline)."""
        import numpy as np

        out = np.zeros(16, np.uint64)
        check(lib().ts_index_debug_timeline(self._h, C.c_void_p(out.ctypes.data)))
        return out

    @_locked
    def reset(self) -> None:
        check(lib().ts_index_reset(self._h))

    @_locked
    def add(self, x, normalize: bool = False) -> None:
        """x: numpy fp32 [n, dim] (host) or torch tensor (cuda: fp32 / storage dtype; cpu: fp32)."""
        import numpy as np
        import torch

        if isinstance(x, np.ndarray):
            x = np.ascontiguousarray(x, dtype=np.float32)
            assert x.ndim == 2 and x.shape[1] == self.dim, x.shape
            check(lib().ts_index_add(self._h, C.c_void_p(x.ctypes.data), x.shape[0], TS_F32, 0, int(normalize),
                                     _stream_ptr(self.device)))
            return
        assert isinstance(x, torch.Tensor) and x.dim() == 2 and x.shape[1] == self.dim, x.shape
        x = x.contiguous()
        on_dev = 1 if x.is_cuda else 0
        if on_dev:
            assert x.device.index == self.device
        check(lib().ts_index_add(self._h, C.c_void_p(x.data_ptr()), x.shape[0], _code_of_torch(x.dtype), on_dev,
                                 int(normalize), _stream_ptr(self.device)))
        if on_dev:
            torch.cuda.current_stream(self.device).synchronize()   # x may be freed by the caller

    @_locked
    def search(self, q, k: int, normalize_q: bool = False, path: str = "auto"):
        """q: cuda tensor [B, dim] (fp32 or storage dtype) -> (scores[B,k] f32, ids[B,k] i64) on the device.
        Asynchronous on the current stream."""
        import torch

        assert q.is_cuda and q.dim() == 2 and q.shape[1] == self.dim, (q.shape, self.dim)
        q = q.contiguous()
        B = q.shape[0]
        dev = torch.device("cuda", self.device)
        scores = torch.empty((B, k), dtype=torch.float32, device=dev)
        ids = torch.empty((B, k), dtype=torch.int64, device=dev)
        check(lib().ts_index_search(self._h, C.c_void_p(q.data_ptr()), _code_of_torch(q.dtype), B, int(k),
                                    TS_FLAG_NORMALIZE_Q if normalize_q else 0, PATHS[path],
                                    C.c_void_p(scores.data_ptr()), C.c_void_p(ids.data_ptr()),
                                    _stream_ptr(self.device)))
        return scores, ids

    @_locked
    def search_packed(self, q, k: int, blob, normalize_q: bool = False, path: str = "auto"):
        """Like search(), but scores and ids land in one uint8 cuda buffer (packed_layout) so a
        single all-gather moves both."""
        q = q.contiguous()
        B = q.shape[0]
        ids_off, nbytes = packed_layout(B, k)
        assert blob.is_cuda and blob.numel() >= nbytes and blob.data_ptr() % 8 == 0
        check(lib().ts_index_search(self._h, C.c_void_p(q.data_ptr()), _code_of_torch(q.dtype), B, int(k),
                                    TS_FLAG_NORMALIZE_Q if normalize_q else 0, PATHS[path],
                                    C.c_void_p(blob.data_ptr()), C.c_void_p(blob.data_ptr() + ids_off),
                                    _stream_ptr(self.device)))
        return blob

    @_locked
    def search_host(self, q, k: int, normalize_q: bool = False, path: str = "auto", out=None):
        """q: numpy fp32 [B, dim] -> (D[B,k] f32, I[B,k] i64) numpy -- the faiss call shape.
        ``out=(D, I)`` reuses caller-owned result arrays (e.g. views of pinned memory, which makes the
        device-to-host copies asynchronous up to the final synchronise)."""
        import numpy as np

        q = np.ascontiguousarray(q, dtype=np.float32)
        assert q.ndim == 2 and q.shape[1] == self.dim, (q.shape, self.dim)
        B = q.shape[0]
        if out is None:
            D = np.empty((B, k), np.float32)
            I = np.empty((B, k), np.int64)
        else:
            D, I = out
            assert D.shape == (B, k) and D.dtype == np.float32 and D.flags.c_contiguous
            assert I.shape == (B, k) and I.dtype == np.int64 and I.flags.c_contiguous
        check(lib().ts_index_search_host(self._h, C.c_void_p(q.ctypes.data), TS_F32, B, int(k),
                                         TS_FLAG_NORMALIZE_Q if normalize_q else 0, PATHS[path],
                                         C.c_void_p(D.ctypes.data), C.c_void_p(I.ctypes.data),
                                         _stream_ptr(self.device)))
        return D, I

    @_locked
    def get_rows(self, start: int, n: int):
        import numpy as np

        out = np.empty((n, self.dim), np.float32)
        check(lib().ts_index_get_rows(self._h, int(start), int(n), C.c_void_p(out.ctypes.data)))
        return out

    @_locked
    def save(self, path: str) -> None:
        check(lib().ts_index_save(self._h, os.fsencode(path)))

    @_locked
    def append_file(self, path: str, row_lo: int, n_rows: int) -> None:
        """Append rows [row_lo, row_lo + n_rows) of an index shard file (re-sharding primitive)."""
        check(lib().ts_index_append_file(self._h, os.fsencode(path), int(row_lo), int(n_rows),
                                         _stream_ptr(self.device)))

    @classmethod
    def load(cls, path: str, device: int = 0) -> "Index":
        h = C.c_void_p()
        check(lib().ts_index_load(C.byref(h), int(device), os.fsencode(path)))
        obj = cls.__new__(cls)
        obj.dim, obj.device, obj._h = int(lib().ts_index_dim(h)), int(device), h
        obj.dtype, obj.metric = int(lib().ts_index_dtype(h)), int(lib().ts_index_metric(h))
        obj._lock = threading.RLock()
        return obj


class IVF:
    """Inverted lists over the rows of an ``Index`` (``ts_ivf``): the approximate mode that stands in for the
    reference's ``faiss.IndexIVFFlat`` branch (``/root/reference/src/stage1_retriever.py:262-273``).  The
    corpus is not copied; the lists hold row numbers.  Training (k-means) is ``tristage_rag_b200/ivf.py``."""

    def __init__(self, base: Index, nlist: int):
        self.base, self.nlist, self.device = base, int(nlist), base.device     # keeps `base` alive
        self._lock = base._lock            # the two handles share rows and are used one call at a time
        h = C.c_void_p()
        check(lib().ts_ivf_create(C.byref(h), base._h, self.nlist))
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h is not None and _lib is not None:
            _lib.ts_ivf_destroy(h)

    @property
    def is_trained(self) -> bool:
        return bool(lib().ts_ivf_is_trained(self._h))

    @property
    def nassigned(self) -> int:
        return int(lib().ts_ivf_nassigned(self._h))

    @property
    def launches(self) -> int:
        return int(lib().ts_ivf_launch_count(self._h))

    @_locked
    def set_centroids(self, centroids) -> None:
        import numpy as np

        c = np.ascontiguousarray(centroids, np.float32)
        assert c.shape == (self.nlist, self.base.dim), (c.shape, self.nlist, self.base.dim)
        check(lib().ts_ivf_set_centroids(self._h, C.c_void_p(c.ctypes.data), _stream_ptr(self.device)))

    @_locked
    def centroids(self):
        import numpy as np

        out = np.empty((self.nlist, self.base.dim), np.float32)
        check(lib().ts_ivf_get_centroids(self._h, C.c_void_p(out.ctypes.data)))
        return out

    @_locked
    def sync(self) -> None:
        """Put every row added to the base index since the last call into its list."""
        check(lib().ts_ivf_sync(self._h, _stream_ptr(self.device)))

    @_locked
    def set_assignments(self, assign) -> None:
        import numpy as np

        a = np.ascontiguousarray(assign, np.int32)
        check(lib().ts_ivf_set_assignments(self._h, C.c_void_p(a.ctypes.data) if a.size else None, a.size,
                                           _stream_ptr(self.device)))

    @_locked
    def assignments(self):
        import numpy as np

        out = np.empty(self.nassigned, np.int32)
        check(lib().ts_ivf_get_assignments(self._h, C.c_void_p(out.ctypes.data) if out.size else None, out.size))
        return out

    def list_sizes(self):
        import numpy as np

        out = np.empty(self.nlist, np.int64)
        check(lib().ts_ivf_list_sizes(self._h, C.c_void_p(out.ctypes.data)))
        return out

    @_locked
    def coarse_host(self, q, nprobe: int, normalize_q: bool = False):
        """quantizer.search(q, nprobe): ([B, nprobe] int32 lists, [B, nprobe] fp32 scores), best first."""
        import numpy as np

        q = np.ascontiguousarray(q, np.float32)
        assert q.ndim == 2 and q.shape[1] == self.base.dim, (q.shape, self.base.dim)
        B = q.shape[0]
        lists, scores = np.empty((B, nprobe), np.int32), np.empty((B, nprobe), np.float32)
        check(lib().ts_ivf_coarse_host(self._h, C.c_void_p(q.ctypes.data), B, int(nprobe),
                                       TS_FLAG_NORMALIZE_Q if normalize_q else 0, C.c_void_p(lists.ctypes.data),
                                       C.c_void_p(scores.ctypes.data), _stream_ptr(self.device)))
        return lists, scores

    @_locked
    def search(self, q, k: int, nprobe: int, normalize_q: bool = False):
        """q: cuda tensor [B, dim] -> (scores [B,k] f32, ids [B,k] i64) on the device, asynchronous."""
        import torch

        assert q.is_cuda and q.dim() == 2 and q.shape[1] == self.base.dim, (q.shape, self.base.dim)
        q = q.contiguous()
        B = q.shape[0]
        dev = torch.device("cuda", self.device)
        scores = torch.empty((B, k), dtype=torch.float32, device=dev)
        ids = torch.empty((B, k), dtype=torch.int64, device=dev)
        check(lib().ts_ivf_search(self._h, C.c_void_p(q.data_ptr()), _code_of_torch(q.dtype), B, int(k), int(nprobe),
                                  TS_FLAG_NORMALIZE_Q if normalize_q else 0, C.c_void_p(scores.data_ptr()),
                                  C.c_void_p(ids.data_ptr()), _stream_ptr(self.device)))
        return scores, ids

    @_locked
    def search_packed(self, q, k: int, nprobe: int, blob, normalize_q: bool = False):
        """Like search(), but scores and ids land in one uint8 cuda buffer (packed_layout): the shape the
        multi-GPU all-gather / peer-memory exchange moves (dist.ShardedIndex over dist.IVFShard)."""
        q = q.contiguous()
        B = q.shape[0]
        ids_off, nbytes = packed_layout(B, k)
        assert blob.is_cuda and blob.numel() >= nbytes and blob.data_ptr() % 8 == 0
        check(lib().ts_ivf_search(self._h, C.c_void_p(q.data_ptr()), _code_of_torch(q.dtype), B, int(k), int(nprobe),
                                  TS_FLAG_NORMALIZE_Q if normalize_q else 0, C.c_void_p(blob.data_ptr()),
                                  C.c_void_p(blob.data_ptr() + ids_off), _stream_ptr(self.device)))
        return blob

    @_locked
    def search_host(self, q, k: int, nprobe: int, normalize_q: bool = False):
        """numpy in / numpy out -- the faiss call shape (``index.nprobe = nprobe; index.search(q, k)``)."""
        import numpy as np

        q = np.ascontiguousarray(q, np.float32)
        assert q.ndim == 2 and q.shape[1] == self.base.dim, (q.shape, self.base.dim)
        B = q.shape[0]
        D, I = np.empty((B, k), np.float32), np.empty((B, k), np.int64)
        check(lib().ts_ivf_search_host(self._h, C.c_void_p(q.ctypes.data), TS_F32, B, int(k), int(nprobe),
                                       TS_FLAG_NORMALIZE_Q if normalize_q else 0, C.c_void_p(D.ctypes.data),
                                       C.c_void_p(I.ctypes.data), _stream_ptr(self.device)))
        return D, I


def file_probe(path: str) -> dict:
    """Header of a shard file as a dict (host only: works without a GPU)."""
    fi = FileInfo()
    check(lib().ts_file_probe(os.fsencode(path), C.byref(fi)))
    return fi.as_dict()


def file_verify(path: str) -> None:
    """Recompute the section checksums of a shard file; raises TristageError on a mismatch."""
    check(lib().ts_file_verify(os.fsencode(path)))


def _storage_view(a, dtype_code: int):
    """numpy array already in the storage dtype: float32, or uint16 bit patterns of bf16 / fp16
    (float16 arrays are accepted for fp16)."""
    import numpy as np

    a = np.ascontiguousarray(a)
    if dtype_code == TS_F32:
        assert a.dtype == np.float32, a.dtype
    elif dtype_code == TS_F16 and a.dtype == np.float16:
        a = a.view(np.uint16)
    else:
        assert a.dtype == np.uint16, f"pass {DTYPE_NAMES[dtype_code]} rows as uint16 bit patterns, got {a.dtype}"
    return a


def write_index_file(path: str, rows, dtype: str = "bf16", metric: str = "ip", id_base: int = 0, inv_norm=None,
                     dim: int = None) -> None:
    """Write an index shard file from host rows ALREADY in the storage dtype ([n, ld] with ld = dim
    rounded up to 16 bytes).  Layout only, no arithmetic; host only."""
    import numpy as np

    code = DTYPES[dtype]
    rows = _storage_view(rows, code)
    n, ld = rows.shape
    dim = ld if dim is None else int(dim)
    per = 4 if code == TS_F32 else 8
    assert (dim + per - 1) // per * per == ld, (dim, ld)
    m = TS_METRIC_COSINE if metric in ("cosine", "cos") else TS_METRIC_IP
    inv = np.ascontiguousarray(inv_norm, np.float32) if inv_norm is not None else None
    check(lib().ts_file_write_index_host(os.fsencode(path), dim, code, m, n, int(id_base),
                                         C.c_void_p(rows.ctypes.data) if n else None,
                                         C.c_void_p(inv.ctypes.data) if inv is not None else None))


def write_tokstore_file(path: str, tok, lens, dtype: str = "bf16", id_base: int = 0) -> None:
    """Write a token shard file from un-padded host token rows already in the storage dtype."""
    import numpy as np

    code = DTYPES[dtype]
    tok = _storage_view(tok, code)
    lens = np.ascontiguousarray(lens, np.int32)
    assert tok.ndim == 2 and tok.shape[0] == int(lens.sum()), (tok.shape, int(lens.sum()))
    check(lib().ts_file_write_tokstore_host(os.fsencode(path), tok.shape[1], code, len(lens), int(id_base),
                                            C.c_void_p(lens.ctypes.data) if len(lens) else None,
                                            C.c_void_p(tok.ctypes.data) if len(lens) else None))


def topk_merge(scores, ids, device: int = 0):
    """[L, B, k] per-shard results (cuda tensors) -> merged ([B,k], [B,k])."""
    import torch

    L, B, k = scores.shape
    scores, ids = scores.contiguous(), ids.contiguous()
    dev = torch.device("cuda", device)
    out_s = torch.empty((B, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((B, k), dtype=torch.int64, device=dev)
    check(lib().ts_topk_merge(device, C.c_void_p(scores.data_ptr()), C.c_void_p(ids.data_ptr()), L, B, k,
                              C.c_void_p(out_s.data_ptr()), C.c_void_p(out_i.data_ptr()), _stream_ptr(device)))
    return out_s, out_i


class Exchange:
    """Fused multi-GPU exchange state (``ts_exchange``): this rank's receive buffer plus the peers' buffers as seen
    from this device.  ``peer_bases``: int sequence [n_ranks] of buffer base addresses (own buffer at [rank]);
    every buffer holds ``Exchange.buffer_bytes(n_ranks, B_max, k_max)`` ZEROED bytes of peer-mapped memory."""

    def __init__(self, device: int, rank: int, n_ranks: int, peer_bases, B_max: int, k_max: int):
        import numpy as np

        self.device, self.rank, self.n_ranks, self.B_max, self.k_max = int(device), int(rank), int(n_ranks), int(B_max), int(k_max)
        bases = np.ascontiguousarray(peer_bases, np.int64)
        assert len(bases) == n_ranks
        h = C.c_void_p()
        check(lib().ts_exchange_create(C.byref(h), self.device, self.rank, self.n_ranks, C.c_void_p(bases.ctypes.data),
                                       self.B_max, self.k_max))
        self._h = h

    @staticmethod
    def buffer_bytes(n_ranks: int, B_max: int, k_max: int) -> int:
        n = int(lib().ts_exchange_buffer_bytes(int(n_ranks), int(B_max), int(k_max)))
        if n <= 0:
            raise ValueError("bad exchange capacity")
        return n

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h is not None and _lib is not None:
            _lib.ts_exchange_destroy(h)

    def search(self, index: "Index", q, k: int, normalize_q: bool = False, path: str = "auto"):
        """q: cuda tensor [B, dim] replicated on every rank -> merged (scores, ids) [B, k], identical on every rank.
        Four launches (query prep, scan, select+push, wait+merge), asynchronous on the current stream."""
        import torch

        q = q.contiguous()
        B = q.shape[0]
        dev = torch.device("cuda", self.device)
        out_s = torch.empty((B, k), dtype=torch.float32, device=dev)
        out_i = torch.empty((B, k), dtype=torch.int64, device=dev)
        check(lib().ts_index_search_sharded(index._h, self._h, C.c_void_p(q.data_ptr()), _code_of_torch(q.dtype), B, int(k),
                                            TS_FLAG_NORMALIZE_Q if normalize_q else 0, PATHS[path],
                                            C.c_void_p(out_s.data_ptr()), C.c_void_p(out_i.data_ptr()), _stream_ptr(self.device)))
        return out_s, out_i

    def search_host(self, index: "Index", q, k: int, normalize_q: bool = False, path: str = "auto", out=None):
        """numpy fp32 [B, dim] in, (D, I) numpy out: H2D, the four launches, D2H and the synchronise in ONE C call."""
        import numpy as np

        q = np.ascontiguousarray(q, dtype=np.float32)
        B = q.shape[0]
        D, I = out if out is not None else (np.empty((B, k), np.float32), np.empty((B, k), np.int64))
        check(lib().ts_index_search_sharded_host(index._h, self._h, C.c_void_p(q.ctypes.data), TS_F32, B, int(k),
                                                 TS_FLAG_NORMALIZE_Q if normalize_q else 0, PATHS[path],
                                                 C.c_void_p(D.ctypes.data), C.c_void_p(I.ctypes.data), _stream_ptr(self.device)))
        return D, I


def packed_layout(B: int, k: int):
    """(ids_offset_bytes, blob_bytes) of one packed [B, k] result: fp32 scores, then int64 ids."""
    ids_off = (B * k * 4 + 7) // 8 * 8
    return ids_off, ids_off + B * k * 8


def topk_merge_packed(blob, n_lists: int, B: int, k: int, device: int = 0):
    """blob: uint8 cuda tensor [n_lists * blob_bytes] (one all-gather of packed results)."""
    import torch

    ids_off, nbytes = packed_layout(B, k)
    dev = torch.device("cuda", device)
    out_s = torch.empty((B, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((B, k), dtype=torch.int64, device=dev)
    check(lib().ts_topk_merge_packed(device, C.c_void_p(blob.data_ptr()), nbytes, ids_off, n_lists, B, k,
                                     C.c_void_p(out_s.data_ptr()), C.c_void_p(out_i.data_ptr()), _stream_ptr(device)))
    return out_s, out_i


def rank_desc(scores, top_k: int, n_cand=None, device: int = 0):
    """Stable descending top_k positions per row of scores[B, C] (cuda)."""
    import torch

    B, Cn = scores.shape
    scores = scores.contiguous()
    dev = torch.device("cuda", device)
    out_s = torch.empty((B, top_k), dtype=torch.float32, device=dev)
    out_p = torch.empty((B, top_k), dtype=torch.int32, device=dev)
    nc = C.c_void_p(n_cand.data_ptr()) if n_cand is not None else None
    check(lib().ts_rank_desc(device, C.c_void_p(scores.data_ptr()), nc, B, Cn, int(top_k),
                             C.c_void_p(out_s.data_ptr()), C.c_void_p(out_p.data_ptr()), _stream_ptr(device)))
    return out_s, out_p


class BM25:
    """Device-resident BM25 postings (``ts_bm25``): CSR term offsets, document ids and one fp64
    weight per posting, as built by ``stage1_retriever.BM25Index._build_postings``."""

    SCRATCH_BYTES = 2 << 30      # budget of the dense per-query score accumulator; ``search`` slices the batch to fit

    def __init__(self, n_docs: int, term_off, post_doc, post_w, device: int = 0):
        import numpy as np

        self.device = int(device)
        self._lock = threading.RLock()     # per-handle scratch: one call at a time
        off = np.ascontiguousarray(term_off, np.int64)
        docs = np.ascontiguousarray(post_doc, np.int32)
        w = np.ascontiguousarray(post_w, np.float64)
        assert len(docs) == len(w) == int(off[-1])
        self.n_docs, self.n_terms = int(n_docs), len(off) - 1
        h = C.c_void_p()
        check(lib().ts_bm25_create(C.byref(h), self.device, self.n_docs, self.n_terms, C.c_void_p(off.ctypes.data),
                                   C.c_void_p(docs.ctypes.data) if len(docs) else None,
                                   C.c_void_p(w.ctypes.data) if len(w) else None))
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h is not None and _lib is not None:
            _lib.ts_bm25_destroy(h)

    @property
    def launches(self) -> int:
        return int(lib().ts_bm25_launch_count(self._h))

    @_locked
    def search(self, query_terms, top_k: int):
        """query_terms: one int sequence of term ids per query (query-token order, repeats kept).
        -> (scores [B, top_k] float64, ids [B, top_k] int64, -1 beyond the corpus)."""
        import numpy as np

        B = len(query_terms)
        scores = np.empty((B, top_k), np.float64)
        ids = np.empty((B, top_k), np.int64)
        # the kernels keep a dense [B, n_docs] fp64 accumulator: walk the batch in slices that keep it under
        # BM25_SCRATCH_BYTES (2 GiB) instead of asking for B * n_docs * 8 bytes at once
        per_query = max(1, int(self.n_docs)) * 8
        step = max(1, min(B, self.SCRATCH_BYTES // per_query))
        for b0 in range(0, B, step):
            qs = query_terms[b0:b0 + step]
            nb = len(qs)
            off = np.zeros(nb + 1, np.int64)
            off[1:] = np.cumsum([len(q) for q in qs])
            flat = np.ascontiguousarray(np.concatenate([np.asarray(q, np.int32) for q in qs])
                                        if int(off[-1]) else np.zeros(0, np.int32), np.int32)
            sc = np.empty((nb, top_k), np.float64)
            ii = np.empty((nb, top_k), np.int64)
            check(lib().ts_bm25_search_host(self._h, C.c_void_p(flat.ctypes.data) if len(flat) else None,
                                            C.c_void_p(off.ctypes.data), nb, int(top_k), C.c_void_p(sc.ctypes.data),
                                            C.c_void_p(ii.ctypes.data), _stream_ptr(self.device)))
            scores[b0:b0 + nb], ids[b0:b0 + nb] = sc, ii
        return scores, ids


def hybrid_fuse(method: str, rrf_k: int, w_dense: float, w_bm25: float, dense_ids, dense_scores, bm25_ids, bm25_scores,
                top_k: int, device: int = 0):
    """RRF / weighted fusion on the device.  dense_ids [B, k1] int64 (-1 = unused slot), dense_scores
    [B, k1] float32, bm25_ids [B, k2] int64, bm25_scores [B, k2] float64 -> (ids [B, top_k] int64,
    scores [B, top_k] float64, n [B]) in the reference's order (stable descending)."""
    import numpy as np

    dense_ids = np.ascontiguousarray(dense_ids, np.int64)
    dense_scores = np.ascontiguousarray(dense_scores, np.float32)
    bm25_ids = np.ascontiguousarray(bm25_ids, np.int64)
    bm25_scores = np.ascontiguousarray(bm25_scores, np.float64)
    B, k1 = dense_ids.shape
    k2 = bm25_ids.shape[1]
    n_dense = np.ascontiguousarray((dense_ids >= 0).sum(axis=1), np.int32)      # valid slots lead (FAISS pads the tail)
    n_bm = np.ascontiguousarray((bm25_ids >= 0).sum(axis=1), np.int32)
    out_ids = np.empty((B, top_k), np.int64)
    out_scores = np.empty((B, top_k), np.float64)
    out_n = np.empty(B, np.int32)
    vp = lambda a: C.c_void_p(a.ctypes.data) if a.size else None                # noqa: E731
    check(lib().ts_hybrid_fuse_host(int(device), 0 if method == "rrf" else 1, int(rrf_k), float(w_dense), float(w_bm25),
                                    vp(dense_ids), vp(dense_scores), vp(n_dense), k1, vp(bm25_ids), vp(bm25_scores),
                                    vp(n_bm), k2, B, int(top_k), vp(out_ids), vp(out_scores), vp(out_n),
                                    _stream_ptr(device)))
    return out_ids, out_scores, out_n


class TokStore:
    """One Stage-2 token-embedding shard on one GPU (``ts_tokstore``)."""

    def __init__(self, dim: int, dtype: str = "bf16", device: int = 0, reserve_docs: int = 0,
                 reserve_tokens: int = 0):
        self.dim, self.device = int(dim), int(device)
        self.dtype = DTYPES[dtype]
        self._lock = threading.RLock()     # one call at a time per handle
        h = C.c_void_p()
        check(lib().ts_tokstore_create(C.byref(h), self.device, self.dim, self.dtype, int(reserve_docs),
                                       int(reserve_tokens)))
        self._h = h

    @_locked
    def save(self, path: str) -> None:
        check(lib().ts_tokstore_save(self._h, os.fsencode(path)))

    @_locked
    def append_file(self, path: str, doc_lo: int, n_docs: int) -> None:
        """Append docs [doc_lo, doc_lo + n_docs) of a token shard file (re-sharding primitive)."""
        check(lib().ts_tokstore_append_file(self._h, os.fsencode(path), int(doc_lo), int(n_docs),
                                            _stream_ptr(self.device)))

    @classmethod
    def load(cls, path: str, dim: int = None, dtype: str = None, device: int = 0) -> "TokStore":
        """dim / dtype come from the file; when given they are checked against it."""
        h = C.c_void_p()
        check(lib().ts_tokstore_load(C.byref(h), int(device), os.fsencode(path)))
        obj = cls.__new__(cls)
        obj.dim, obj.dtype = int(lib().ts_tokstore_dim(h)), int(lib().ts_tokstore_dtype(h))
        obj.device, obj._h = int(device), h
        obj._lock = threading.RLock()
        if (dim is not None and int(dim) != obj.dim) or (dtype is not None and DTYPES[dtype] != obj.dtype):
            raise ValueError(f"{path}: holds dim {obj.dim} {DTYPE_NAMES[obj.dtype]}, expected dim {dim} {dtype}")
        return obj

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h is not None and _lib is not None:
            _lib.ts_tokstore_destroy(h)

    @property
    def ndocs(self) -> int:
        return int(lib().ts_tokstore_ndocs(self._h))

    @property
    def layout(self) -> int:
        """0 = row-major shard, 1 = tile layout (the Stage-2 tensor kernel's operand image)."""
        return int(lib().ts_tokstore_layout(self._h))

    @property
    def ntokens(self) -> int:
        return int(lib().ts_tokstore_ntokens(self._h))

    @property
    def launches(self) -> int:
        return int(lib().ts_tokstore_launch_count(self._h))

    def set_id_base(self, base: int) -> None:
        check(lib().ts_tokstore_set_id_base(self._h, int(base)))

    def set_profiling(self, on: bool) -> None:
        check(lib().ts_tokstore_set_profiling(self._h, int(on)))

    def scan_time_ms(self):
        ms, n = C.c_float(), C.c_int()
        check(lib().ts_tokstore_scan_time(self._h, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    @_locked
    def reset(self) -> None:
        check(lib().ts_tokstore_reset(self._h))

    @_locked
    def add(self, tok, lens, normalize: bool = True) -> None:
        """tok: [sum(lens), dim] numpy fp32 (host) or torch tensor; lens: int sequence."""
        import numpy as np
        import torch

        lens = np.ascontiguousarray(np.asarray(lens, dtype=np.int32))
        total = int(lens.sum())
        if isinstance(tok, np.ndarray):
            tok = np.ascontiguousarray(tok, dtype=np.float32)
            assert tok.shape == (total, self.dim), (tok.shape, total, self.dim)
            check(lib().ts_tokstore_add(self._h, C.c_void_p(tok.ctypes.data), TS_F32, 0,
                                        C.c_void_p(lens.ctypes.data), len(lens), int(normalize),
                                        _stream_ptr(self.device)))
            return
        assert isinstance(tok, torch.Tensor) and tuple(tok.shape) == (total, self.dim), (tok.shape, total)
        tok = tok.contiguous()
        check(lib().ts_tokstore_add(self._h, C.c_void_p(tok.data_ptr()), _code_of_torch(tok.dtype),
                                    1 if tok.is_cuda else 0, C.c_void_p(lens.ctypes.data), len(lens),
                                    int(normalize), _stream_ptr(self.device)))

    @_locked
    def maxsim(self, q_tok, cand, q_len=None, n_cand=None, mode: int = TS_S2_MAXSIM, normalize_q: bool = True):
        """q_tok [B, Lq, dim] cuda, cand [B, C] int64 cuda -> scores [B, C] f32 cuda (async)."""
        import torch

        assert q_tok.is_cuda and q_tok.dim() == 3 and q_tok.shape[2] == self.dim
        q_tok, cand = q_tok.contiguous(), cand.contiguous()
        B, Lq, _ = q_tok.shape
        Cn = cand.shape[1]
        out = torch.empty((B, Cn), dtype=torch.float32, device=q_tok.device)
        check(lib().ts_maxsim(self._h, C.c_void_p(q_tok.data_ptr()), _code_of_torch(q_tok.dtype),
                              C.c_void_p(q_len.data_ptr()) if q_len is not None else None, B, Lq,
                              C.c_void_p(cand.data_ptr()),
                              C.c_void_p(n_cand.data_ptr()) if n_cand is not None else None, Cn, int(mode),
                              TS_FLAG_NORMALIZE_Q if normalize_q else 0, C.c_void_p(out.data_ptr()),
                              _stream_ptr(self.device)))
        return out

    @_locked
    def maxsim_scatter(self, q_tok, cand, peer_bases_dev, n_ranks: int, rank: int, matrix_offset: int, flags_offset: int,
                       seq: int, q_len=None, n_cand=None, mode: int = TS_S2_MAXSIM, normalize_q: bool = True) -> None:
        """ts_maxsim_scatter: score the candidates this shard owns and store every score into all ranks' [B, C]
        matrices (peer_bases_dev: int64 cuda tensor [n_ranks] of the receive buffers' addresses); async."""
        assert q_tok.is_cuda and q_tok.dim() == 3 and q_tok.shape[2] == self.dim
        q_tok, cand = q_tok.contiguous(), cand.contiguous()
        B, Lq, _ = q_tok.shape
        check(lib().ts_maxsim_scatter(self._h, C.c_void_p(q_tok.data_ptr()), _code_of_torch(q_tok.dtype),
                                      C.c_void_p(q_len.data_ptr()) if q_len is not None else None, B, Lq,
                                      C.c_void_p(cand.data_ptr()),
                                      C.c_void_p(n_cand.data_ptr()) if n_cand is not None else None, cand.shape[1], int(mode),
                                      TS_FLAG_NORMALIZE_Q if normalize_q else 0, C.c_void_p(peer_bases_dev.data_ptr()),
                                      int(n_ranks), int(rank), int(matrix_offset), int(flags_offset), int(seq),
                                      _stream_ptr(self.device)))

    @_locked
    def maxsim_host(self, q_tok, cand, q_len=None, n_cand=None, mode: int = TS_S2_MAXSIM,
                    normalize_q: bool = True):
        """numpy in / numpy out variant (copies inside the call, synchronises)."""
        import numpy as np

        q_tok = np.ascontiguousarray(q_tok, dtype=np.float32)
        cand = np.ascontiguousarray(cand, dtype=np.int64)
        B, Lq, _ = q_tok.shape
        Cn = cand.shape[1]
        out = np.empty((B, Cn), np.float32)
        ql = np.ascontiguousarray(q_len, dtype=np.int32) if q_len is not None else None
        nc = np.ascontiguousarray(n_cand, dtype=np.int32) if n_cand is not None else None
        check(lib().ts_maxsim_host(self._h, C.c_void_p(q_tok.ctypes.data), TS_F32,
                                   C.c_void_p(ql.ctypes.data) if ql is not None else None, B, Lq,
                                   C.c_void_p(cand.ctypes.data),
                                   C.c_void_p(nc.ctypes.data) if nc is not None else None, Cn, int(mode),
                                   TS_FLAG_NORMALIZE_Q if normalize_q else 0, C.c_void_p(out.ctypes.data),
                                   _stream_ptr(self.device)))
        return out
