"""ctypes binding of oracle/c/liboracle.so (TEST INFRASTRUCTURE, see oracle/__init__.py)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "c", "liboracle.so")
_lib = None


def build() -> str:
    subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        f32p, i64p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int64)
        L.oracle_flat_ip_search.argtypes = [f32p, ctypes.c_int64, ctypes.c_int, f32p, ctypes.c_int, ctypes.c_int, f32p, i64p]
        L.oracle_flat_ip_search.restype = ctypes.c_int
        L.oracle_normalize_rows.argtypes = [f32p, ctypes.c_int64, ctypes.c_int]
        L.oracle_maxsim.argtypes = [f32p, ctypes.c_int, f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.oracle_maxsim.restype = ctypes.c_float
        L.oracle_maxsim_batch.argtypes = [f32p, ctypes.c_int, f32p, i64p, ctypes.c_int, ctypes.c_int, ctypes.c_int, f32p]
        L.oracle_num_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _f(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def flat_ip_search(X, Q, k):
    X = np.ascontiguousarray(X, np.float32)
    Q = np.ascontiguousarray(Q, np.float32)
    B = Q.shape[0]
    D = np.empty((B, k), np.float32)
    I = np.empty((B, k), np.int64)
    rc = lib().oracle_flat_ip_search(_f(X), X.shape[0], X.shape[1], _f(Q), B, k, _f(D),
                                     I.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
    assert rc == 0
    return D, I


def normalize_rows(X):
    X = np.array(X, np.float32, order="C", copy=True)
    lib().oracle_normalize_rows(_f(X), X.shape[0], X.shape[1])
    return X


def maxsim(q, d, mode=0):
    q = np.ascontiguousarray(q, np.float32).reshape(-1, q.shape[-1])
    d = np.ascontiguousarray(d, np.float32).reshape(-1, d.shape[-1])
    return float(lib().oracle_maxsim(_f(q), q.shape[0], _f(d), d.shape[0], q.shape[1], mode))


def maxsim_batch(q, tok, off, mode=0):
    q = np.ascontiguousarray(q, np.float32).reshape(-1, q.shape[-1])
    tok = np.ascontiguousarray(tok, np.float32)
    off = np.ascontiguousarray(off, np.int64)
    out = np.empty(len(off) - 1, np.float32)
    lib().oracle_maxsim_batch(_f(q), q.shape[0], _f(tok), off.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                              len(off) - 1, q.shape[1], mode, _f(out))
    return out


def num_threads() -> int:
    return int(lib().oracle_num_threads())
