"""CPU: the C-ABI library loads, exports every symbol include/tristage.h
declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from tristage_rag_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "tristage.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ts_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _header_symbols()
    assert len(names) >= 25
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in tristage.h but not exported"
    assert sorted(_lib.SYMBOLS) == names, "python binding and header disagree"


def test_abi_version_and_error_string():
    L = _lib.lib()
    assert L.ts_abi_version() == 1
    assert isinstance(L.ts_last_error(), bytes)


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert _lib.lib().ts_device_count() == 0
    with pytest.raises(_lib.TristageError) as e:
        _lib.Index(16)
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(_lib.TristageError):
        _lib.TokStore(16)


def test_invalid_arguments_are_rejected_before_any_device_work():
    L = _lib.lib()
    h = ctypes.c_void_p()
    assert L.ts_index_create(ctypes.byref(h), 0, 0, _lib.TS_BF16, 0, 0) == -1      # dim 0
    assert L.ts_index_create(ctypes.byref(h), 0, 16, 7, 0, 0) == -1                 # bad dtype
    assert L.ts_tokstore_create(ctypes.byref(h), 0, -4, _lib.TS_BF16, 0, 0) == -1
    assert b"invalid" in L.ts_last_error()
    assert L.ts_index_ntotal(None) == -1


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tristage_rag_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text, f
