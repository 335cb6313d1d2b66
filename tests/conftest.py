import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_addoption(parser):
    parser.addoption("--emulate", action="store_true",
                     help="run gpu-marked tests on the CPU emulator build of the kernels (tests/cudasim) instead of a "
                          "B200: a development aid to debug the TESTS before GPU time is spent; tests that need torch "
                          "CUDA tensors cannot run this way (deselect them with -k)")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _enter_emulation():
    """Point the ctypes binding at build/cudasim/libtristage_cudasim.so (test infrastructure) for this session."""
    import ctypes as C
    import subprocess

    import numpy as np
    import torch

    from tristage_rag_b200 import _lib

    env = dict(os.environ)
    env.pop("CXX", None)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cudasim"), "-j", "8"], env=env,
                          stdout=subprocess.DEVNULL)
    L = C.CDLL(os.path.join(ROOT, "build", "cudasim", "libtristage_cudasim.so"))
    for name, (res, args) in _lib.SYMBOLS.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib._lib = L
    _lib._stream_ptr = lambda device: None

    def rank_desc(scores, top_k, n_cand=None, device=0):      # host tensors in, host tensors out
        s = np.ascontiguousarray(scores.numpy(), np.float32)
        B, Cn = s.shape
        out_s, out_p = np.empty((B, top_k), np.float32), np.empty((B, top_k), np.int32)
        nc = np.ascontiguousarray(n_cand.numpy(), np.int32) if n_cand is not None else None
        _lib.check(L.ts_rank_desc(device, C.c_void_p(s.ctypes.data), C.c_void_p(nc.ctypes.data) if nc is not None else None,
                                  B, Cn, int(top_k), C.c_void_p(out_s.ctypes.data), C.c_void_p(out_p.ctypes.data), None))
        return torch.from_numpy(out_s), torch.from_numpy(out_p)

    _lib.rank_desc = rank_desc
    real_to = torch.Tensor.to

    def to(self, *a, **kw):                                    # "device" memory is host memory here
        a = tuple(x for x in a if not (isinstance(x, torch.device) and x.type == "cuda"))
        kw = {k: v for k, v in kw.items() if not (k == "device" and str(v).startswith("cuda"))}
        return real_to(self, *a, **kw) if (a or kw) else self

    torch.Tensor.to = to


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def cuda_device(request):
    """GPU tests fail (not skip) when the native path cannot run: a silent
    fallback would void the parity claim."""
    import torch

    if request.config.getoption("--emulate"):
        _enter_emulation()
        return 0
    assert torch.cuda.is_available(), "gpu-marked test but torch sees no CUDA device"
    from tristage_rag_b200 import _lib

    assert _lib.lib().ts_device_count() >= 1, "libtristage sees no sm_100 device"
    return 0


@pytest.fixture(scope="session")
def sim_lib():
    """The kernels' own sources compiled for the CPU emulator (tests/cudasim): test infrastructure, loaded
    explicitly by the emulator tests and never by the package."""
    import ctypes as C
    import subprocess

    from tristage_rag_b200 import _lib

    env = dict(os.environ)
    env.pop("CXX", None)
    out = subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cudasim"), "-j", "8"], env=env,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert out.returncode == 0, out.stdout[-4000:]
    L = C.CDLL(os.path.join(ROOT, "build", "cudasim", "libtristage_cudasim.so"))
    assert L.hostsim_is_simulation() == 2
    for name, (res, args) in _lib.SYMBOLS.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    for name in ("cudasim_launches", "cudasim_blocks", "cudasim_switches"):
        getattr(L, name).restype = C.c_ulonglong
    return L


@pytest.fixture()
def sim(sim_lib, monkeypatch):
    from tristage_rag_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", sim_lib)
    monkeypatch.setattr(_lib, "_stream_ptr", lambda device: None)
    # the emulator tests compare each scan variant with the plain single-CTA two-launch scan: start from that,
    # whatever the product's build-time defaults are (csrc/ts_internal.h); a test enables a variant with "1"
    for var in ("TS_PAIR", "TS_FUSE", "TS_TF32"):
        monkeypatch.setenv(var, "0")
    return sim_lib
