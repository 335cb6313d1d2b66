"""GPU: fp32 storage on the tensor path (kind::tf32; csrc/s1_umma.cu) against the oracle.  The variant was written
after the round's GPU budget was spent and is opt-in (path="umma" on an fp32 index, or TS_TF32=1 under
TS_PATH_AUTO), so -- like the other unvalidated variants -- its hardware test runs only with
TS_TEST_EXPERIMENTAL=1 (tools/gpu/round2_variants.sh).  The same cases run on the emulator in
tests/test_cudasim.py::test_tensor_scan_over_fp32_storage_reads_tf32.

Tolerance: against the oracle on tf32-rounded operands the Stage-1 rule (1e-3 relative); against the fp32
oracle (what FAISS computes) 1e-3 relative + 1e-3 / sqrt(d) absolute (tf32 keeps 10 mantissa bits: ~2^-11 / sqrt(d)
on unit vectors), ids equal outside that band."""
import os

import numpy as np
import pytest

from test_cudasim import make, oracle_search

from oracle import flat_ip
from tristage_rag_b200 import _lib

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("TS_TEST_EXPERIMENTAL", "0") in ("", "0"),
                                 reason="tf32 tensor path is opt-in until validated on hardware (TS_TEST_EXPERIMENTAL=1)")]
REL = 1e-3


@pytest.mark.parametrize("N,d,B,k", [(100_000, 768, 32, 100), (100_000, 768, 200, 100), (50_000, 1024, 1, 100),
                                     (30_000, 100, 1024, 10), (20_000, 768, 8, 500)])
def test_fp32_storage_on_the_tensor_path(cuda_device, N, d, B, k):
    X, Q = make(N, d, B, seed=N + B, planted=10)
    idx = _lib.Index(d, "fp32", "ip", cuda_device)
    idx.add(X)
    D, I = idx.search_host(Q, k, path="umma")
    rD, rI, sc = oracle_search(X, Q, k, "tf32")
    assert not flat_ip.check_topk(D, I, sc, rD, rI, rel=REL)
    fD, fI, fsc = oracle_search(X, Q, k, "fp32")
    for b in range(B):
        ref = fsc(b, I[b])
        assert (np.abs(D[b] - ref) <= REL * np.abs(ref) + 1e-3 / np.sqrt(d)).all()
        extra = np.setdiff1d(I[b], fI[b])
        if extra.size:
            kth = float(fD[b].min())
            assert (fsc(b, extra) >= kth - 2 * (REL * abs(kth) + 1e-3 / np.sqrt(d))).all()
    sD, sI = idx.search_host(Q[:4], k, path="stream")           # the CUDA-core scan of the same index: exact fp32 products
    assert not flat_ip.check_topk(sD, sI, fsc, fD[:4], fI[:4], rel=REL)
