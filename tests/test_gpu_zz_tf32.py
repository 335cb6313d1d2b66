"""GPU: fp32 storage on the tensor path (kind::tf32; csrc/s1_umma.cu) against the oracle -- the default for an
fp32 index and B > 4 (TS_TF32=0 keeps the exact-fp32 CUDA-core scan, which batches <= 4 always take).  The same
cases run on the emulator in tests/test_cudasim.py::test_tensor_scan_over_fp32_storage_reads_tf32.

Tolerance: against the oracle on tf32-rounded operands the Stage-1 rule (1e-3 relative); against the fp32
oracle (what FAISS computes) 1e-3 relative + 1e-3 / sqrt(d) absolute (tf32 keeps 10 mantissa bits: ~2^-11 / sqrt(d)
on unit vectors), ids equal outside that band."""
import os

import numpy as np
import pytest

from test_cudasim import make, oracle_search

from oracle import flat_ip
from tristage_rag_b200 import _lib

pytestmark = pytest.mark.gpu
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
    aD, aI = idx.search_host(Q, k)                               # auto: the tensor path for B > 4, the exact scan below
    if B > 4:
        assert (aI == I).all() and (aD == D).all()
    else:
        assert not flat_ip.check_topk(aD, aI, fsc, fD, fI, rel=REL)


def test_config2_fp32_corpus_full_size(cuda_device):
    """BASELINE configs[1] with the reference's own storage dtype: 1 M x 768 fp32, B = 32 on the tf32 tensor path
    -- planted rows recovered, scores within the tf32 rule of the exact fp32 oracle over the first 200 k rows."""
    import torch

    from test_gpu_zzz_fullsize import check_common

    N, d, k, B = 1_000_000, 768, 100, 32
    dev = torch.device("cuda", cuda_device)
    g = torch.Generator(device=dev).manual_seed(21)
    q = torch.randn((B, d), generator=g, device=dev)
    q /= q.norm(dim=1, keepdim=True) + 1e-8
    rng = np.random.default_rng(21)
    pos = rng.choice(N, size=200, replace=False).reshape(2, 100)
    pos_t = torch.from_numpy(pos).to(dev)
    idx = _lib.Index(d, "fp32", "ip", cuda_device, reserve_rows=N)
    for s in range(0, N, 250_000):
        x = torch.randn((250_000, d), generator=g, device=dev)
        for b in range(2):
            sel = pos_t[b][(pos_t[b] >= s) & (pos_t[b] < s + 250_000)] - s
            if len(sel):
                x[sel] = q[b][None, :] + 0.7 * torch.randn((len(sel), d), generator=g, device=dev) / d ** 0.5
        x /= x.norm(dim=1, keepdim=True) + 1e-8
        idx.add(x)
    s, i = idx.search(q, k)
    torch.cuda.synchronize()
    sn, inn = check_common(s, i, N)
    for b in range(2):
        assert set(inn[b].tolist()) == set(pos[b].tolist())
    X = idx.get_rows(0, 200_000)
    Qh = q.cpu().numpy()
    S = Qh[2:6] @ X.T
    for j, b in enumerate(range(2, 6)):
        inside = inn[b] < 200_000
        ref = S[j, inn[b][inside]]
        assert (np.abs(sn[b][inside] - ref) <= REL * np.abs(ref) + 1e-3 / np.sqrt(d)).all()
        kth = float(sn[b, -1])
        assert not (S[j] > kth + 2 * (REL * abs(kth) + 1e-3 / np.sqrt(d))).sum() > int(inside.sum())
