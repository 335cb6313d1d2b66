"""GPU parity: Stage-1 exact top-k through the C ABI vs the CPU oracle.

Rule (BASELINE.json north_star / SURVEY.md §8d): id lists identical to the
oracle's except swaps inside a 1e-3 relative near-tie band; scores within
1e-3 relative.  The oracle scores the SAME storage-rounded values (bf16/fp16
rows and queries rounded like the kernels round them) in fp32.
"""
import os
import tempfile

import numpy as np
import pytest
import torch

from oracle import flat_ip
from tristage_rag_b200 import _lib

pytestmark = pytest.mark.gpu
REL = 1e-3


def make(N, d, B, seed=0, planted=0):
    rng = np.random.default_rng(seed)
    X = flat_ip.normalize_rows(rng.standard_normal((N, d)).astype(np.float32)).astype(np.float32)
    Q = flat_ip.normalize_rows(rng.standard_normal((B, d)).astype(np.float32)).astype(np.float32)
    if planted and N > planted * B:
        for b in range(B):
            pos = rng.choice(N, size=planted, replace=False)
            noise = rng.standard_normal((planted, d)).astype(np.float32) / np.sqrt(d)
            X[pos] = flat_ip.normalize_rows(Q[b][None, :] + rng.choice([0.5, 1.0]) * noise)
    return X, Q


def oracle_search(X, Q, k, dtype):
    Xr, Qr = flat_ip.round_to(X, dtype), flat_ip.round_to(Q, dtype)
    idx = flat_ip.IndexFlatIP(X.shape[1])
    idx.add(Xr)
    D, I = idx.search(Qr, k)
    return D, I, (lambda b, ids: (Xr[ids].astype(np.float64) @ Qr[b].astype(np.float64)))


def run_case(cuda_device, N, d, B, k, dtype, path, seed=0, planted=0, chunks=1):
    X, Q = make(N, d, B, seed, planted)
    idx = _lib.Index(d, dtype, "ip", cuda_device)
    for part in np.array_split(X, chunks):
        idx.add(part)
    assert idx.ntotal == N
    D, I = idx.search_host(Q, k, path=path)
    rD, rI, sc = oracle_search(X, Q, k, dtype)
    bad = flat_ip.check_topk(D, I, sc, rD, rI, rel=REL)
    assert not bad, f"N={N} d={d} B={B} k={k} {dtype} {path}: {bad[:4]}"
    return idx, X, Q, (D, I)


@pytest.mark.parametrize("N,d,B,k", [(5, 768, 1, 50), (1000, 768, 1, 100), (4099, 64, 3, 10), (20000, 128, 4, 100),
                                     (3000, 100, 2, 7), (9000, 1024, 4, 128), (7000, 256, 1, 500)])
@pytest.mark.parametrize("dtype", ["bf16", "fp16", "fp32"])
def test_stream_path(cuda_device, N, d, B, k, dtype):
    run_case(cuda_device, N, d, B, k, dtype, "stream", seed=N + B)


@pytest.mark.parametrize("N,d,B,k", [(5, 768, 1, 50), (256, 64, 8, 5), (1000, 768, 32, 100), (20000, 128, 32, 100),
                                     (4099, 64, 5, 10), (3000, 100, 17, 7), (30000, 1024, 64, 100),
                                     (30000, 768, 65, 100), (25000, 256, 128, 128), (40000, 128, 200, 100),
                                     (12000, 256, 33, 500), (50000, 64, 1024, 100), (300, 1024, 1, 100),
                                     (40000, 128, 16, 500), (70000, 64, 300, 128), (38000, 128, 1, 1)])
@pytest.mark.parametrize("dtype", ["bf16", "fp16"])
def test_umma_path(cuda_device, N, d, B, k, dtype):
    if dtype == "fp16" and N > 20000:
        pytest.skip("fp16 covered at the smaller sizes")
    run_case(cuda_device, N, d, B, k, dtype, "umma", seed=N + B)


def test_planted_neighbours_and_path_agreement(cuda_device):
    """Non-degenerate top-k (SURVEY.md §8d 'planted' variant); both kernels and
    the auto dispatch agree with the oracle and with each other."""
    N, d, B, k = 60000, 768, 4, 100
    idx, X, Q, (D1, I1) = run_case(cuda_device, N, d, B, k, "bf16", "stream", seed=7, planted=100)
    D2, I2 = idx.search_host(Q, k, path="umma")
    D3, I3 = idx.search_host(Q, k, path="auto")
    rD, rI, sc = oracle_search(X, Q, k, "bf16")
    assert not flat_ip.check_topk(D2, I2, sc, rD, rI, rel=REL)
    assert (I3 == I2).all() and (D3 == D2).all()     # auto == umma for 16-bit storage, bit-identical rerun
    assert rD[:, 0].min() > 0.5                       # planted rows dominate


def test_exact_duplicates_tie_break_by_id(cuda_device):
    rng = np.random.default_rng(1)
    d = 64
    X = flat_ip.normalize_rows(rng.standard_normal((500, d)).astype(np.float32)).astype(np.float32)
    for dup in (17, 250, 499):
        X[dup] = X[3]
    Q = X[3:4].copy()
    for path in ("stream", "umma"):
        idx = _lib.Index(d, "bf16", "ip", cuda_device)
        idx.add(X)
        D, I = idx.search_host(Q, 10, path=path)
        assert I[0, :4].tolist() == [3, 17, 250, 499], path
        assert (D[0, :4] == D[0, 0]).all()


def test_multi_add_growth_reset_and_padding(cuda_device):
    idx, X, Q, _ = run_case(cuda_device, 5000, 96, 2, 20, "bf16", "auto", seed=3, chunks=7)
    back = idx.get_rows(100, 50)
    np.testing.assert_array_equal(back, flat_ip.round_to(X[100:150], "bf16"))
    idx.reset()
    assert idx.ntotal == 0
    with pytest.raises(_lib.TristageError) as e:
        idx.search_host(Q, 5)
    assert e.value.code == _lib.TS_ERR_EMPTY and "No documents indexed" in str(e.value)
    idx.add(X[:3])
    D, I = idx.search_host(Q, 8)
    assert (I[:, 3:] == -1).all() and (I[:, :3] >= 0).all()
    assert (D[:, 3:] == flat_ip.LOWEST_F32).all()


def test_normalize_flags_match_reference_formula(cuda_device):
    """ingest + query normalisation on the device == numpy x/(|x|+1e-8)."""
    rng = np.random.default_rng(5)
    X = (rng.standard_normal((3000, 128)) * 3).astype(np.float32)
    Q = (rng.standard_normal((3, 128)) * 0.2).astype(np.float32)
    idx = _lib.Index(128, "fp32", "ip", cuda_device)
    idx.add(X, normalize=True)
    np.testing.assert_allclose(idx.get_rows(0, 3000), flat_ip.normalize_rows(X), rtol=2e-6, atol=1e-7)
    D, I = idx.search_host(Q, 10, normalize_q=True)
    Xn, Qn = flat_ip.normalize_rows(X).astype(np.float32), flat_ip.normalize_rows(Q).astype(np.float32)
    rD, rI = flat_ip.topk_desc(Qn @ Xn.T, 10)
    assert not flat_ip.check_topk(D, I, lambda b, ids: Xn[ids].astype(np.float64) @ Qn[b].astype(np.float64), rD, rI, rel=1e-5)


@pytest.mark.parametrize("path,B", [("stream", 2), ("umma", 40)])
def test_cosine_metric_fuses_norm_scaling(cuda_device, path, B):
    """METRIC_COSINE: rows stored un-normalised, 1/(|x|+1e-8) applied in the scan epilogue."""
    rng = np.random.default_rng(9)
    d = 128
    X = (rng.standard_normal((8000, d)) * rng.uniform(0.1, 5.0, size=(8000, 1))).astype(np.float32)
    Q = flat_ip.normalize_rows(rng.standard_normal((B, d)).astype(np.float32)).astype(np.float32)
    idx = _lib.Index(d, "bf16", "cosine", cuda_device)
    idx.add(X)
    D, I = idx.search_host(Q, 50, path=path)
    Xr, Qr = flat_ip.round_to(X, "bf16"), flat_ip.round_to(Q, "bf16")
    inv = (1.0 / (np.linalg.norm(X, axis=1) + 1e-8)).astype(np.float32)
    S = (Qr @ Xr.T) * inv[None, :]
    rD, rI = flat_ip.topk_desc(S, 50)
    sc = lambda b, ids: (Xr[ids].astype(np.float64) @ Qr[b].astype(np.float64)) * inv[ids]   # noqa: E731
    assert not flat_ip.check_topk(D, I, sc, rD, rI, rel=REL)


def test_save_load_roundtrip(cuda_device):
    idx, X, Q, (D, I) = run_case(cuda_device, 4000, 72, 3, 25, "bf16", "auto", seed=11)
    with tempfile.TemporaryDirectory() as tmp:
        p = os.path.join(tmp, "shard.tsidx")
        idx.save(p)
        idx2 = _lib.Index.load(p, cuda_device)
        assert idx2.ntotal == 4000
        D2, I2 = idx2.search_host(Q, 25)
        assert (I2 == I).all() and (D2 == D).all()
    with pytest.raises(_lib.TristageError):
        _lib.Index.load("/nonexistent/file", cuda_device)


def test_device_pointer_api_and_merge_of_shards(cuda_device):
    """Partition property: top-k over the union == merge of per-shard top-k
    (what the multi-GPU all-gather path computes), via ts_topk_merge."""
    N, d, B, k = 30000, 128, 16, 100
    X, Q = make(N, d, B, seed=21)
    dev = torch.device("cuda", cuda_device)
    full = _lib.Index(d, "bf16", "ip", cuda_device)
    full.add(X)
    q = torch.from_numpy(Q).to(dev)
    Df, If = full.search(q, k)
    parts_s, parts_i = [], []
    bounds = [0, 7001, 7002, 19000, N]               # uneven shards, one of a single row
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        sh = _lib.Index(d, "bf16", "ip", cuda_device)
        sh.add(torch.from_numpy(X[lo:hi]).to(dev))   # device-resident fp32 source
        sh.set_id_base(lo)
        s, i = sh.search(q.to(torch.bfloat16), k)    # storage-dtype queries
        parts_s.append(s)
        parts_i.append(i)
    Dm, Im = _lib.topk_merge(torch.stack(parts_s), torch.stack(parts_i), cuda_device)
    torch.cuda.synchronize()
    assert (Im == If).all() and (Dm == Df).all()
    rD, rI, sc = oracle_search(X, Q, k, "bf16")
    assert not flat_ip.check_topk(Dm.cpu().numpy(), Im.cpu().numpy(), sc, rD, rI, rel=REL)


def test_argument_errors(cuda_device):
    idx = _lib.Index(32, "bf16", "ip", cuda_device)
    idx.add(np.ones((4, 32), np.float32))
    q = np.ones((1, 32), np.float32)
    for k in (0, 513):
        with pytest.raises(_lib.TristageError):
            idx.search_host(q, k)
    f32 = _lib.Index(32, "fp32", "ip", cuda_device)
    f32.add(np.ones((4, 32), np.float32))
    with pytest.raises(KeyError):
        f32.search_host(q, 2, path="nonsense")        # (path="umma" on fp32 storage = the opt-in tf32 scan, tests/test_gpu_zz_tf32.py)
    D, I = f32.search_host(np.ones((9, 32), np.float32), 2)   # fp32 + B > 4: stream passes of 4
    assert I.shape == (9, 2) and (I >= 0).all()

