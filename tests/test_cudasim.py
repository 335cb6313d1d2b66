"""CPU: the CUDA-core kernels of libtristage EXECUTED on a SIMT emulator (tests/cudasim/: every
simulated thread is a fiber; __syncthreads and the *_sync warp primitives are rendezvous points)
and checked against the oracle with the rule the GPU parity tests use.  The kernel sources are
the product's own (.cu files compiled with g++ under TS_CUDASIM).  The tensor-path kernels run
too, over a functional model of the inline-PTX wrappers (tests/cudasim/ts_ptx_sim.cuh: mbarrier
phases, TMA boxes with the 128-byte swizzle, UMMA descriptors, TMEM) in which every asynchronous
operation completes at issue time -- it checks data movement, descriptor arithmetic, phase
bookkeeping, masks and the fused selection, not latency or the memory model.  Test infrastructure
only -- the emulated library is built into build/cudasim/, loaded explicitly here, and never used
by the package."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import flat_ip, maxsim
from tristage_rag_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REL = 1e-3


def make(N, d, B, seed=0, planted=0):
    rng = np.random.default_rng(seed)
    X = flat_ip.normalize_rows(rng.standard_normal((N, d)).astype(np.float32)).astype(np.float32)
    Q = flat_ip.normalize_rows(rng.standard_normal((B, d)).astype(np.float32)).astype(np.float32)
    if planted and N > planted * B:
        for b in range(B):
            pos = rng.choice(N, size=planted, replace=False)
            noise = rng.standard_normal((planted, d)).astype(np.float32) / np.sqrt(d)
            X[pos] = flat_ip.normalize_rows(Q[b][None, :] + 0.7 * noise)
    return X, Q


def oracle_search(X, Q, k, dtype):
    Xr, Qr = flat_ip.round_to(X, dtype), flat_ip.round_to(Q, dtype)
    idx = flat_ip.IndexFlatIP(X.shape[1])
    idx.add(Xr)
    D, I = idx.search(Qr, k)
    return D, I, (lambda b, ids: (Xr[ids].astype(np.float64) @ Qr[b].astype(np.float64)))


def p(a):
    return C.c_void_p(a.ctypes.data)


# ------------------------------------------------------------------- Stage 1 ---
@pytest.mark.parametrize("N,d,B,k,dtype", [(5, 40, 1, 50, "bf16"), (700, 64, 3, 10, "bf16"), (900, 100, 2, 7, "fp16"),
                                           (1500, 32, 4, 128, "fp32"), (1100, 72, 6, 20, "bf16"), (600, 16, 1, 500, "fp32")])
def test_stream_scan_and_select_match_the_oracle(sim, N, d, B, k, dtype):
    X, Q = make(N, d, B, seed=N + B, planted=5)
    idx = _lib.Index(d, dtype, "ip", 0)
    for part in np.array_split(X, 3):
        idx.add(part)
    before = sim.cudasim_launches()
    D, I = idx.search_host(Q, k, path="stream")
    assert sim.cudasim_launches() - before >= 3          # query prep, scan pass(es), selection
    rD, rI, sc = oracle_search(X, Q, k, dtype)
    assert not flat_ip.check_topk(D, I, sc, rD, rI, rel=REL)
    if k > N:
        assert (I[:, N:] == -1).all() and (D[:, N:] == np.float32(flat_ip.LOWEST_F32)).all()
    idx.set_id_base(10_000_000_000)
    D2, I2 = idx.search_host(Q, min(k, N), path="stream")
    assert (I2 == I[:, : min(k, N)] + 10_000_000_000).all()


@pytest.mark.parametrize("order", ["ascending", "descending", "constant"])
def test_adversarial_score_orders_exercise_the_in_scan_prune(sim, order):
    """Scores chosen per row: ascending order overflows every candidate list (warp bitonic prune in
    the scan), constant makes every row tie (the k smallest ids must win)."""
    N, d, k = 2600, 32, 100
    rng = np.random.default_rng(4)
    u = flat_ip.normalize_rows(rng.standard_normal((1, d)).astype(np.float32))[0].astype(np.float32)
    v = {"ascending": np.linspace(0.05, 1.0, N), "descending": np.linspace(1.0, 0.05, N),
         "constant": np.full(N, 0.5)}[order].astype(np.float32)
    X = (v[:, None] * u[None, :]).astype(np.float32)
    Q = np.stack([u, -u]).astype(np.float32)
    idx = _lib.Index(d, "bf16", "ip", 0)
    idx.add(X)
    D, I = idx.search_host(Q, k, path="stream")
    rD, rI, sc = oracle_search(X, Q, k, "bf16")
    assert not flat_ip.check_topk(D, I, sc, rD, rI, rel=REL)
    if order == "constant":
        assert I[0].tolist() == list(range(k)) and I[1].tolist() == list(range(k))


def test_cosine_metric_and_query_normalisation_flags(sim):
    rng = np.random.default_rng(2)
    N, d, k = 800, 48, 15
    X = (rng.standard_normal((N, d)) * rng.uniform(0.1, 5.0, size=(N, 1))).astype(np.float32)
    Q = (3.0 * rng.standard_normal((3, d))).astype(np.float32)
    idx = _lib.Index(d, "bf16", "cosine", 0)
    idx.add(X)
    D, I = idx.search_host(Q, k, normalize_q=True, path="stream")
    Xr = flat_ip.round_to(X, "bf16")
    Qr = flat_ip.round_to(flat_ip.normalize_rows(Q), "bf16")
    inv = (1.0 / (np.linalg.norm(X, axis=1) + 1e-8)).astype(np.float32)
    S = (Qr @ Xr.T) * inv[None, :]
    rD, rI = flat_ip.topk_desc(S, k)
    sc = lambda b, ids: (Xr[ids].astype(np.float64) @ Qr[b].astype(np.float64)) * inv[ids]   # noqa: E731
    assert not flat_ip.check_topk(D, I, sc, rD, rI, rel=REL)


def test_shard_merge_equals_full_search(sim):
    """ts_topk_merge / ts_topk_merge_packed: merge of per-shard top-k == top-k over the union, ties by id."""
    N, d, B, k = 1200, 32, 5, 40
    X, Q = make(N, d, B, seed=21)
    X[700] = X[3]
    X[1100] = X[3]                                           # exact duplicates across shards
    full = _lib.Index(d, "bf16", "ip", 0)
    full.add(X)
    Df, If = full.search_host(Q, k, path="stream")
    bounds = [0, 401, 402, 900, N]
    S = np.empty((4, B, k), np.float32)
    Id = np.empty((4, B, k), np.int64)
    for i, (lo, hi) in enumerate(zip(bounds[:-1], bounds[1:])):
        sh = _lib.Index(d, "bf16", "ip", 0)
        sh.add(X[lo:hi])
        sh.set_id_base(lo)
        S[i], Id[i] = sh.search_host(Q, k, path="stream")    # 1-row shard: k - 1 slots are -1 padding
    out_s, out_i = np.empty((B, k), np.float32), np.empty((B, k), np.int64)
    _lib.check(sim.ts_topk_merge(0, p(S), p(Id), 4, B, k, p(out_s), p(out_i), None))
    assert (out_i == If).all() and (out_s == Df).all()
    ids_off, nbytes = _lib.packed_layout(B, k)
    blob = np.zeros(4 * nbytes, np.uint8)
    for i in range(4):
        blob[i * nbytes: i * nbytes + B * k * 4] = S[i].view(np.uint8).ravel()
        blob[i * nbytes + ids_off: i * nbytes + ids_off + B * k * 8] = Id[i].view(np.uint8).ravel()
    out_s2, out_i2 = np.empty((B, k), np.float32), np.empty((B, k), np.int64)
    _lib.check(sim.ts_topk_merge_packed(0, p(blob), nbytes, ids_off, 4, B, k, p(out_s2), p(out_i2), None))
    assert (out_i2 == If).all() and (out_s2 == Df).all()
    rD, rI, sc = oracle_search(X, Q, k, "bf16")
    assert not flat_ip.check_topk(out_s, out_i, sc, rD, rI, rel=REL)


def test_rank_desc_is_the_reference_stable_sort(sim):
    rng = np.random.default_rng(5)
    B, Cn, top = 4, 300, 100
    scores = rng.integers(0, 12, size=(B, Cn)).astype(np.float32) / 4          # many exact ties
    n_cand = np.array([300, 17, 0, 100], np.int32)
    out_s, out_p = np.empty((B, top), np.float32), np.empty((B, top), np.int32)
    _lib.check(sim.ts_rank_desc(0, p(scores), p(n_cand), B, Cn, top, p(out_s), p(out_p), None))
    for b in range(B):
        order = maxsim.rescore_order(scores[b, : n_cand[b]], top)
        n = len(order)
        assert out_p[b, :n].tolist() == order.tolist() and (out_p[b, n:] == -1).all()
        assert (out_s[b, :n] == scores[b, order]).all()


# ------------------------------------------- Stage 1, tensor path (tcgen05 / TMA model) ---
@pytest.mark.parametrize("N,d,B,k,dtype", [(5, 768, 1, 50, "bf16"), (256, 64, 8, 5, "bf16"), (1000, 768, 32, 100, "bf16"),
                                           (9000, 128, 32, 100, "bf16"), (4099, 64, 5, 10, "fp16"), (3000, 100, 17, 7, "bf16"),
                                           (3000, 1024, 64, 100, "bf16"), (4000, 768, 65, 100, "bf16"), (6000, 256, 128, 128, "fp16"),
                                           (9000, 128, 200, 100, "bf16"), (5000, 256, 33, 500, "bf16"), (20000, 64, 1024, 100, "bf16"),
                                           (300, 1024, 1, 100, "bf16"), (9000, 128, 16, 500, "bf16"), (12000, 64, 300, 128, "bf16"),
                                           (38000, 128, 1, 1, "bf16")])
def test_tensor_scan_matches_the_oracle(sim, N, d, B, k, dtype):
    """s1_umma_kernel (threshold pre-pass + scan with the fused, threshold-sharing top-k) and
    select_kernel over its lists, executed on the functional tcgen05 / TMA / mbarrier model; the
    shapes of tests/test_gpu_stage1.py::test_umma_path at reduced corpus sizes."""
    X, Q = make(N, d, B, seed=N + B, planted=10 if N > 10 * B else 0)
    idx = _lib.Index(d, dtype, "ip", 0)
    for part in np.array_split(X, 2):
        idx.add(part)
    D, I = idx.search_host(Q, k, path="umma")
    rD, rI, sc = oracle_search(X, Q, k, dtype)
    assert not flat_ip.check_topk(D, I, sc, rD, rI, rel=REL)
    Da, Ia = idx.search_host(Q, k)                                  # auto = tensor path for 16-bit storage
    assert (Ia == I).all() and (Da == D).all()


@pytest.mark.parametrize("N,d,B,k", [(5, 768, 1, 50), (3000, 100, 17, 7), (1500, 768, 65, 100), (2500, 36, 130, 20),
                                     (1500, 64, 8, 500)])
def test_tensor_scan_over_fp32_storage_reads_tf32(sim, monkeypatch, N, d, B, k):
    """fp32 storage on the tensor path (kind::tf32, TFLOAT32 tensor maps, 32-element K chunks): scores are those of
    the tf32-rounded operands with fp32 accumulation; ids / scores within the Stage-1 tolerance of the fp32 oracle
    as well.  TS_PATH_AUTO keeps the CUDA-core scan for fp32 storage unless TS_TF32=1 and B > 4."""
    X, Q = make(N, d, B, seed=N + B, planted=10 if N > 10 * B else 0)
    idx = _lib.Index(d, "fp32", "ip", 0)
    for part in np.array_split(X, 2):
        idx.add(part)
    D, I = idx.search_host(Q, k, path="umma")
    rD, rI, sc = oracle_search(X, Q, k, "tf32")
    assert not flat_ip.check_topk(D, I, sc, rD, rI, rel=REL)           # against the tf32-rounded operands
    # against the fp32 reference (what FAISS computes): two 10-bit-mantissa operands cost ~2^-11 / sqrt(d) ABSOLUTE on
    # unit vectors (6e-5 at d = 64, 2e-5 at d = 768) on top of the relative rule; ids may differ only inside that band
    fD, fI, fsc = oracle_search(X, Q, k, "fp32")
    for b in range(B):
        ok = I[b] >= 0
        assert (ok == (fI[b] >= 0)).all()
        ref = fsc(b, I[b][ok])
        assert (np.abs(D[b][ok] - ref) <= REL * np.abs(ref) + 1e-3 / np.sqrt(d)).all()
        extra = np.setdiff1d(I[b][ok], fI[b][ok])
        if extra.size:
            kth = float(fD[b][ok].min())
            assert (fsc(b, extra) >= kth - 2 * (REL * abs(kth) + 1e-3 / np.sqrt(d))).all()
    sD, sI = idx.search_host(Q[:4], k)                                 # auto: CUDA-core scan, exact fp32 products
    assert not flat_ip.check_topk(sD, sI, fsc, fD[:4], fI[:4], rel=REL)
    monkeypatch.setenv("TS_TF32", "1")
    if B > 4:
        aD, aI = idx.search_host(Q, k)                                 # auto now takes the tensor path
        assert (aI == I).all() and (aD == D).all()
    aD, aI = idx.search_host(Q[:4], k)                                 # ... but not for one CUDA-core pass
    assert (aI == sI).all() and (aD == sD).all()


@pytest.mark.parametrize("order", ["ascending", "descending", "constant"])
@pytest.mark.parametrize("k", [100, 500])
def test_tensor_scan_adversarial_score_orders(sim, order, k):
    """The cases of tests/test_gpu_zzz_fullsize.py::test_adversarial_score_orders_on_device: ascending
    scores overflow every list (in-scan warp prune under a stale shared bound), constant scores tie
    everywhere (the k smallest ids must win through scan, prune and select)."""
    N, d = 60_000, 64
    rng = np.random.default_rng(4)
    u = flat_ip.normalize_rows(rng.standard_normal((1, d)).astype(np.float32))[0].astype(np.float32)
    v = {"ascending": np.linspace(0.05, 1.0, N), "descending": np.linspace(1.0, 0.05, N),
         "constant": np.full(N, 0.5)}[order].astype(np.float32)
    X = (v[:, None] * u[None, :]).astype(np.float32)
    Q = np.stack([u, -u, flat_ip.normalize_rows(rng.standard_normal((1, d)).astype(np.float32))[0]]).astype(np.float32)
    idx = _lib.Index(d, "bf16", "ip", 0)
    idx.add(X)
    D, I = idx.search_host(Q, k, path="umma")
    rD, rI, sc = oracle_search(X, Q, k, "bf16")
    bad = flat_ip.check_topk(D, I, sc, rD, rI, rel=REL)
    assert not bad, bad[:3]
    if order == "constant":
        assert I[0].tolist() == list(range(k)) and I[1].tolist() == list(range(k))


def test_tensor_scan_variants_agree_bit_for_bit(sim, monkeypatch):
    """tests/test_gpu_zzz_fullsize.py::test_scan_variants_agree_bit_for_bit on the emulator: without
    threshold sharing, without the small-batch spread, and
    with the first select kernel -- identical ids and scores."""
    N, d, B, k = 20000, 128, 48, 100
    X, Q = make(N, d, B, seed=77, planted=20)
    idx = _lib.Index(d, "bf16", "ip", 0)
    idx.add(X)
    base = idx.search_host(Q, k, path="umma")
    rD, rI, sc = oracle_search(X, Q, k, "bf16")
    assert not flat_ip.check_topk(base[0], base[1], sc, rD, rI, rel=REL)
    Q2 = make(10, d, 300, seed=78)[1]
    base2 = idx.search_host(Q2, k, path="umma")
    for var in ("TS_DBG_NOSHARE", "TS_DBG_NOSPREAD", "TS_SELECT_V1"):
        monkeypatch.setenv(var, "1")
        D, I = idx.search_host(Q, k, path="umma")
        D2, I2 = idx.search_host(Q2, k, path="umma")
        monkeypatch.delenv(var)
        assert (I == base[1]).all() and (D == base[0]).all(), var
        assert (I2 == base2[1]).all() and (D2 == base2[0]).all(), var
    # TS_FUSE: pre-pass + grid barrier + scan in ONE cooperative launch.  Every CTA must be alive at once,
    # so the emulated "GPU" is shrunk to 12 SMs (12 x 192 fibers); the 148-SM layout of `base` differs, so
    # the comparison is against the two-launch result on the same 12 SMs.
    monkeypatch.setenv("HOSTSIM_SM_COUNT", "12")
    small = _lib.Index(d, "bf16", "ip", 0)
    small.add(X)
    two = small.search_host(Q, k, path="umma")
    two2 = small.search_host(Q2, k, path="umma")
    assert (two[1] == base[1]).all() and (two[0] == base[0]).all()          # the SM count never changes the answer
    monkeypatch.setenv("TS_FUSE", "1")
    before = sim.cudasim_launches()
    D, I = small.search_host(Q, k, path="umma")
    assert sim.cudasim_launches() - before == 3                             # query prep, ONE scan launch, select
    assert (I == two[1]).all() and (D == two[0]).all()
    D2, I2 = small.search_host(Q2, k, path="umma")                          # three query tiles x 4 slices
    assert (I2 == two2[1]).all() and (D2 == two2[0]).all()
    D, I = small.search_host(Q, k, path="umma")                             # the barrier resets itself: run it again
    assert (I == two[1]).all() and (D == two[0]).all()


@pytest.mark.parametrize("sms,B,k,order", [(12, 48, 10, "random"), (16, 5, 16, "random"), (12, 128, 12, "ascending"), (40, 33, 7, "constant")])
def test_kth_of_slices_bound_never_drops_a_result(sim, monkeypatch, sms, B, k, order):
    """kth_rule (topk_select.cu): with one published best per slice and n_slices >= k the select kernel filters the
    candidate lists with the k-th largest of the slices' bests instead of their minimum.  The result must be the
    oracle's and bit-equal to the result without the rule, after the two-launch and the single-launch (TS_FUSE)
    scan, for random rows, ascending scores and all-equal scores (ties exactly at the bound)."""
    monkeypatch.setenv("HOSTSIM_SM_COUNT", str(sms))
    N, d = 9000, 64
    if order == "random":
        X, Q = make(N, d, B, seed=sms + B, planted=6)
    else:
        rng = np.random.default_rng(4)
        u = flat_ip.normalize_rows(rng.standard_normal((1, d)).astype(np.float32))[0].astype(np.float32)
        v = (np.linspace(0.05, 1.0, N) if order == "ascending" else np.full(N, 0.5)).astype(np.float32)
        X = (v[:, None] * u[None, :]).astype(np.float32)
        Q = np.concatenate([u[None, :], flat_ip.normalize_rows(rng.standard_normal((B - 1, d)).astype(np.float32))]).astype(np.float32)
    idx = _lib.Index(d, "bf16", "ip", 0)
    idx.add(X)
    rD, rI, sc = oracle_search(X, Q, k, "bf16")
    monkeypatch.setenv("TS_DBG_NOKTH", "1")
    D0, I0 = idx.search_host(Q, k, path="umma")
    monkeypatch.delenv("TS_DBG_NOKTH")
    for fuse in ("0", "1"):
        monkeypatch.setenv("TS_FUSE", fuse)
        D, I = idx.search_host(Q, k, path="umma")
        assert not flat_ip.check_topk(D, I, sc, rD, rI, rel=REL), (fuse, order)
        assert (I == I0).all() and (D == D0).all(), (fuse, order)
        if order == "constant":
            assert I[0].tolist() == list(range(k))


@pytest.mark.parametrize("N,d,B,k,dtype", [(3000, 64, 200, 10, "bf16"), (5000, 128, 300, 100, "bf16"), (600, 100, 129, 7, "fp16"),
                                           (6000, 64, 1024, 100, "bf16"), (4000, 256, 256, 500, "bf16"), (5, 72, 130, 50, "bf16")])
def test_cta_pair_scan_equals_the_single_cta_scan(sim, monkeypatch, N, d, B, k, dtype):
    """s1_pair_kernel (TS_PAIR=1: tcgen05.mma cta_group::2 over a 2-CTA cluster, each CTA loading half of
    every corpus chunk, remote mbarrier arrives, multicast commits) on the emulator's cluster model:
    same ids and scores as the single-CTA scan, bit for bit, and parity with the oracle.  Odd numbers
    of query tiles (B = 300, 129, 130) exercise the padded idle tile."""
    X, Q = make(N, d, B, seed=N + B, planted=5 if N > 5 * B else 0)
    idx = _lib.Index(d, dtype, "ip" if N > 10 else "cosine", 0)
    idx.add(X)
    D0, I0 = idx.search_host(Q, k, path="umma")
    monkeypatch.setenv("TS_PAIR", "1")
    D, I = idx.search_host(Q, k, path="umma")
    assert (I == I0).all() and (D == D0).all()
    if N > 10:
        rD, rI, sc = oracle_search(X, Q, k, dtype)
        assert not flat_ip.check_topk(D, I, sc, rD, rI, rel=REL)
    monkeypatch.setenv("TS_DBG_NOSHARE", "1")                       # no pre-pass launch, private thresholds only
    D2, I2 = idx.search_host(Q, k, path="umma")
    assert (I2 == I0).all() and (D2 == D0).all()


def test_tensor_scan_cosine_metric_ties_and_shards(sim):
    rng = np.random.default_rng(2)
    N, d, k = 3000, 72, 50
    X = (rng.standard_normal((N, d)) * rng.uniform(0.1, 5.0, size=(N, 1))).astype(np.float32)
    Q = (3.0 * rng.standard_normal((4, d))).astype(np.float32)
    idx = _lib.Index(d, "bf16", "cosine", 0)
    idx.add(X)
    D, I = idx.search_host(Q, k, normalize_q=True, path="umma")
    Xr = flat_ip.round_to(X, "bf16")
    Qr = flat_ip.round_to(flat_ip.normalize_rows(Q), "bf16")
    inv = (1.0 / (np.linalg.norm(X, axis=1) + 1e-8)).astype(np.float32)
    rD, rI = flat_ip.topk_desc((Qr @ Xr.T) * inv[None, :], k)
    sc = lambda b, ids: (Xr[ids].astype(np.float64) @ Qr[b].astype(np.float64)) * inv[ids]   # noqa: E731
    assert not flat_ip.check_topk(D, I, sc, rD, rI, rel=REL)
    # exact duplicates: ties resolved by ascending id; half shards merged == full scan (what 2 GPUs compute)
    X2, Q2 = make(2000, 64, 6, seed=3)
    X2[5] = X2[1500] = X2[900] = X2[17]
    full = _lib.Index(64, "bf16", "ip", 0)
    full.add(X2)
    Df, If = full.search_host(Q2, 20, path="umma")
    S, Id = np.empty((2, 6, 20), np.float32), np.empty((2, 6, 20), np.int64)
    for i, (lo, hi) in enumerate(((0, 1000), (1000, 2000))):
        sh = _lib.Index(64, "bf16", "ip", 0)
        sh.add(X2[lo:hi])
        sh.set_id_base(lo)
        S[i], Id[i] = sh.search_host(Q2, 20, path="umma")
    out_s, out_i = np.empty((6, 20), np.float32), np.empty((6, 20), np.int64)
    _lib.check(sim.ts_topk_merge(0, p(S), p(Id), 2, 6, 20, p(out_s), p(out_i), None))
    assert (out_i == If).all() and (out_s == Df).all()
    rD2, rI2, sc2 = oracle_search(X2, Q2, 20, "bf16")
    assert (If == rI2).all()


# ------------------------------------------------------------------- Stage 2 ---
@pytest.mark.parametrize("dtype,mode", [("bf16", _lib.TS_S2_MAXSIM), ("fp32", _lib.TS_S2_COLBERT), ("fp16", _lib.TS_S2_MAXSIM)])
def test_simt_maxsim_matches_the_oracle(sim, dtype, mode):
    rng = np.random.default_rng(8)
    dim, ndocs, Bq, Cn, Lq = 24, 60, 3, 12, 9
    lens = rng.integers(1, 40, size=ndocs)
    tok = (rng.standard_normal((int(lens.sum()), dim)) * 2).astype(np.float32)
    st = _lib.TokStore(dim, dtype, 0)
    st.add(tok, lens, normalize=True)
    st.set_id_base(100)
    q = rng.standard_normal((Bq, Lq, dim)).astype(np.float32)
    cand = (rng.integers(0, ndocs, size=(Bq, Cn)) + 100).astype(np.int64)
    cand[0, 3] = 5            # below the shard's id range
    cand[1, 0] = 100 + ndocs  # above it
    cand[2, 5] = -1
    q_len = np.array([9, 4, 1], np.int32)
    n_cand = np.array([12, 7, 12], np.int32)
    got = st.maxsim_host(q, cand, q_len=q_len, n_cand=n_cand, mode=mode)
    off = np.concatenate([[0], np.cumsum(lens)])
    nr = lambda x: flat_ip.round_to(maxsim.l2_normalize_tokens(x), dtype)      # noqa: E731
    fn = maxsim.maxsim_score if mode == _lib.TS_S2_MAXSIM else maxsim.colbert_score
    for b in range(Bq):
        for j in range(Cn):
            c = cand[b, j] - 100
            if j >= n_cand[b] or c < 0 or c >= ndocs:
                assert got[b, j] == 0.0
            else:
                want = fn(nr(q[b, : q_len[b]]), nr(tok[off[c]: off[c + 1]]), normalize=False)
                assert got[b, j] == pytest.approx(want, rel=1e-3, abs=2e-4)


# ------------------------------------------- selection over the tensor-path scan's lists ---
def f2ord(x):
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    return np.where(u & 0x80000000, (~u) & 0xFFFFFFFF, u | 0x80000000).astype(np.uint64)


def make_keys(scores, rows):
    return (f2ord(scores) << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - np.asarray(rows, np.uint64))


def scan_output_contract(Sc, n_slices, k, cap, rng, extra_frac=0.5):
    """What s1_umma_kernel leaves for select_kernel (csrc/s1_umma.cu header): per (slice, query) an
    UNSORTED list holding at least every row of the slice at or above the final shared bound
    min_c pub[c][q] (plus rows appended while the bound was still lower), its count, and
    pub[c][q] = the J-th best score of the slice, J = ceil(k / n_slices) (sharing off when J > 8)."""
    B, N = Sc.shape
    n_mt = (B + 127) // 128
    bpad = n_mt * 128
    spread = 1 if B <= 64 else 0
    J = -(-k // n_slices)
    jrank = J if J <= 8 else 0
    tiles = np.arange(N) // 256
    pub = np.full((n_slices, bpad), -np.inf, np.float32)
    lists = np.zeros((n_mt * n_slices * 128, cap), np.uint64)
    counts = np.zeros(n_mt * n_slices * 128, np.int32)
    members = [np.nonzero(tiles % n_slices == c)[0] for c in range(n_slices)]
    for b in range(B):
        if jrank:
            for c in range(n_slices):
                s = np.sort(Sc[b, members[c]])[::-1]
                pub[c, b] = s[J - 1] if len(s) >= J else -np.inf
        bound = pub[:, b].min() if jrank else -np.inf
        mt, qi = b >> 7, b & 127
        row = 64 + qi if spread else qi
        for c in range(n_slices):
            rows = members[c]
            must = rows[Sc[b, rows] >= bound]
            rest = rows[Sc[b, rows] < bound]
            if len(must) > cap:                       # the in-scan prune keeps the best k of an overflowing list
                order = np.lexsort((must, -Sc[b, must].astype(np.float64)))
                must = must[order[:k]]
            take = rng.permutation(rest)[: min(int(len(rest) * extra_frac), cap - len(must))]
            sel = rng.permutation(np.concatenate([must, take]))
            at = (mt * n_slices + c) * 128 + row
            lists[at, : len(sel)] = make_keys(Sc[b, sel], sel)
            lists[at, len(sel):] = rng.integers(1, 2 ** 63, size=cap - len(sel)).astype(np.uint64)   # stale keys beyond count
            counts[at] = len(sel)
    return lists, counts, pub, dict(n_slices=n_slices, n_mt=n_mt, cap=cap, spread=spread, bpad=bpad, jrank=jrank)


@pytest.mark.parametrize("N,B,k,n_slices,ties", [(6000, 3, 100, 20, False), (9000, 130, 100, 35, False), (5000, 2, 500, 10, False),
                                                 (4000, 5, 10, 148, False), (3000, 2, 100, 7, True), (300, 1, 100, 2, False),
                                                 (40000, 1, 100, 148, True)])   # 14 800 tied survivors: more than the sort buffer holds
@pytest.mark.parametrize("serial", [False, True])
def test_select_over_scan_lists_matches_the_oracle(sim, monkeypatch, N, B, k, n_slices, ties, serial):
    """select_kernel in its kLists mode (parallel count prefix + batched loads by default, the first
    version with TS_SELECT_V1=1) over lists that satisfy the scan's output contract."""
    rng = np.random.default_rng(N + B + k)
    X, Q = make(N, 32, B, seed=N, planted=0)
    if ties:
        X[:] = X[0]                                   # every row identical: one giant tie, ids must ascend
    Sc = (flat_ip.round_to(Q, "bf16") @ flat_ip.round_to(X, "bf16").T).astype(np.float32)
    n_slices = min(n_slices, (N + 255) // 256)
    cap = 256 if k <= 128 else 1024
    lists, counts, pub, lay = scan_output_contract(Sc, n_slices, k, cap, rng)
    if serial:
        monkeypatch.setenv("TS_SELECT_V1", "1")
    out_s, out_i = np.empty((B, k), np.float32), np.empty((B, k), np.int64)
    sim.cudasim_merge_lists.argtypes = [C.c_void_p] * 3 + [C.c_int] * 8 + [C.c_int64, C.c_void_p, C.c_void_p]
    rc = sim.cudasim_merge_lists(p(lists), p(counts), p(pub), lay["n_slices"], lay["n_mt"], lay["cap"], lay["spread"],
                                 lay["bpad"], lay["jrank"], B, k, 7_000_000_000, p(out_s), p(out_i))
    assert rc == 0, sim.ts_last_error()
    rD, rI = flat_ip.topk_desc(Sc, k)
    valid = rI >= 0
    assert (out_i[valid] == rI[valid] + 7_000_000_000).all() and (out_i[~valid] == -1).all()
    assert (out_s[valid] == rD[valid]).all()


# ------------------------------------------------- hybrid: BM25 search + rank fusion ---
def _zipf_corpus(n_docs, vocab_size=300, seed=0):
    rng = np.random.default_rng(seed)
    vocab = [f"w{i}" for i in range(vocab_size)]
    zipf = 1.0 / np.arange(1, vocab_size + 1)
    zipf /= zipf.sum()
    return [" ".join(rng.choice(vocab, size=int(rng.integers(0, 50)), p=zipf)) for _ in range(n_docs)], vocab


def test_device_bm25_is_bit_identical_to_the_host_index(sim):
    """bm25_score_kernel + bm25_topk_kernel (emulated) == BM25Index.search, itself pinned bit for bit
    against the reference's BM25Index in tests/test_host.py: fp64 scores, stable order, zero-score tail."""
    from tristage_rag_b200.stage1_retriever import BM25Index, DeviceBM25

    docs, vocab = _zipf_corpus(1500)
    bm = BM25Index()
    bm.fit(docs)
    dev = DeviceBM25(bm, 0)
    queries = ["w0 w1 w2", "w7 w7 w250", "w299", "w5 unknownword w5 w0", " ".join(vocab[:30]), "", "nothing known here",
               "w0 " * 12]
    for top_k in (1, 10, 300, 1024):
        got = dev.search_batch(queries, top_k)
        for q, g in zip(queries, got):
            assert g == bm.search(q, top_k), (q, top_k)
    # 2600 touched documents for one query: the selection walks several buffer refills
    many = [" ".join(["w0", f"w{i % 40 + 1}"] * 3) for i in range(2600)]
    bm2 = BM25Index()
    bm2.fit(many)
    dev2 = DeviceBM25(bm2, 0)
    for top_k in (5, 700):
        assert dev2.search_batch(["w0 w3", "w17"], top_k) == [bm2.search("w0 w3", top_k), bm2.search("w17", top_k)]
    # documents split over several CTAs per query (grid (B, R)): same sums, same order
    import os as _os
    for split in ("3", "7", "64"):
        _os.environ["TS_BM25_SPLIT"] = split
        try:
            for q, g in zip(queries, dev.search_batch(queries, 50)):
                assert g == bm.search(q, 50), (q, split)
            assert dev2.search_batch(["w0 w3", "w17"], 700) == [bm2.search("w0 w3", 700), bm2.search("w17", 700)]
        finally:
            del _os.environ["TS_BM25_SPLIT"]
    # tiny corpus, top_k beyond it: the list just ends (ids -1 are dropped)
    bm3 = BM25Index()
    bm3.fit(["a b", "b c", "zz"])
    assert DeviceBM25(bm3, 0).search_batch(["b", "q"], 10) == [bm3.search("b", 10), bm3.search("q", 10)]


def test_device_bm25_refuses_non_positive_weights(sim):
    from tristage_rag_b200.stage1_retriever import BM25Index, DeviceBM25

    bm = BM25Index()
    bm.fit(["a a b", "a c", "a d"])
    bm._build_postings()
    bm._w[0] = 0.0                                     # what a stale refit can produce (idf <= 0)
    with pytest.raises(_lib.TristageError) as e:
        DeviceBM25(bm, 0)
    assert e.value.code == -4


@pytest.mark.parametrize("method", ["rrf", "weighted"])
def test_device_fusion_is_bit_identical_to_the_reference_arithmetic(sim, method):
    """fuse_kernel (emulated) == Stage1Retriever._reciprocal_rank_fusion / _weighted_fusion (same
    arithmetic as the reference, :326-366): fp64 values, dict insertion order, stable sort."""
    import types

    from tristage_rag_b200.stage1_retriever import Stage1Retriever

    rng = np.random.default_rng(3)
    cfg = types.SimpleNamespace(rrf_k=60, dense_weight=0.7, bm25_weight=0.3)
    host = types.SimpleNamespace(config=cfg)
    fuse = Stage1Retriever._reciprocal_rank_fusion if method == "rrf" else Stage1Retriever._weighted_fusion
    B, k1, k2, top_k = 6, 50, 30, 40
    dense_ids = np.full((B, k1), -1, np.int64)
    dense_sc = np.full((B, k1), flat_ip.LOWEST_F32, np.float32)
    bm_ids = np.full((B, k2), -1, np.int64)
    bm_sc = np.zeros((B, k2), np.float64)
    want = []
    for b in range(B):
        nd = [50, 50, 7, 50, 1, 20][b]
        nb = [30, 30, 30, 0, 30, 3][b]
        universe = 60 if b != 1 else 5000                  # b == 1: (almost) no overlap between the lists
        d_ids = rng.choice(universe, size=nd, replace=False)
        d_sc = np.sort(rng.random(nd).astype(np.float32))[::-1] + np.float32(0.01)
        if b == 5:
            d_sc[:] = d_sc[0]                              # exact ties everywhere
        m_ids = rng.choice(universe, size=nb, replace=False)
        m_sc = np.sort(rng.random(nb) * 9)[::-1] + 0.5
        dense_ids[b, :nd], dense_sc[b, :nd] = d_ids, d_sc
        bm_ids[b, :nb], bm_sc[b, :nb] = m_ids, m_sc
        dense = [(int(i), float(s)) for i, s in zip(d_ids, d_sc)]
        bm25 = [(int(i), float(s)) for i, s in zip(m_ids, m_sc)]
        want.append((fuse(host, dense, bm25) if bm25 else dense)[:top_k] if bm25 else None)
    ids, scores, n = _lib.hybrid_fuse(method, 60, 0.7, 0.3, dense_ids, dense_sc, bm_ids, bm_sc, top_k, 0)
    for b in range(B):
        got = [(int(ids[b, r]), float(scores[b, r])) for r in range(int(n[b]))]
        if want[b] is None:                                # empty BM25 list: the kernel still returns the dense part
            ref = fuse(host, [(int(i), float(s)) for i, s in zip(dense_ids[b], dense_sc[b]) if i >= 0], [])[:top_k]
            assert got == ref
        else:
            assert got == want[b], b
        assert (ids[b, int(n[b]):] == -1).all()


def test_stage1_search_batch_hybrid_on_device_equals_host_path(sim, tmp_path, monkeypatch):
    """Stage1Retriever.search_batch with hybrid_on_device=True (device BM25 + device fusion, emulated)
    returns exactly what the host BM25 + fusion path returns; the dense index is the oracle's here.
    After a second add_documents the reference's stale-refit quirk makes idf negative for common
    words: the device index refuses such weights and the call takes the host path."""
    from oracle import fakes
    from tristage_rag_b200 import stage1_retriever as s1

    monkeypatch.setattr(flat_ip.IndexFlatIP, "search", _search_ignoring_path, raising=True)
    docs, _ = _zipf_corpus(400, vocab_size=120, seed=5)
    queries = ["w0 w1", "w3 w3 w40", "", "w119 w2 w7", "unknown"]
    out = {}
    for refit in (False, True):
        for on_device in (False, True):
            for fusion in ("rrf", "weighted"):
                cfg = s1.Stage1Config(device="cpu", cache_dir=str(tmp_path / "m"), index_dir=str(tmp_path / "i"),
                                      top_k_candidates=50, enable_bm25=True, bm25_top_k=30, fusion_method=fusion,
                                      hybrid_on_device=on_device)
                r = s1.Stage1Retriever(cfg, model=fakes.FakeSentenceEncoder(64))
                r._create_faiss_index = lambda emb, r=r: (setattr(r, "faiss_index", flat_ip.IndexFlatIP(emb.shape[1])),
                                                           r.faiss_index.add(emb))
                if refit:
                    r.add_documents(docs[:250])
                    r.add_documents(docs[250:])
                else:
                    r.add_documents(docs)
                # weighted fusion divides by the best score of each list: the reference raises ZeroDivisionError
                # when a query has no lexical hit (or, after a stale refit, only negative ones)
                qs = queries if fusion == "rrf" else (["w3 w3 w40", "w2 w7"] if refit else ["w0 w1", "w3 w3 w40", "w2 w7"])
                out[(refit, on_device, fusion)] = r.search_batch(qs, 20)
                single = r.search(qs[0], 20)                   # the single-query call takes the same route (the numpy
                assert [x["doc_id"] for x in single] == [x["doc_id"] for x in out[(refit, on_device, fusion)][0]]   # stand-in
                assert [x["score"] for x in single] == pytest.approx(                       # index rounds B=1 differently)
                    [x["score"] for x in out[(refit, on_device, fusion)][0]], rel=1e-6)
                if on_device:
                    assert (getattr(r, "_device_bm25", None) is not None) == (not refit)
    for refit in (False, True):
        for fusion in ("rrf", "weighted"):
            assert out[(refit, True, fusion)] == out[(refit, False, fusion)]
            assert all(len(x) == 20 for x in out[(refit, True, fusion)])


_orig_search = flat_ip.IndexFlatIP.search


def _search_ignoring_path(self, q, k, path="auto", **kw):
    return _orig_search(self, q, k)


# ------------------------------------------- Stage 2, tensor path (tcgen05 / TMA model) ---
def _norm_round(x, dtype):
    return flat_ip.round_to(maxsim.l2_normalize_tokens(x), dtype)


def _make_store(lens, dim, dtype, seed=0):
    rng = np.random.default_rng(seed)
    tok = rng.standard_normal((int(np.sum(lens)), dim)).astype(np.float32)
    st = _lib.TokStore(dim, dtype, 0)
    cut = len(lens) // 3
    off = np.concatenate([[0], np.cumsum(lens)])
    for a, b in ((0, cut), (cut, len(lens))):
        if b > a:
            st.add(tok[off[a]:off[b]], lens[a:b], normalize=True)
    return st, [tok[off[i]:off[i + 1]] for i in range(len(lens))]


def _oracle_scores(q, docs, cand, mode, dtype, n_cand=None, q_len=None):
    B, Cn = cand.shape
    out = np.zeros((B, Cn), np.float32)
    for b in range(B):
        lq = q.shape[1] if q_len is None else int(q_len[b])
        qb = _norm_round(q[b, :lq], dtype)
        for j in range(Cn if n_cand is None else int(n_cand[b])):
            c = int(cand[b, j])
            if 0 <= c < len(docs):
                out[b, j] = maxsim.score(qb, _norm_round(docs[c], dtype), mode, normalize=False)
    return out


@pytest.mark.parametrize("dim,Lq,dtype", [(128, 32, "bf16"), (128, 32, "fp16"), (64, 7, "bf16"), (128, 1, "bf16"),
                                          (96, 33, "bf16"), (128, 128, "bf16"), (256, 32, "bf16"), (128, 70, "bf16")])
@pytest.mark.parametrize("mode", [0, 1])
def test_tensor_maxsim_matches_the_oracle_both_epilogues(sim, monkeypatch, dim, Lq, dtype, mode):
    """maxsim_umma_kernel on the functional model (ragged TMA gather, tile packing, per-quarter drain,
    finalize): the cases of tests/test_gpu_stage2.py::test_batch_vs_oracle_ragged, for the validated
    epilogue and for the opt-in TS_S2_V2 one, which must agree bit for bit."""
    rng = np.random.default_rng(dim + Lq)
    ndocs, B, Cn = 150, 3, 70
    lens = rng.integers(1, 181, size=ndocs)
    lens[:6] = [1, 8, 9, 255, 256, 180]
    st_flow, docs = _make_store(lens, dim, dtype, seed=dim)        # default: tile-layout shard, maxsim_flow_kernel
    q = rng.standard_normal((B, Lq, dim)).astype(np.float32)
    cand = rng.integers(0, ndocs, size=(B, Cn)).astype(np.int64)
    cand[0, :6] = np.arange(6)
    ref = _oracle_scores(q, docs, cand, mode, dtype)
    got_flow = st_flow.maxsim_host(q, cand, mode=mode)
    np.testing.assert_allclose(got_flow, ref, rtol=1e-3, atol=2e-4)
    np.testing.assert_allclose(st_flow.maxsim_host(q, cand, mode=mode | _lib.TS_S2_FORCE_SIMT), ref, rtol=1e-3, atol=2e-4)
    monkeypatch.setenv("TS_S2_FLOW", "0")                            # row-major shard: the first kernel and its variants
    st, _ = _make_store(lens, dim, dtype, seed=dim)
    got = st.maxsim_host(q, cand, mode=mode)
    np.testing.assert_allclose(got, ref, rtol=1e-3, atol=2e-4)
    np.testing.assert_allclose(got_flow, got, rtol=2e-6, atol=1e-7)  # same products and maxima; only the final sum's order differs
    monkeypatch.setenv("TS_S2_V2", "1")
    got2 = st.maxsim_host(q, cand, mode=mode)
    assert np.array_equal(got2, got)
    monkeypatch.setenv("TS_S2_STAGES", "4")                          # 4-stage ring, single maxima buffer
    assert np.array_equal(st.maxsim_host(q, cand, mode=mode), got)
    monkeypatch.delenv("TS_S2_V2")
    assert np.array_equal(st.maxsim_host(q, cand, mode=mode), got)
    # two epilogue warpgroups (TS_S2_EPI2: 320 threads, one group per accumulator), with either epilogue
    monkeypatch.setenv("TS_S2_EPI2", "1")
    assert np.array_equal(st.maxsim_host(q, cand, mode=mode), got)
    monkeypatch.setenv("TS_S2_V2", "1")
    assert np.array_equal(st.maxsim_host(q, cand, mode=mode), got)
    for sms in (("1", "3") if (dim, Lq, dtype) == (128, 32, "bf16") else ()):   # few SMs: many tiles per CTA, odd and even counts
        monkeypatch.setenv("HOSTSIM_SM_COUNT", sms)
        st2, _ = _make_store(lens, dim, dtype, seed=dim)
        assert np.array_equal(st2.maxsim_host(q, cand, mode=mode), got)


@pytest.mark.parametrize("v2,epi2", [(False, False), (True, False), (False, True), (True, True)])
def test_tensor_maxsim_n_cand_q_len_unowned_and_length_mixes(sim, monkeypatch, v2, epi2):
    monkeypatch.setenv("TS_S2_FLOW", "0")
    if v2:
        monkeypatch.setenv("TS_S2_V2", "1")
    if epi2:
        monkeypatch.setenv("TS_S2_EPI2", "1")
    rng = np.random.default_rng(2)
    dim, ndocs, B, Cn, Lq = 128, 200, 6, 64, 32
    lens = rng.integers(16, 181, size=ndocs)
    st, docs = _make_store(lens, dim, "bf16", seed=4)
    q = rng.standard_normal((B, Lq, dim)).astype(np.float32)
    cand = rng.integers(0, ndocs, size=(B, Cn)).astype(np.int64)
    cand[1, 5] = -1
    cand[2, 7] = ndocs + 10
    n_cand = np.array([64, 10, 0, 33, 1, 64], np.int32)
    q_len = np.array([32, 5, 32, 17, 1, 31], np.int32)
    for mode in (0, 1):
        got = st.maxsim_host(q, cand, q_len=q_len, n_cand=n_cand, mode=mode)
        ref = _oracle_scores(q, docs, cand, mode, "bf16", n_cand=n_cand, q_len=q_len)
        np.testing.assert_allclose(got, ref, rtol=1e-3, atol=2e-4)
        assert got[1, 5] == 0.0 and (got[2] == 0.0).all() and (got[1, 10:] == 0.0).all()
    st.set_id_base(1000)
    np.testing.assert_allclose(st.maxsim_host(q, cand + 1000, mode=0), _oracle_scores(q, docs, cand, 0, "bf16"),
                               rtol=1e-3, atol=2e-4)
    assert (st.maxsim_host(q, cand, mode=0) == 0.0).all()
    # doc-length mixes that stress the packing: only tiny docs (32 per tile), only 256-token docs, quarter straddlers
    for name, ln in (("tiny", np.full(90, 3)), ("max", np.full(12, 256)), ("eights", np.full(70, 8)),
                     ("straddle", np.array([60, 10, 120, 7, 59, 130, 3, 3, 3, 250, 5, 1, 63, 1, 64, 1, 127, 129] * 3))):
        st2, docs2 = _make_store(ln, 64, "bf16", seed=len(ln))
        q2 = rng.standard_normal((2, 20, 64)).astype(np.float32)
        c2 = np.stack([rng.permutation(len(ln))[: min(len(ln), 40)] for _ in range(2)]).astype(np.int64)
        np.testing.assert_allclose(st2.maxsim_host(q2, c2), _oracle_scores(q2, docs2, c2, 0, "bf16"), rtol=1e-3, atol=2e-4,
                                   err_msg=name)


@pytest.mark.parametrize("stages,a_bufs,sms,async_seed", [(0, 0, 0, 0), (2, 1, 2, 0), (3, 2, 3, 5), (5, 1, 1, 9), (4, 2, 5, 0)])
def test_flow_maxsim_splits_docs_across_tiles_and_keeps_the_query_resident(sim, monkeypatch, stages, a_bufs, sms, async_seed):
    """maxsim_flow_kernel (s2_flow.cu) on tile-layout shards: docs split across tile boundaries (their maxima
    meet in the doc's slot), slot generations (more than 96 docs per CTA), query changes inside a CTA with one
    and two query-tile buffers, n_cand / q_len / un-owned ids, ring depths, both tile widths, few SMs (long
    pipelines) and adversarial timing -- against the oracle and against the first kernel on a row-major shard."""
    if stages in (3, 4):
        monkeypatch.setenv("TS_S2_TILE", "128")
    if stages:
        monkeypatch.setenv("TS_S2_STAGES", str(stages))
    if a_bufs:
        monkeypatch.setenv("TS_S2_ABUFS", str(a_bufs))
    if sms:
        monkeypatch.setenv("HOSTSIM_SM_COUNT", str(sms))
    monkeypatch.setenv("CUDASIM_ASYNC", str(async_seed))
    rng = np.random.default_rng(11 + stages)
    dim, ndocs = 128, 220
    lens = rng.integers(1, 257, size=ndocs)
    lens[:12] = [256, 255, 249, 248, 1, 8, 9, 7, 200, 56, 256, 64]
    st, docs = _make_store(lens, dim, "bf16", seed=3)
    monkeypatch.setenv("TS_S2_FLOW", "0")
    st_first, _ = _make_store(lens, dim, "bf16", seed=3)             # row-major shard: first kernel
    monkeypatch.delenv("TS_S2_FLOW")
    assert st.layout == 1 and st_first.layout == 0
    for B, Cn, Lq in ((5, 37, 32), (3, 90, 70), (2, 130, 128), (9, 9, 20)):
        q = rng.standard_normal((B, Lq, dim)).astype(np.float32)
        cand = rng.integers(0, ndocs, size=(B, Cn)).astype(np.int64)
        cand[0, : min(Cn, 12)] = np.arange(min(Cn, 12))
        cand[1, 3] = -1
        cand[1, 4] = ndocs + 3
        n_cand = rng.integers(0, Cn + 1, size=B).astype(np.int32)
        n_cand[0] = Cn
        q_len = rng.integers(1, Lq + 1, size=B).astype(np.int32)
        q_len[0] = Lq
        for mode in (0, 1):
            got = st.maxsim_host(q, cand, q_len=q_len, n_cand=n_cand, mode=mode)
            ref = _oracle_scores(q, docs, cand, mode, "bf16", n_cand=n_cand, q_len=q_len)
            np.testing.assert_allclose(got, ref, rtol=1e-3, atol=2e-4, err_msg=f"{B}x{Cn} Lq={Lq} mode={mode}")
            assert got[1, 3] == 0.0 and got[1, 4] == 0.0
            assert np.array_equal(st.maxsim_host(q, cand, q_len=q_len, n_cand=n_cand, mode=mode), got)     # deterministic
            np.testing.assert_allclose(st_first.maxsim_host(q, cand, q_len=q_len, n_cand=n_cand, mode=mode), got, rtol=2e-6, atol=1e-7)
    # length mixes: only tiny docs (the 16-segment cap closes tiles early), only 256-token docs (every doc but
    # the first is split), exact multiples of the tile, quarter straddlers
    for name, ln in (("tiny", np.full(90, 3)), ("max", np.full(12, 256)), ("eights", np.full(70, 8)), ("128s", np.full(20, 128)),
                     ("248", np.full(15, 248)),
                     ("straddle", np.array([60, 10, 120, 7, 59, 130, 3, 3, 3, 250, 5, 1, 63, 1, 64, 1, 127, 129] * 3))):
        st2, docs2 = _make_store(ln, 64, "bf16", seed=len(ln))
        q2 = rng.standard_normal((2, 20, 64)).astype(np.float32)
        c2 = np.stack([rng.permutation(len(ln))[: min(len(ln), 40)] for _ in range(2)]).astype(np.int64)
        for mode in (0, 1):
            np.testing.assert_allclose(st2.maxsim_host(q2, c2, mode=mode), _oracle_scores(q2, docs2, c2, mode, "bf16"), rtol=1e-3,
                                       atol=2e-4, err_msg=name)


# ------------------------------------------- multi-GPU exchange over peer memory ---
@pytest.mark.parametrize("N", [1500, 90])
def test_peer_memory_exchange_equals_gather_and_merge(sim, N):
    """(N = 90: 30 rows per shard < k, so every list ends in a -1 tail and the merged result is padded too.)
    ts_exchange_push / ts_exchange_wait_merge with G ranks living in one process (every rank's receive
    buffer is plain host memory here, so "peer" pointers are ordinary pointers): after all ranks pushed step
    s, every rank's wait+merge returns what ts_topk_merge_packed returns over the gathered lists; steps
    alternate the two parities and the sequence numbers."""
    G, B, k = 3, 5, 40
    ids_off, nbytes = _lib.packed_layout(B, k)
    slot = (nbytes + 15) // 16 * 16
    flags_off = 2 * G * slot
    bufs = [np.zeros(flags_off + 2 * G * 4 + 16, np.uint8) for _ in range(G)]
    bases = np.array([b.ctypes.data for b in bufs], np.int64)
    d = 32
    X, _ = make(N, d, 1, seed=4)
    X[N // 2] = X[3]
    X[N - 7] = X[3]                                            # exact ties across shards: ids must ascend
    shards = []
    for r in range(G):
        lo, hi = r * N // G, (r + 1) * N // G
        sh = _lib.Index(d, "bf16", "ip", 0)
        sh.add(X[lo:hi])
        sh.set_id_base(lo)
        shards.append(sh)
    rng = np.random.default_rng(9)
    for step in range(4):
        Q = flat_ip.normalize_rows(rng.standard_normal((B, d)).astype(np.float32)).astype(np.float32)
        parity, seq = step & 1, step + 1
        blobs = []
        for r in range(G):
            D, I = shards[r].search_host(Q, k, path="stream")
            blob = np.zeros(slot, np.uint8)
            blob[: B * k * 4] = D.view(np.uint8).ravel()
            blob[ids_off: ids_off + B * k * 8] = I.view(np.uint8).ravel()
            blobs.append(blob)
        for r in range(G):                                     # every rank publishes its list into every buffer
            _lib.check(sim.ts_exchange_push(0, p(blobs[r]), nbytes, p(bases), G, r, slot, flags_off, parity, seq, None))
        gathered = np.concatenate(blobs)
        ref_s, ref_i = np.empty((B, k), np.float32), np.empty((B, k), np.int64)
        _lib.check(sim.ts_topk_merge_packed(0, p(gathered), slot, ids_off, G, B, k, p(ref_s), p(ref_i), None))
        for r in range(G):
            out_s, out_i = np.empty((B, k), np.float32), np.empty((B, k), np.int64)
            _lib.check(sim.ts_exchange_wait_merge(0, p(bufs[r]), G, B, k, slot, ids_off, flags_off, parity, seq, p(out_s), p(out_i), None))
            assert (out_i == ref_i).all() and (out_s == ref_s).all(), (step, r)
        full = _lib.Index(d, "bf16", "ip", 0)
        full.add(X)
        Df, If = full.search_host(Q, k, path="stream")
        assert (ref_i == If).all() and (ref_s == Df).all()
    # flags now hold the sequence numbers of the last step of each parity
    for b in bufs:
        assert np.frombuffer(b, np.uint32, 2 * G, flags_off).tolist() == [3] * G + [4] * G
    # argument checks
    assert sim.ts_exchange_push(0, p(blobs[0]), nbytes, p(bases), G, G, slot, flags_off, 0, 1, None) == -1       # rank out of range
    assert sim.ts_exchange_wait_merge(0, p(bufs[0]), G, B, k, slot + 8, ids_off, flags_off, 0, 1, p(out_s), p(out_i), None) == -1


@pytest.mark.parametrize("path", ["stream", "umma"])
def test_fused_exchange_select_pushes_rows_and_merge_waits_per_query(sim, path):
    """The fused exchange (ts_exchange_*): the select kernel of every "rank" stores its rows into every rank's
    receive buffer and publishes per-query flags, the merge kernel waits for its query's rows.  G ranks live in
    one process here (buffers are host memory, "peer" pointers ordinary pointers; all pushes of a step run before
    the merges, as separate launches cannot wait on each other in the emulator).  Result == the oracle over all
    rows, on every rank, over several steps (both parities) and changing batch sizes."""
    G, N, d, k = 3, 1800, 64, 40
    X, _ = make(N, d, 1, seed=14)
    X[900] = X[5]
    X[1700] = X[5]                                             # exact ties across shards: ids must ascend
    B_max, k_max = 16, 64
    nbytes = _lib.Exchange.buffer_bytes(G, B_max, k_max)
    bufs = [np.zeros(nbytes + 16, np.uint8) for _ in range(G)]
    bases = [b.ctypes.data + (-b.ctypes.data % 16) for b in bufs]
    shards, xs = [], []
    for r in range(G):
        lo, hi = r * N // G, (r + 1) * N // G
        sh = _lib.Index(d, "bf16", "ip", 0)
        sh.add(X[lo:hi])
        sh.set_id_base(lo)
        shards.append(sh)
        xs.append(_lib.Exchange(0, r, G, bases, B_max, k_max))
    rng = np.random.default_rng(19)
    for step, B in enumerate((5, 16, 1, 7, 7)):
        Q = flat_ip.normalize_rows(rng.standard_normal((B, d)).astype(np.float32)).astype(np.float32)
        Q[0] = X[5]
        for r in range(G):
            _lib.check(sim.ts_index_search_push(shards[r]._h, xs[r]._h, p(Q), _lib.TS_F32, B, k, 0, _lib.PATHS[path], None))
        rD, rI, sc = oracle_search(X, Q, k, "bf16")
        for r in range(G):
            out_s, out_i = np.empty((B, k), np.float32), np.empty((B, k), np.int64)
            _lib.check(sim.ts_exchange_merge(xs[r]._h, B, k, p(out_s), p(out_i), None))
            assert not flat_ip.check_topk(out_s, out_i, sc, rD, rI, rel=REL), (step, r)
            assert out_i[0, :3].tolist() == [5, 900, 1700]
            if r == 0:
                first = (out_s.copy(), out_i.copy())
            assert (out_i == first[1]).all() and (out_s == first[0]).all()      # identical on every rank
    assert sim.ts_index_search_push(shards[0]._h, xs[0]._h, p(Q), _lib.TS_F32, B_max + 1, k, 0, 0, None) == -1   # over capacity
    assert sim.ts_exchange_create(C.byref(C.c_void_p()), 0, G, G, p(np.array(bases, np.int64)), B_max, k_max) == -1   # rank out of range


# ------------------------------------------- long pipelines: few SMs, many tiles per CTA ---
@pytest.mark.parametrize("sms,async_seed", [(2, 0), (6, 0), (3, 7)])
def test_few_sms_many_tiles_per_cta_wrap_every_ring(sim, monkeypatch, sms, async_seed):
    """On a 148-SM "GPU" a small corpus gives every CTA one or two tiles, so the shared-memory rings,
    the TMEM double buffer and the mbarrier phase bits hardly ever wrap.  Shrinking the emulated GPU to
    a few SMs makes each CTA walk dozens of tiles: every ring wraps many times, in all scan variants
    and in the Stage-2 kernel, with the results still equal to the oracle / to each other.
    async_seed != 0 adds adversarial timing (CUDASIM_ASYNC): TMA loads and tensor-core operations complete a
    random number of scheduler rounds after issue, the thread order is reshuffled every round and random
    threads stall -- a kernel that reads data before its barrier, or reuses a stage too early, fails
    (checked by removing single waits from the kernels: both the emulator's diagnostics and these
    comparisons catch it)."""
    monkeypatch.setenv("HOSTSIM_SM_COUNT", str(sms))
    monkeypatch.setenv("CUDASIM_ASYNC", str(async_seed))
    N, d, k = 6500, 200, 100
    for B in (7, 200):
        X, Q = make(N, d, B, seed=B, planted=10)
        idx = _lib.Index(d, "bf16", "ip", 0)
        idx.add(X)
        D, I = idx.search_host(Q, k, path="umma")
        rD, rI, sc = oracle_search(X, Q, k, "bf16")
        assert not flat_ip.check_topk(D, I, sc, rD, rI, rel=REL)
        for var in ("TS_FUSE", "TS_PAIR", "TS_DBG_NOSHARE"):
            monkeypatch.setenv(var, "1")
            D2, I2 = idx.search_host(Q, k, path="umma")
            monkeypatch.delenv(var)
            assert (I2 == I).all() and (D2 == D).all(), (var, B)
    rng = np.random.default_rng(sms)
    lens = rng.integers(1, 200, size=300)
    st, docs = _make_store(lens, 128, "bf16", seed=sms)
    for Lq in (32, 80):
        q = rng.standard_normal((4, Lq, 128)).astype(np.float32)
        cand = rng.integers(0, 300, size=(4, 160)).astype(np.int64)
        got = st.maxsim_host(q, cand)
        np.testing.assert_allclose(got, _oracle_scores(q, docs, cand, 0, "bf16"), rtol=1e-3, atol=2e-4)
        monkeypatch.setenv("TS_S2_FLOW", "0")
        st_first, _ = _make_store(lens, 128, "bf16", seed=sms)
        got_first = st_first.maxsim_host(q, cand)
        np.testing.assert_allclose(got_first, got, rtol=2e-6, atol=1e-7)
        monkeypatch.setenv("TS_S2_V2", "1")
        assert np.array_equal(st_first.maxsim_host(q, cand), got_first)
        monkeypatch.delenv("TS_S2_V2")
        monkeypatch.delenv("TS_S2_FLOW")


def test_load_index_imports_a_reference_written_faiss_file(sim, tmp_path):
    """Migration: index_dir holds what the REFERENCE saved (pickle side-car + a flat FAISS index file, layout as in
    tristage_rag_b200/faiss_io.py); load_index imports the vectors into a ts_index and search works on them."""
    import pickle
    import struct

    from oracle import fakes
    from tristage_rag_b200 import stage1_retriever as s1

    docs = [f"document {i} about topic {i % 5}" for i in range(40)]
    enc = fakes.FakeSentenceEncoder(48)
    emb = flat_ip.normalize_rows(np.asarray(enc.encode(docs), np.float32)).astype(np.float32)
    idir = tmp_path / "i"
    idir.mkdir()
    with open(idir / "stage1_faiss.index", "wb") as f:
        f.write(b"IxFI" + struct.pack("<iqqqBi", 48, 40, 1 << 20, 1 << 20, 1, 0) + struct.pack("<Q", 40 * 48) + emb.tobytes())
    with open(idir / "stage1_index.pkl", "wb") as f:
        pickle.dump({"documents": docs, "doc_metadata": [{} for _ in docs], "config": {}, "bm25_index": None}, f)
    cfg = s1.Stage1Config(device="cpu", cache_dir=str(tmp_path / "m"), index_dir=str(idir), enable_bm25=False,
                          top_k_candidates=5, storage_dtype="fp32")
    r = s1.Stage1Retriever(cfg, model=enc)
    r.load_index()
    assert r.faiss_index.ntotal == 40 and len(r.documents) == 40
    res = r.search(docs[7])
    assert res[0]["doc_id"] == 7 and res[0]["score"] == pytest.approx(1.0, abs=1e-5)
    r.save_index()                                      # ... and is written back as a tristage shard file
    assert _lib.file_probe(str(idir / "stage1_faiss.index"))["n"] == 40
    r2 = s1.Stage1Retriever(cfg, model=enc)
    r2.load_index()
    assert [x["doc_id"] for x in r2.search(docs[7])] == [x["doc_id"] for x in res]


def test_peer_memory_sum_of_stage2_scores(sim):
    """ts_exchange_push + ts_exchange_wait_sum with three in-process ranks: the element-wise sum of the
    ownership-filtered per-rank MaxSim outputs, i.e. what all-reduce(SUM) delivers, over several steps."""
    G, B, Cn = 3, 4, 37                                        # 148 floats: not a multiple of 4 -> padded slot
    n = B * Cn
    slot = (n * 4 + 15) // 16 * 16
    flags_off = 2 * G * slot
    bufs = [np.zeros(flags_off + 2 * G * 4 + 16, np.uint8) for _ in range(G)]
    bases = np.array([b.ctypes.data for b in bufs], np.int64)
    rng = np.random.default_rng(1)
    lens = rng.integers(2, 60, size=90)
    tok = rng.standard_normal((int(lens.sum()), 32)).astype(np.float32)
    off = np.concatenate([[0], np.cumsum(lens)])
    stores = []
    for r in range(G):
        lo, hi = r * 30, (r + 1) * 30
        st = _lib.TokStore(32, "bf16", 0)
        st.add(tok[off[lo]:off[hi]], lens[lo:hi], normalize=True)
        st.set_id_base(lo)
        stores.append(st)
    whole = _lib.TokStore(32, "bf16", 0)
    whole.add(tok, lens, normalize=True)
    for step in range(3):
        q = rng.standard_normal((B, 8, 32)).astype(np.float32)
        cand = rng.integers(0, 90, size=(B, Cn)).astype(np.int64)
        parity, seq = step & 1, step + 1
        for r in range(G):
            part = np.zeros(slot // 4, np.float32)
            part[:n] = stores[r].maxsim_host(q, cand).ravel()
            _lib.check(sim.ts_exchange_push(0, p(part), slot, p(bases), G, r, slot, flags_off, parity, seq, None))
        ref = whole.maxsim_host(q, cand)
        for r in range(G):
            out = np.empty((B, Cn), np.float32)
            _lib.check(sim.ts_exchange_wait_sum(0, p(bufs[r]), G, n, slot, flags_off, parity, seq, p(out), None))
            assert np.array_equal(out, ref), (step, r)


def test_stage2_scatter_fused_into_the_scoring_kernel(sim):
    """ts_maxsim_scatter + ts_exchange_wait_take with three in-process ranks: every rank's flow kernel stores the scores
    of the candidates it owns into ALL ranks' matrices and its last CTA publishes the step; the consumer takes the
    matrix and leaves zeros.  Result == the single-store kernel on the whole corpus (un-owned / invalid ids and
    positions beyond n_cand read 0.0), over several steps so both parities are reused."""
    G, B, Cn, dim, Lq = 3, 5, 41, 32, 8
    n = B * Cn
    slot = (n * 4 + 15) // 16 * 16
    flags_off = 2 * slot
    bufs = [np.zeros(flags_off + 2 * G * 4 + 16, np.uint8) for _ in range(G)]
    bases = np.array([b.ctypes.data for b in bufs], np.int64)
    rng = np.random.default_rng(5)
    lens = rng.integers(2, 70, size=90)
    tok = rng.standard_normal((int(lens.sum()), dim)).astype(np.float32)
    off = np.concatenate([[0], np.cumsum(lens)])
    stores = []
    for r in range(G):
        lo, hi = r * 30, (r + 1) * 30
        st = _lib.TokStore(dim, "bf16", 0)
        assert st.layout == 1
        st.add(tok[off[lo]:off[hi]], lens[lo:hi], normalize=True)
        st.set_id_base(lo)
        stores.append(st)
    whole = _lib.TokStore(dim, "bf16", 0)
    whole.add(tok, lens, normalize=True)
    for step in range(4):
        q = rng.standard_normal((B, Lq, dim)).astype(np.float32)
        cand = rng.integers(-2, 95, size=(B, Cn)).astype(np.int64)         # a few ids nobody owns
        n_cand = rng.integers(Cn // 2, Cn + 1, size=B).astype(np.int32)
        parity, seq = step & 1, step + 1
        mat_off, f_off = parity * slot, flags_off + parity * G * 4
        for r in range(G):
            _lib.check(sim.ts_maxsim_scatter(stores[r]._h, p(q), _lib.TS_F32, None, B, Lq, p(cand), p(n_cand), Cn, 0,
                                             _lib.TS_FLAG_NORMALIZE_Q, p(bases), G, r, mat_off, f_off, seq, None))
        ref = whole.maxsim_host(q, cand, n_cand=n_cand)
        for r in range(G):
            out = np.empty((B, Cn), np.float32)
            mat = bufs[r][mat_off:].ctypes.data
            flg = bufs[r][f_off:].ctypes.data
            _lib.check(sim.ts_exchange_wait_take(0, C.c_void_p(mat), C.c_void_p(flg), G, seq, n, p(out), None))
            assert np.array_equal(out, ref), (step, r, np.abs(out - ref).max())
            assert not bufs[r][mat_off:mat_off + n * 4].any()              # left clean for step + 2
