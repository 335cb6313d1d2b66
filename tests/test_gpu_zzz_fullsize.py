"""(Named zzz so that it runs LAST under `pytest -x`: it is the longest file and the only one whose logic cannot be
rehearsed on the CPU emulator, because it builds its corpora with torch CUDA tensors.)
GPU: BASELINE.json's full sizes through size-independent properties
(planted known answers, sortedness, idempotence, shard-merge == full search,
agreement of the two scan kernels) plus oracle checks on a few queries."""
import os

import numpy as np
import pytest
import torch

from test_gpu_stage1 import REL, make, oracle_search

from oracle import flat_ip, maxsim
from tristage_rag_b200 import _lib

pytestmark = pytest.mark.gpu


def build_planted(cuda_device, N, d, B_plant, n_plant, seed, chunk=500_000):
    """randn corpus normalised like the reference, with n_plant rows per planted query
    overwritten by normalize(q + sigma*noise).  Returns index, queries, planted id sets."""
    dev = torch.device("cuda", cuda_device)
    g = torch.Generator(device=dev).manual_seed(seed)
    q = torch.randn((B_plant, d), generator=g, device=dev)
    q /= q.norm(dim=1, keepdim=True) + 1e-8
    rng = np.random.default_rng(seed)
    pos = rng.choice(N, size=B_plant * n_plant, replace=False).reshape(B_plant, n_plant)
    pos_t = torch.from_numpy(pos).to(dev)
    idx = _lib.Index(d, "bf16", "ip", cuda_device, reserve_rows=N)
    for s in range(0, N, chunk):
        n = min(chunk, N - s)
        x = torch.randn((n, d), generator=g, device=dev)
        for b in range(B_plant):
            sel = pos_t[b][(pos_t[b] >= s) & (pos_t[b] < s + n)] - s
            if len(sel):
                noise = torch.randn((len(sel), d), generator=g, device=dev) / d ** 0.5
                x[sel] = q[b][None, :] + 0.7 * noise
        x /= x.norm(dim=1, keepdim=True) + 1e-8
        idx.add(x.to(torch.bfloat16))
        del x
    return idx, q, pos


def check_common(s, i, N):
    s, i = s.cpu().numpy(), i.cpu().numpy()
    assert (np.diff(s, axis=1) <= 0).all(), "scores must be descending"
    assert (i >= 0).all() and (i < N).all()
    assert all(len(set(r.tolist())) == len(r) for r in i), "duplicate ids"
    return s, i


@pytest.mark.parametrize("B", [1, 32, 1024])
def test_config2_1Mx768_properties(cuda_device, B):
    N, d, k, BP = 1_000_000, 768, 100, 4
    idx, qp, pos = build_planted(cuda_device, N, d, BP, 100, seed=2)
    dev = qp.device
    g = torch.Generator(device=dev).manual_seed(5)
    q = torch.randn((B, d), generator=g, device=dev)
    q /= q.norm(dim=1, keepdim=True) + 1e-8
    q[: min(B, BP)] = qp[: min(B, BP)]
    s, i = idx.search(q, k)
    s2, i2 = idx.search(q, k)                                   # idempotent, bit-identical
    torch.cuda.synchronize()
    assert torch.equal(i, i2) and torch.equal(s, s2)
    sn, inn = check_common(s, i, N)
    for b in range(min(B, BP)):                                  # known answer: the planted rows
        assert set(inn[b].tolist()) == set(pos[b].tolist()), f"query {b}: planted rows not recovered"
        assert sn[b, -1] > 0.5
    # the CUDA-core scan agrees (ids up to near-ties) with the tensor-path scan
    if B <= 4:
        s3, i3 = idx.search(q, k, path="stream")
        a, b_ = i.cpu().numpy(), i3.cpu().numpy()
        sa = s.cpu().numpy()
        for r in range(B):
            diff = set(a[r].tolist()) ^ set(b_[r].tolist())
            assert len(diff) <= 4, "paths disagree beyond near-ties"
        np.testing.assert_allclose(s3.cpu().numpy()[:, 0], sa[:, 0], rtol=1e-3)
    # oracle on the unplanted tail of the batch (fp32 scores of the stored bf16 rows)
    rows = torch.from_numpy(idx.get_rows(0, 200_000))            # first 200k rows, as stored
    sub = _lib.Index(d, "bf16", "ip", cuda_device)
    sub.add(rows.numpy())
    nq = min(B, 6)
    qs = q[-nq:].contiguous()
    ss, si = sub.search(qs, k)
    Qr = flat_ip.round_to(qs.cpu().numpy(), "bf16")
    Xr = rows.numpy()
    rD, rI = flat_ip.topk_desc(Qr @ Xr.T, k)
    bad = flat_ip.check_topk(ss.cpu().numpy(), si.cpu().numpy(),
                             lambda b, ids: Xr[ids].astype(np.float64) @ Qr[b].astype(np.float64), rD, rI)
    assert not bad, bad[:3]


def _oracle_over_the_whole_shard(idx, N, q, k, chunk):
    """exact fp32 top-k of the STORED rows, fetched back shard by shard (oracle/flat_ip.search_streamed)"""
    Qr = flat_ip.round_to(q.cpu().numpy(), "bf16")
    rD, rI = flat_ip.search_streamed(((s, idx.get_rows(s, min(chunk, N - s))) for s in range(0, N, chunk)), Qr, k)

    def scores_of(b, ids):
        return np.array([idx.get_rows(int(i), 1)[0].astype(np.float64) @ Qr[b].astype(np.float64) for i in ids])
    return rD, rI, scores_of


def test_config2_1Mx768_whole_corpus_against_the_oracle(cuda_device):
    """BASELINE configs[1] at its full size: 8 queries of a batch of 32 checked against the oracle's exact
    search over ALL 1 M stored rows (not a sub-index), B = 32 on the tensor path and B = 1 on both paths."""
    N, d, k, B = 1_000_000, 768, 100, 32
    idx, qp, pos = build_planted(cuda_device, N, d, 2, 100, seed=12)
    dev = qp.device
    g = torch.Generator(device=dev).manual_seed(15)
    q = torch.randn((B, d), generator=g, device=dev)
    q /= q.norm(dim=1, keepdim=True) + 1e-8
    q[:2] = qp
    s, i = idx.search(q, k)
    torch.cuda.synchronize()
    rD, rI, scores_of = _oracle_over_the_whole_shard(idx, N, q[:8], k, 250_000)
    bad = flat_ip.check_topk(s[:8].cpu().numpy(), i[:8].cpu().numpy(), scores_of, rD, rI)
    assert not bad, bad[:3]
    for path in ("umma", "stream"):
        s1, i1 = idx.search(q[2:3].contiguous(), k, path=path)
        bad = flat_ip.check_topk(s1.cpu().numpy(), i1.cpu().numpy(), lambda b, ids: scores_of(2, ids), rD[2:3], rI[2:3])
        assert not bad, (path, bad[:3])


def test_config3_10Mx1024_two_queries_against_the_oracle(cuda_device):
    """BASELINE configs[2] at its full size on one GPU: 2 queries of the batch of 32 (one planted, one not)
    against the oracle's exact search over all 10 M stored rows, fetched back in 1 M-row shards."""
    free, _ = torch.cuda.mem_get_info(cuda_device)
    if free < 60 * 2 ** 30:
        pytest.skip("needs ~45 GB of free HBM")
    N, d, k, B = 10_000_000, 1024, 100, 32
    idx, qp, pos = build_planted(cuda_device, N, d, 1, 100, seed=13, chunk=1_000_000)
    dev = qp.device
    g = torch.Generator(device=dev).manual_seed(19)
    q = torch.randn((B, d), generator=g, device=dev)
    q /= q.norm(dim=1, keepdim=True) + 1e-8
    q[0] = qp[0]
    s, i = idx.search(q, k)
    torch.cuda.synchronize()
    sel = q[[0, 31]].contiguous()
    rD, rI, scores_of = _oracle_over_the_whole_shard(idx, N, sel, k, 1_000_000)
    got_s, got_i = s[[0, 31]].cpu().numpy(), i[[0, 31]].cpu().numpy()
    bad = flat_ip.check_topk(got_s, got_i, scores_of, rD, rI)
    assert not bad, bad[:3]
    assert set(got_i[0].tolist()) == set(pos[0].tolist())


def test_config4_full_store_all_queries_against_the_c_oracle(cuda_device, monkeypatch):
    """BASELINE configs[3] at its full size: 64 queries x 1000 candidates drawn from a 1 M-doc token store
    (Ld ~ U[16,180], dim 128, Lq 32); EVERY score is checked against oracle/c/oracle.c (maxsim_batch over the
    stored bf16 values in fp32), both scoring modes, and the two tensor kernels must agree bit for bit."""
    from oracle import c_oracle
    free, _ = torch.cuda.mem_get_info(cuda_device)
    if free < 70 * 2 ** 30:
        pytest.skip("needs ~55 GB of free HBM")
    rng = np.random.default_rng(77)
    dim, ndocs, B, C, Lq = 128, 1_000_000, 64, 1000, 32
    lens = rng.integers(16, 181, size=ndocs).astype(np.int32)
    dev = torch.device("cuda", cuda_device)
    g = torch.Generator(device=dev).manual_seed(77)
    st = _lib.TokStore(dim, "bf16", cuda_device, reserve_docs=ndocs, reserve_tokens=int(lens.sum()))
    toks = []
    for s0 in range(0, ndocs, 100_000):
        ln = lens[s0:s0 + 100_000]
        t = torch.nn.functional.normalize(torch.randn((int(ln.sum()), dim), generator=g, device=dev), dim=-1).to(torch.bfloat16)
        st.add(t, ln, normalize=False)
        toks.append(t)
    tok = torch.cat(toks)
    del toks
    off = torch.from_numpy(np.concatenate([[0], np.cumsum(lens.astype(np.int64))])).to(dev)
    q = torch.nn.functional.normalize(torch.randn((B, Lq, dim), generator=g, device=dev), dim=-1).to(torch.bfloat16)
    cand = torch.stack([torch.randperm(ndocs, generator=g, device=dev)[:C] for _ in range(B)])
    lens_t = torch.from_numpy(lens.astype(np.int64)).to(dev)
    for mode in (0, 1):
        scores = st.maxsim(q, cand, normalize_q=False, mode=mode)
        torch.cuda.synchronize()
        monkeypatch.setenv("TS_S2_FLOW", "0")
        first = st.maxsim(q, cand, normalize_q=False, mode=mode)
        torch.cuda.synchronize()
        monkeypatch.delenv("TS_S2_FLOW")
        assert torch.equal(first, scores), "the two tensor kernels disagree"
        sh = scores.cpu().numpy()
        qh = q.float().cpu().numpy()
        worst = 0.0
        for b in range(B):
            ln = lens_t[cand[b]]
            o = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(ln, 0)])
            # rows of the query's 1000 candidates, in candidate order
            ridx = torch.repeat_interleave(off[cand[b]] - o[:-1], ln) + torch.arange(int(o[-1]), device=dev)
            rows = tok[ridx].float().cpu().numpy()
            ref = c_oracle.maxsim_batch(qh[b], rows, o.cpu().numpy(), mode=mode)
            np.testing.assert_allclose(sh[b], ref, rtol=1e-3, atol=2e-4, err_msg=f"query {b} mode {mode}")
            worst = max(worst, float(np.abs(sh[b] - ref).max()))
        assert worst < 1e-3


def test_config3_10Mx1024_properties(cuda_device):
    free, _ = torch.cuda.mem_get_info(cuda_device)
    if free < 60 * 2 ** 30:
        pytest.skip("needs ~45 GB of free HBM")
    N, d, k, B = 10_000_000, 1024, 100, 32
    idx, qp, pos = build_planted(cuda_device, N, d, 3, 100, seed=3, chunk=1_000_000)
    dev = qp.device
    g = torch.Generator(device=dev).manual_seed(9)
    q = torch.randn((B, d), generator=g, device=dev)
    q /= q.norm(dim=1, keepdim=True) + 1e-8
    q[:3] = qp
    s, i = idx.search(q, k)
    torch.cuda.synchronize()
    sn, inn = check_common(s, i, N)
    for b in range(3):
        assert set(inn[b].tolist()) == set(pos[b].tolist())
    # partition property: two half shards + ts_topk_merge == the full scan (what 2 GPUs compute)
    half = N // 2
    parts = []
    for lo, hi in ((0, half), (half, N)):
        sh = _lib.Index(d, "bf16", "ip", cuda_device, reserve_rows=hi - lo)
        for c in range(lo, hi, 1_000_000):
            sh.add(torch.from_numpy(idx.get_rows(c, min(1_000_000, hi - c))).to(dev).to(torch.bfloat16))
        sh.set_id_base(lo)
        parts.append(sh.search(q, k))
        del sh
    ms, mi = _lib.topk_merge(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]), cuda_device)
    torch.cuda.synchronize()
    assert torch.equal(mi, i) and torch.equal(ms, s)


def test_config5_shapes_stage1_k500_into_stage2(cuda_device):
    """BASELINE config #5 shapes on one GPU at reduced corpus size: Stage 1 k=500 over d=768,
    Stage 2 over those 500 candidates (Ld <= 192), keep 100."""
    N, d, k1, k2, B, dim, Lq = 300_000, 768, 500, 100, 8, 128, 32
    dev = torch.device("cuda", cuda_device)
    rng = np.random.default_rng(55)
    X = flat_ip.normalize_rows(rng.standard_normal((N, d)).astype(np.float32)).astype(np.float32)
    idx = _lib.Index(d, "bf16", "ip", cuda_device)
    idx.add(X)
    lens = rng.integers(16, 193, size=N).astype(np.int32)
    g = torch.Generator(device=dev).manual_seed(55)
    st = _lib.TokStore(dim, "bf16", cuda_device, reserve_docs=N, reserve_tokens=int(lens.sum()))
    for s0 in range(0, N, 100_000):
        ln = lens[s0:s0 + 100_000]
        t = torch.nn.functional.normalize(torch.randn((int(ln.sum()), dim), generator=g, device=dev), dim=-1)
        st.add(t.to(torch.bfloat16), ln, normalize=False)
    Q = flat_ip.normalize_rows(rng.standard_normal((B, d)).astype(np.float32)).astype(np.float32)
    s1, i1 = idx.search(torch.from_numpy(Q).to(dev), k1)
    Xr, Qr = flat_ip.round_to(X, "bf16"), flat_ip.round_to(Q, "bf16")
    rD, rI = flat_ip.topk_desc(Qr @ Xr.T, k1)
    assert not flat_ip.check_topk(s1.cpu().numpy(), i1.cpu().numpy(),
                                  lambda b, ids: Xr[ids].astype(np.float64) @ Qr[b].astype(np.float64), rD, rI)
    qt = torch.nn.functional.normalize(torch.randn((B, Lq, dim), generator=g, device=dev), dim=-1).to(torch.bfloat16)
    s2 = st.maxsim(qt, i1, normalize_q=False)
    top_s, top_p = _lib.rank_desc(s2, k2, device=cuda_device)
    torch.cuda.synchronize()
    s2h, i1h = s2.cpu().numpy(), i1.cpu().numpy()
    off = np.concatenate([[0], np.cumsum(((lens + 7) // 8) * 8)])   # store rows are padded to 8
    for b in (0, B - 1):
        order = maxsim.rescore_order(s2h[b], k2)
        assert top_p[b].cpu().tolist() == order.tolist()
    assert (s2h > 0).all() and s2h.shape == (B, k1)
    assert st.ndocs == N and int(off[-1]) >= int(lens.sum())


def test_scan_variants_agree_bit_for_bit(cuda_device, monkeypatch):
    """The scan with and without cross-CTA threshold sharing must return identical results: the
    shared bound only prunes rows that cannot be in the top-k; so must the scan layouts."""
    N, d, B, k = 60000, 256, 48, 100
    X, Q = make(N, d, B, seed=77, planted=20)
    idx = _lib.Index(d, "bf16", "ip", cuda_device)
    idx.add(X)
    base = idx.search_host(Q, k, path="umma")
    rD, rI, sc = oracle_search(X, Q, k, "bf16")
    assert not flat_ip.check_topk(base[0], base[1], sc, rD, rI, rel=REL)
    Q2 = make(10, d, 300, seed=78)[1]                 # 300 queries: three query tiles
    base2 = idx.search_host(Q2, k, path="umma")
    # defaults: one cooperative launch (TS_FUSE) and CTA pairs for B >= 129 (TS_PAIR); the alternatives -- two
    # launches, single-CTA tiles, no threshold sharing, two query tiles per CTA -- must give the same bits
    variants = [("TS_DBG_NOSHARE", "1"), ("TS_FUSE", "0"), ("TS_PAIR", "0")]
    for var, val in variants:
        monkeypatch.setenv(var, val)
        D, I = idx.search_host(Q, k, path="umma")
        D2, I2 = idx.search_host(Q2, k, path="umma")
        monkeypatch.delenv(var)
        assert (I == base[1]).all() and (D == base[0]).all(), var
        assert (I2 == base2[1]).all() and (D2 == base2[0]).all(), var


@pytest.mark.parametrize("order", ["ascending", "descending", "constant"])
@pytest.mark.parametrize("k", [100, 500])
def test_adversarial_score_orders_on_device(cuda_device, order, k):
    """Rows are multiples of one direction, so the score of row i is a chosen value: ascending
    order makes every later row beat the running threshold (list overflow + in-scan prune),
    'constant' makes every row tie (k smallest ids must win).  tests/test_topk_model.py checks
    the same cases on the CPU model of the selection algorithm."""
    N, d = 60_000, 64
    rng = np.random.default_rng(4)
    u = flat_ip.normalize_rows(rng.standard_normal((1, d)).astype(np.float32))[0].astype(np.float32)
    v = {"ascending": np.linspace(0.05, 1.0, N), "descending": np.linspace(1.0, 0.05, N),
         "constant": np.full(N, 0.5)}[order].astype(np.float32)
    X = (v[:, None] * u[None, :]).astype(np.float32)
    Q = np.stack([u, -u, flat_ip.normalize_rows(rng.standard_normal((1, d)).astype(np.float32))[0]]).astype(np.float32)
    idx = _lib.Index(d, "bf16", "ip", cuda_device)
    idx.add(X)
    for path in ("umma", "stream"):
        D, I = idx.search_host(Q, k, path=path)
        rD, rI, sc = oracle_search(X, Q, k, "bf16")
        bad = flat_ip.check_topk(D, I, sc, rD, rI, rel=REL)
        assert not bad, (order, k, path, bad[:3])
        if order == "constant":
            # every row has the same stored value: exact ties, ids ascending
            assert I[0].tolist() == list(range(k)) and I[1].tolist() == list(range(k))


def test_tokstore_save_load_roundtrip(cuda_device, tmp_path):
    rng = np.random.default_rng(8)
    dim, ndocs = 128, 500
    lens = rng.integers(1, 200, size=ndocs)
    tok = rng.standard_normal((int(lens.sum()), dim)).astype(np.float32)
    st = _lib.TokStore(dim, "bf16", cuda_device)
    st.add(tok, lens, normalize=True)
    st.set_id_base(40)
    q = rng.standard_normal((3, 32, dim)).astype(np.float32)
    cand = (rng.integers(0, ndocs, size=(3, 64)) + 40).astype(np.int64)
    ref = st.maxsim_host(q, cand)
    p = str(tmp_path / "tok.tstok")
    st.save(p)
    st2 = _lib.TokStore.load(p, dim, "bf16", cuda_device)
    assert st2.ndocs == ndocs and st2.ntokens == int(lens.sum())
    got = st2.maxsim_host(q, cand)                    # id_base travels with the file
    assert np.array_equal(got, ref) and (ref != 0).all()
    with pytest.raises(_lib.TristageError):
        _lib.TokStore.load(str(tmp_path / "missing"), dim, "bf16", cuda_device)


def test_batched_rescoring_equals_per_query(cuda_device):
    """ColBERTScorer.rescore_candidates_batch (one launch for the batch) == the per-query call."""
    from oracle import fakes
    from tristage_rag_b200 import ColBERTScorer, Stage2Config

    docs = [f"document number {i} talks about topic {i % 7} and item {i * 3 % 11} in some detail" for i in range(60)]
    queries = ["topic 3 item 5", "document number 12", "what about detail", ""]
    tok = fakes.FakeTokenizer()
    sc = ColBERTScorer(Stage2Config(device="cpu", top_k_candidates=10, gpu_index=cuda_device), tokenizer=tok,
                       model=fakes.FakeTokenModel(tok, 64))
    rng = np.random.default_rng(3)
    cands = []
    for n in (25, 7, 0, 60):
        ids = rng.choice(60, size=n, replace=False)
        cands.append([{"doc_id": int(i), "document": docs[int(i)], "score": 1.0} for i in ids])
    batch = sc.rescore_candidates_batch(queries, cands)
    for q, c, got in zip(queries, cands, batch):
        ref = sc.rescore_candidates(q, c)
        assert [x["doc_id"] for x in got] == [x["doc_id"] for x in ref]
        assert [x["stage2_score"] for x in got] == pytest.approx([x["stage2_score"] for x in ref], rel=1e-6, abs=1e-7)
        assert all(x["stage"] == "stage2" for x in got) and len(got) == min(10, len(c))
