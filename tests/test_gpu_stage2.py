"""GPU parity: Stage-2 MaxSim / colbert scoring through the C ABI vs the
reference's own golden values and the CPU oracle."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import flat_ip, maxsim
from tristage_rag_b200 import _lib

pytestmark = pytest.mark.gpu
SIMT = _lib.TS_S2_FORCE_SIMT


def _norm_round(x, dtype):
    """what the store holds: F.normalize in fp32, then the storage rounding"""
    return flat_ip.round_to(maxsim.l2_normalize_tokens(x), dtype)


def make_store(cuda_device, lens, dim, dtype, seed=0, scale=1.0):
    rng = np.random.default_rng(seed)
    tok = (rng.standard_normal((int(np.sum(lens)), dim)) * scale).astype(np.float32)
    st = _lib.TokStore(dim, dtype, cuda_device)
    # several adds: exercises growth and offset bookkeeping
    cut = len(lens) // 3
    off = np.concatenate([[0], np.cumsum(lens)])
    for a, b in ((0, cut), (cut, len(lens))):
        if b > a:
            st.add(tok[off[a]:off[b]], lens[a:b], normalize=True)
    docs = [tok[off[i]:off[i + 1]] for i in range(len(lens))]
    return st, docs


def oracle_scores(q, docs, cand, mode, dtype, n_cand=None, q_len=None):
    B, C = cand.shape
    out = np.zeros((B, C), np.float32)
    for b in range(B):
        lq = q.shape[1] if q_len is None else int(q_len[b])
        qb = _norm_round(q[b, :lq], dtype)
        for j in range(C if n_cand is None else int(n_cand[b])):
            c = int(cand[b, j])
            if 0 <= c < len(docs):
                out[b, j] = maxsim.score(qb, _norm_round(docs[c], dtype), mode, normalize=False)
    return out


@pytest.mark.parametrize("force", [0, SIMT])
def test_reference_golden_pairs(cuda_device, golden_dir, force):
    """The reference's own _maxsim_score/_colbert_score values (gen_golden.py)."""
    with open(os.path.join(golden_dir, "stage2_reference.json")) as f:
        g = json.load(f)
    for c in g["cases"]:
        if c["kind"] != "numpy":
            continue
        r = np.random.default_rng(c["seed"])
        q = r.standard_normal((c["Lq"], c["H"])).astype(np.float32) * np.float32(c["q_scale"])
        d = r.standard_normal((c["Ld"], c["H"])).astype(np.float32) * np.float32(c["d_scale"])
        for dtype, tol in (("bf16", 2e-3), ("fp32", 2e-5)):
            st = _lib.TokStore(c["H"], dtype, cuda_device)
            st.add(d, [c["Ld"]], normalize=True)
            for mode, key in ((0, "maxsim"), (1, "colbert")):
                s = st.maxsim_host(q[None], np.zeros((1, 1), np.int64), mode=mode | force)[0, 0]
                assert s == pytest.approx(c[key], rel=tol, abs=tol), (c, dtype, key, force)


@pytest.mark.parametrize("dim,Lq,dtype", [(128, 32, "bf16"), (128, 32, "fp16"), (64, 7, "bf16"), (128, 1, "bf16"),
                                          (96, 33, "bf16"), (128, 128, "bf16"), (768, 32, "bf16"), (128, 70, "bf16")])
@pytest.mark.parametrize("mode", [0, 1])
def test_batch_vs_oracle_ragged(cuda_device, dim, Lq, dtype, mode):
    rng = np.random.default_rng(dim + Lq)
    ndocs, B, C = 400, 5, 97
    lens = rng.integers(1, 181, size=ndocs)
    lens[:6] = [1, 8, 9, 255, 256, 180]              # edge lengths
    st, docs = make_store(cuda_device, lens, dim, dtype, seed=dim)
    assert st.ndocs == ndocs and st.ntokens == int(lens.sum())
    q = rng.standard_normal((B, Lq, dim)).astype(np.float32)
    cand = rng.integers(0, ndocs, size=(B, C)).astype(np.int64)
    cand[0, :6] = np.arange(6)
    got = st.maxsim_host(q, cand, mode=mode)
    ref = oracle_scores(q, docs, cand, mode, dtype)
    np.testing.assert_allclose(got, ref, rtol=1e-3, atol=2e-4)
    simt = st.maxsim_host(q, cand, mode=mode | SIMT)
    np.testing.assert_allclose(simt, ref, rtol=1e-3, atol=2e-4)


def test_n_cand_q_len_and_unowned_ids(cuda_device):
    rng = np.random.default_rng(2)
    dim, ndocs, B, C, Lq = 128, 300, 6, 64, 32
    lens = rng.integers(16, 181, size=ndocs)
    st, docs = make_store(cuda_device, lens, dim, "bf16", seed=4)
    q = rng.standard_normal((B, Lq, dim)).astype(np.float32)
    cand = rng.integers(0, ndocs, size=(B, C)).astype(np.int64)
    cand[1, 5] = -1
    cand[2, 7] = ndocs + 10                          # not in this shard
    n_cand = np.array([64, 10, 0, 33, 1, 64], np.int32)
    q_len = np.array([32, 5, 32, 17, 1, 31], np.int32)
    for mode in (0, 1):
        got = st.maxsim_host(q, cand, q_len=q_len, n_cand=n_cand, mode=mode)
        ref = oracle_scores(q, docs, cand, mode, "bf16", n_cand=n_cand, q_len=q_len)
        np.testing.assert_allclose(got, ref, rtol=1e-3, atol=2e-4)
        assert got[1, 5] == 0.0 and (got[2] == 0.0).all() and (got[1, 10:] == 0.0).all()
    # ownership filter: shard with id_base -> scores of a split store sum to the whole
    st.set_id_base(1000)
    got = st.maxsim_host(q, cand + 1000, mode=0)
    ref = oracle_scores(q, docs, cand, 0, "bf16")
    np.testing.assert_allclose(got, ref, rtol=1e-3, atol=2e-4)
    assert (st.maxsim_host(q, cand, mode=0) == 0.0).all()


def test_baseline_c4_shape_and_rank(cuda_device):
    """BASELINE config #4 shape at reduced store size: 64 q x 1000 cand x (32 x <=180 x 128),
    device tensors, then the stable descending top-k (K6)."""
    rng = np.random.default_rng(77)
    dim, ndocs, B, C, Lq = 128, 20000, 64, 1000, 32
    lens = rng.integers(16, 181, size=ndocs)
    dev = torch.device("cuda", cuda_device)
    g = torch.Generator(device=dev).manual_seed(77)
    tok = torch.randn((int(lens.sum()), dim), generator=g, device=dev)
    tok = torch.nn.functional.normalize(tok, dim=-1).to(torch.bfloat16)
    st = _lib.TokStore(dim, "bf16", cuda_device, reserve_docs=ndocs, reserve_tokens=int(lens.sum()))
    st.add(tok, lens, normalize=False)               # device-resident, already normalised bf16
    q = torch.nn.functional.normalize(torch.randn((B, Lq, dim), generator=g, device=dev), dim=-1).to(torch.bfloat16)
    cand = torch.stack([torch.randperm(ndocs, generator=g, device=dev)[:C] for _ in range(B)])
    scores = st.maxsim(q, cand, normalize_q=False)
    top_s, top_p = _lib.rank_desc(scores, 100, device=cuda_device)
    torch.cuda.synchronize()
    tok_h = tok.float().cpu().numpy()
    off = np.concatenate([[0], np.cumsum(lens)])
    qh, ch, sh = q.float().cpu().numpy(), cand.cpu().numpy(), scores.cpu().numpy()
    for b in (0, 31, 63):
        ref = np.array([maxsim.maxsim_score(qh[b], tok_h[off[c]:off[c + 1]], normalize=False) for c in ch[b]])
        np.testing.assert_allclose(sh[b], ref, rtol=1e-3, atol=2e-4)
        order = maxsim.rescore_order(sh[b], 100)
        assert top_p[b].cpu().tolist() == order.tolist()
        assert np.array_equal(top_s[b].cpu().numpy(), sh[b][order])


def test_rank_desc_is_stable_and_pads(cuda_device):
    dev = torch.device("cuda", cuda_device)
    s = torch.tensor([[0.5, 0.9, 0.5, 0.9, 0.1, 0.0], [1.0, 1.0, 1.0, 1.0, 1.0, 1.0]], device=dev)
    n = torch.tensor([5, 2], dtype=torch.int32, device=dev)
    ts, tp = _lib.rank_desc(s, 4, n_cand=n, device=cuda_device)
    assert tp.cpu().tolist() == [[1, 3, 0, 2], [0, 1, -1, -1]]
    assert ts[0].cpu().tolist() == pytest.approx([0.9, 0.9, 0.5, 0.5])
