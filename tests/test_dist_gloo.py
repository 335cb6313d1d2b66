"""CPU, world_size 2, gloo: the sharding/exchange plumbing of dist.py (shard
ranges, global ids, rank-major all-gather layout, merge, ownership-filtered
Stage-2 sum).  The per-rank compute is stood in by the oracle here (test
infrastructure); the GPU tests run the same wrappers over libtristage."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import flat_ip, maxsim
from tristage_rag_b200 import dist as tdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _OracleLocalIndex:
    def __init__(self, X):
        self.X, self.base = X, 0

    def set_id_base(self, b):
        self.base = b

    def search(self, q, k):
        D, I = flat_ip.topk_desc(q.numpy() @ self.X.T, k)
        I = np.where(I >= 0, I + self.base, -1)
        return torch.from_numpy(D), torch.from_numpy(I)


class _OracleLocalIVF:
    """What _lib.IVF offers dist.IVFShard (base.set_id_base, search(q, k, nprobe)), over oracle/ivf.py."""

    class _Base:
        def __init__(self):
            self.id_base = 0

        def set_id_base(self, b):
            self.id_base = b

    def __init__(self, X, cent):
        from oracle import ivf as oivf

        self.X, self.cent, self.base, self.oivf = X, cent, self._Base(), oivf
        self.assign = oivf.assign_lists(X, cent)

    def search(self, q, k, nprobe):
        lists, _ = self.oivf.coarse_probe(q.numpy(), self.cent, nprobe)
        D, I = self.oivf.ivf_search(self.X, q.numpy(), self.assign, lists, k)
        return torch.from_numpy(D), torch.from_numpy(np.where(I >= 0, I + self.base.id_base, -1))


def _numpy_merge(all_s, all_i):
    G, B, k = all_s.shape
    s = all_s.permute(1, 0, 2).reshape(B, G * k).numpy()
    i = all_i.permute(1, 0, 2).reshape(B, G * k).numpy()
    outD = np.full((B, k), flat_ip.LOWEST_F32, np.float32)
    outI = np.full((B, k), -1, np.int64)
    for b in range(B):
        ok = i[b] >= 0
        order = np.lexsort((i[b][ok], -s[b][ok].astype(np.float64)))[:k]
        outD[b, : len(order)] = s[b][ok][order]
        outI[b, : len(order)] = i[b][ok][order]
    return torch.from_numpy(outD), torch.from_numpy(outI)


class _OracleLocalStore:
    def __init__(self, docs, lo):
        self.docs, self.base = docs, lo

    def set_id_base(self, b):
        self.base = b

    def maxsim(self, q_tok, cand):
        out = torch.zeros(cand.shape, dtype=torch.float32)
        for b in range(cand.shape[0]):
            for j, cid in enumerate(cand[b].tolist()):
                loc = cid - self.base
                if 0 <= loc < len(self.docs):
                    out[b, j] = maxsim.maxsim_score(q_tok[b].numpy(), self.docs[loc])
        return out


def _worker(rank, world, port, n_total, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        X = rng.standard_normal((n_total, 24)).astype(np.float32)
        Q = rng.standard_normal((3, 24)).astype(np.float32)
        lo, hi = tdist.shard_range(n_total, rank, world)
        idx = tdist.ShardedIndex(_OracleLocalIndex(X[lo:hi]), n_total, merge_fn=_numpy_merge)
        assert (idx.lo, idx.hi) == (lo, hi)
        for k in (7, 40):
            D, I = idx.search(torch.from_numpy(Q), k)
            rD, rI = flat_ip.topk_desc(Q @ X.T, k)
            assert (I.numpy() == rI).all() and np.allclose(D.numpy(), rD)
        # approximate mode: same centroids on every rank, lists over the local rows -> merged result == one index
        from oracle import ivf as oivf

        cent = X[:5].copy()
        shard = tdist.ShardedIndex(tdist.IVFShard(_OracleLocalIVF(X[lo:hi], cent), nprobe=2), n_total, merge_fn=_numpy_merge)
        D, I = shard.search(torch.from_numpy(Q), 7)
        lists, _ = oivf.coarse_probe(Q, cent, 2)
        rD, rI = oivf.ivf_search(X, Q, oivf.assign_lists(X, cent), lists, 7)
        assert (I.numpy() == rI).all() and np.allclose(D.numpy(), rD)
        # Stage 2: ownership-filtered scores, summed across ranks
        lens = rng.integers(2, 9, size=n_total)
        docs = [rng.standard_normal((int(L), 8)).astype(np.float32) for L in lens]
        qt = torch.from_numpy(rng.standard_normal((2, 4, 8)).astype(np.float32))
        cand = torch.from_numpy(rng.integers(0, n_total, size=(2, 6)).astype(np.int64))
        st = tdist.ShardedTokStore(_OracleLocalStore(docs[lo:hi], lo), n_total)
        got = st.maxsim(qt, cand).numpy()
        ref = np.array([[maxsim.maxsim_score(qt[b].numpy(), docs[c]) for c in cand[b].tolist()] for b in range(2)])
        assert np.allclose(got, ref, atol=1e-6)
        ret[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [41, 64])
def test_sharded_search_world2_gloo(n_total):
    world, port = 2, _free_port()
    ret = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, port, n_total, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 10_000_000):
        for w in (1, 2, 4, 8):
            spans = [tdist.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    ids = torch.arange(0, 103)
    own = tdist.owner_of(ids, 103, 4)
    for r in range(4):
        lo, hi = tdist.shard_range(103, r, 4)
        assert (own[lo:hi] == r).all()


# ------------------------------------------------- the drop-in classes, sharded (Stage1Config.sharded) ---
def _dropin_worker(rank, world, port, ret):
    """Every rank makes the SAME calls on the drop-in classes; the kernels run on the CPU emulator build here
    (test infrastructure), the exchange over gloo.  Results must equal the un-sharded classes' on every rank."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import tempfile

    import conftest
    from oracle import fakes
    from tristage_rag_b200 import ColBERTScorer, Stage1Config, Stage1Retriever, Stage2Config

    conftest._enter_emulation()
    for var in ("TS_PAIR", "TS_FUSE", "TS_TF32"):
        os.environ[var] = "0"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        docs = [f"document number {i} talks about topic {i % 7} and item {i * 3 % 11} in some detail {i}" for i in range(23)]
        queries = ["topic 3 item 5", "document number 12 in detail", "item 9"]
        with tempfile.TemporaryDirectory() as tmp:
            kw = dict(device="cpu", cache_dir=os.path.join(tmp, "m"), index_dir=os.path.join(tmp, "i"), top_k_candidates=12,
                      enable_bm25=False, storage_dtype="fp32")
            one = Stage1Retriever(Stage1Config(**kw), model=fakes.FakeSentenceEncoder(64))
            many = Stage1Retriever(Stage1Config(sharded=True, **kw), model=fakes.FakeSentenceEncoder(64))
            many._shard_merge_fn = _numpy_merge
            for part in (docs[:1], docs[1:9], docs[9:]):          # a 1-doc batch leaves rank 1 without rows at first
                one.add_documents(list(part))
                many.add_documents(list(part))
                assert many.faiss_index.ntotal == one.faiss_index.ntotal == len(one.documents)
                a, b = one.search(queries[0], 5), many.search(queries[0], 5)
                assert [x["doc_id"] for x in a] == [x["doc_id"] for x in b]
            assert 0 < many.faiss_index.local.ntotal < len(docs)  # this rank keeps only its share
            tok = fakes.FakeTokenizer()
            s_one = ColBERTScorer(Stage2Config(device="cpu", top_k_candidates=6, storage_dtype="fp32"), tokenizer=tok,
                                  model=fakes.FakeTokenModel(tok, 32))
            s_many = ColBERTScorer(Stage2Config(device="cpu", top_k_candidates=6, storage_dtype="fp32", sharded=True),
                                   tokenizer=tok, model=fakes.FakeTokenModel(tok, 32))
            for q in queries:
                a, b = one.search(q), many.search(q)
                assert [x["doc_id"] for x in a] == [x["doc_id"] for x in b] and len(b) == 12
                assert np.allclose([x["score"] for x in a], [x["score"] for x in b], rtol=1e-6)
                ra, rb = s_one.rescore_candidates(q, a), s_many.rescore_candidates(q, b)
                assert [x["doc_id"] for x in ra] == [x["doc_id"] for x in rb] and len(rb) == 6
                assert np.allclose([x["stage2_score"] for x in ra], [x["stage2_score"] for x in rb], rtol=1e-6)
            assert s_many._store.ndocs < s_one._store.ndocs      # token embeddings live on their owner only
            ba = s_one.rescore_candidates_batch(queries, [one.search(q) for q in queries])
            bb = s_many.rescore_candidates_batch(queries, [many.search(q) for q in queries])
            assert [[x["doc_id"] for x in r] for r in ba] == [[x["doc_id"] for x in r] for r in bb]
            m1 = s_one.compute_similarity_matrix(queries[1], docs[:10])
            m2 = s_many.compute_similarity_matrix(queries[1], docs[:10])
            assert np.allclose(m1, m2, rtol=1e-6)
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_drop_in_classes_sharded_world2_gloo():
    world, port = 2, _free_port()
    ret = mp.get_context("spawn").Manager().dict()
    mp.spawn(_dropin_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))
