#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU, NCCL):
row-sharded Stage-1 search + all-gather + merge, and ownership-filtered Stage-2
scoring + all-reduce, against the CPU oracle on the full data."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import flat_ip, maxsim  # noqa: E402
from tristage_rag_b200 import _lib  # noqa: E402
from tristage_rag_b200.dist import ShardedIndex, ShardedTokStore, shard_range  # noqa: E402


def main():
    import bench

    bench.arm_watchdog(int(os.environ.get("TS_CHECK_TIMEOUT", "240")))     # a wedged exchange must not hold the box
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rng = np.random.default_rng(123)                      # same data on every rank
    N, d, k = 50_001, 256, 100
    X = flat_ip.normalize_rows(rng.standard_normal((N, d)).astype(np.float32)).astype(np.float32)
    Xr = flat_ip.round_to(X, "bf16")
    lo, hi = shard_range(N, rank, world)
    idx = _lib.Index(d, "bf16", "ip", local)
    idx.add(X[lo:hi])
    sh = ShardedIndex(idx, N)
    for B in (1, 32, 200):
        Q = flat_ip.normalize_rows(rng.standard_normal((B, d)).astype(np.float32)).astype(np.float32)
        Qr = flat_ip.round_to(Q, "bf16")
        s, i = sh.search(torch.from_numpy(Q).to(dev), k)
        torch.cuda.synchronize()
        rD, rI = flat_ip.topk_desc(Qr @ Xr.T, k)
        bad = flat_ip.check_topk(s.cpu().numpy(), i.cpu().numpy(),
                                 lambda b, ids: Xr[ids].astype(np.float64) @ Qr[b].astype(np.float64), rD, rI)
        assert not bad, (rank, B, bad[:3])
        # every rank holds the identical merged result
        ref = [torch.empty_like(i) for _ in range(world)]
        dist.all_gather(ref, i)
        assert all((r == ref[0]).all() for r in ref)
        # the host-buffer call (one C call per step on the peer-memory plane) returns the same
        hD, hI = sh.search_host(Q, k)
        assert (hI == i.cpu().numpy()).all() and (hD == s.cpu().numpy()).all(), (rank, B, "host call")
        # one kernel for the whole exchange (default when B <= SM count) and the two-kernel form agree bit for bit
        if sh._p2p:
            os.environ["TS_XFUSE"] = "0"
            s2, i2 = sh.search(torch.from_numpy(Q).to(dev), k)
            del os.environ["TS_XFUSE"]
            torch.cuda.synchronize()
            assert torch.equal(i2, i) and torch.equal(s2, s), (rank, B, "two-kernel exchange")
    # approximate mode: the same centroids on every rank, lists over the local rows -> the merged result is what
    # one GPU holding all rows returns for the same probes (oracle/ivf.py on the full data)
    from oracle import ivf as oivf
    from tristage_rag_b200.dist import IVFShard
    from tristage_rag_b200.ivf import train_centroids

    nlist, nprobe = 50, 5
    cent = train_centroids(X[:20000], nlist)
    iv = _lib.IVF(idx, nlist)
    iv.set_centroids(cent)
    iv.sync()
    shi = ShardedIndex(IVFShard(iv, nprobe), N)
    parts = [None] * world                       # the lists the kernels built (near-tied rows may differ from fp64)
    dist.all_gather_object(parts, iv.assignments())
    assign = np.concatenate(parts)
    assert assign.shape == (N,) and (assign != oivf.assign_lists(Xr, cent)).mean() < 1e-3
    Q = flat_ip.normalize_rows(X[rng.integers(0, N, size=16)] + 0.05 * rng.standard_normal((16, d)).astype(np.float32))
    Q = Q.astype(np.float32)
    Qr = flat_ip.round_to(Q, "bf16")
    s, i = shi.search(torch.from_numpy(Q).to(dev), k)
    torch.cuda.synchronize()
    lists, _ = iv.coarse_host(Q, nprobe)
    rD, rI = oivf.ivf_search(Xr, Qr, assign, lists, k)
    bad = flat_ip.check_topk(s.cpu().numpy(), i.cpu().numpy(),
                             lambda b, ids: Xr[ids].astype(np.float64) @ Qr[b].astype(np.float64), rD, rI)
    assert not bad, (rank, "ivf", bad[:3])
    # Stage 2
    ndocs, dim = 3000, 128
    lens = rng.integers(16, 181, size=ndocs)
    tok = rng.standard_normal((int(lens.sum()), dim)).astype(np.float32)
    off = np.concatenate([[0], np.cumsum(lens)])
    dlo, dhi = shard_range(ndocs, rank, world)
    st = _lib.TokStore(dim, "bf16", local)
    st.add(tok[off[dlo]:off[dhi]], lens[dlo:dhi], normalize=True)
    sst = ShardedTokStore(st, ndocs)
    q = rng.standard_normal((4, 32, dim)).astype(np.float32)
    cand = rng.integers(0, ndocs, size=(4, 200)).astype(np.int64)
    got = sst.maxsim(torch.from_numpy(q).to(dev), torch.from_numpy(cand).to(dev)).cpu().numpy()
    nr = lambda x: flat_ip.round_to(maxsim.l2_normalize_tokens(x), "bf16")  # noqa: E731
    ref = np.array([[maxsim.maxsim_score(nr(q[b]), nr(tok[off[c]:off[c + 1]]), normalize=False) for c in cand[b]]
                    for b in range(4)])
    assert np.allclose(got, ref, rtol=1e-3, atol=2e-4), float(np.abs(got - ref).max())
    # the scatter fused into the scoring kernel (default on CUDA) and the all-reduce of the shards' outputs give the same
    # matrix, bit for bit, over several steps (both parities of the receive buffer, ids nobody owns, ragged n_cand)
    qd = torch.from_numpy(q).to(dev)
    for step in range(5):
        cand2 = torch.from_numpy(rng.integers(-3, ndocs + 5, size=(4, 200)).astype(np.int64)).to(dev)
        ncd = torch.from_numpy(rng.integers(50, 201, size=4).astype(np.int32)).to(dev)
        a1 = sst.maxsim(qd, cand2, n_cand=ncd)
        os.environ["TS_S2_SCATTER"] = "0"
        saved = os.environ.pop("TS_P2P", None)
        os.environ["TS_P2P"] = "0"
        a2 = sst.maxsim(qd, cand2, n_cand=ncd)
        del os.environ["TS_S2_SCATTER"], os.environ["TS_P2P"]
        if saved is not None:
            os.environ["TS_P2P"] = saved
        torch.cuda.synchronize()
        assert torch.equal(a1, a2), (rank, step, float((a1 - a2).abs().max()))
    stage2_plane = "scatter fused into the scoring kernel" if sst._scatter_ok(qd) else "all-reduce / push + sum"
    # load balance of candidate ownership
    own = np.bincount(np.searchsorted([shard_range(ndocs, r, world)[1] for r in range(world)], cand.ravel(), side="right"),
                      minlength=world)
    # sharded persistence: every rank writes its shard, rank 0 the manifest; reload under the same world size
    # and as ONE shard on rank 0's GPU (re-sharding), both must search like the live index
    import tempfile

    box = [tempfile.mkdtemp(prefix="ts_shards_") if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    sh.save(box[0])
    again = ShardedIndex.load(box[0], local)
    Q = flat_ip.normalize_rows(rng.standard_normal((8, d)).astype(np.float32)).astype(np.float32)
    qd = torch.from_numpy(Q).to(dev)
    s0, i0 = sh.search(qd, k)
    s1, i1 = again.search(qd, k)
    torch.cuda.synchronize()
    assert torch.equal(i0, i1) and torch.equal(s0, s1)
    if rank == 0:
        from tristage_rag_b200 import dist as tdist

        man = tdist.read_manifest(box[0], "index")
        one = _lib.Index(d, "bf16", "ip", local, reserve_rows=N)
        for fname, first, n in tdist.plan_reshard(man["shards"], 0, N):
            one.append_file(os.path.join(box[0], fname), first, n)
        s2, i2 = one.search(qd, k)
        torch.cuda.synchronize()
        assert torch.equal(i2, i0) and torch.equal(s2, s0)
    # the drop-in classes striped over the group (Stage1Config.sharded / Stage2Config.sharded): every rank makes the calls the
    # reference's orchestrator makes and gets what the un-sharded classes return (fake encoders: no weights on the box)
    from oracle import fakes
    from tristage_rag_b200 import ColBERTScorer, Stage1Config, Stage1Retriever, Stage2Config

    docs = [f"document number {i} talks about topic {i % 7} and item {i * 3 % 11} in some detail {i}" for i in range(40)]
    queries = ["topic 3 item 5", "document number 12 in detail", "item 9"]
    with tempfile.TemporaryDirectory() as tmp:
        kw = dict(device="cpu", cache_dir=os.path.join(tmp, "m"), index_dir=os.path.join(tmp, "i"), top_k_candidates=12,
                  enable_bm25=False, storage_dtype="fp32", gpu_index=local)
        one = Stage1Retriever(Stage1Config(**kw), model=fakes.FakeSentenceEncoder(64))
        many = Stage1Retriever(Stage1Config(sharded=True, **kw), model=fakes.FakeSentenceEncoder(64))
        for part in (docs[:1], docs[1:9], docs[9:]):
            one.add_documents(list(part))
            many.add_documents(list(part))
        tok = fakes.FakeTokenizer()
        s_one = ColBERTScorer(Stage2Config(device="cpu", top_k_candidates=6, gpu_index=local), tokenizer=tok, model=fakes.FakeTokenModel(tok, 32))
        s_many = ColBERTScorer(Stage2Config(device="cpu", top_k_candidates=6, gpu_index=local, sharded=True), tokenizer=tok,
                               model=fakes.FakeTokenModel(tok, 32))
        for qq in queries:
            a, b2 = one.search(qq), many.search(qq)
            assert [x["doc_id"] for x in a] == [x["doc_id"] for x in b2] and len(b2) == 12, (rank, qq)
            assert np.allclose([x["score"] for x in a], [x["score"] for x in b2], rtol=1e-5)
            ra, rb = s_one.rescore_candidates(qq, a), s_many.rescore_candidates(qq, b2)
            assert [x["doc_id"] for x in ra] == [x["doc_id"] for x in rb] and len(rb) == 6, (rank, qq)
            assert np.allclose([x["stage2_score"] for x in ra], [x["stage2_score"] for x in rb], rtol=1e-3, atol=2e-4)
        assert many.faiss_index.local.ntotal < len(docs) and s_many._store.ndocs < s_one._store.ndocs
    dist.barrier()
    if rank == 0:
        mode = "peer-memory exchange (fused select+push, wait+merge)" if sh._p2p else "nccl all-gather + merge kernel"
        print(f"dist_check ok: world={world} merge via {mode}; stage2 via {stage2_plane}; sharded drop-in classes ok; stage2 candidates per rank max/mean = {own.max() / own.mean():.3f}",
              flush=True)
    sys.stdout.flush()
    os._exit(0)          # no NCCL teardown (it can block for minutes after the work is done)


if __name__ == "__main__":
    main()
