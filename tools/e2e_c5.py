#!/usr/bin/env python
"""BASELINE config #5 end to end: Stage 1 (k=500) -> Stage 2 over those 500 candidates
(Ld ~ U[16,192], dim 128) -> keep 100, on a synthetic corpus row-sharded over the ranks
(torchrun, one rank per GPU, NCCL).  5 M docs need 8 GPUs for the token store (~133 GB);
smaller --docs fit fewer.  Prints one JSON line on rank 0.  NOT part of bench.py's contract.

  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/e2e_c5.py
  python tools/e2e_c5.py --docs 500000            # single GPU
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from tristage_rag_b200 import _lib  # noqa: E402
from tristage_rag_b200.dist import ShardedIndex, ShardedTokStore, shard_range  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=5_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--tok-dim", type=int, default=128)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--k1", type=int, default=500)
    ap.add_argument("--k2", type=int, default=100)
    ap.add_argument("--lq", type=int, default=32)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    bench.arm_watchdog(900)
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N = args.docs
    lo, hi = shard_range(N, rank, world)
    idx = _lib.Index(args.dim, "bf16", "ip", local, reserve_rows=hi - lo)
    bench.build_shard(idx, lo, hi, args.dim, dev, seed=1234 + rank)
    rng = np.random.default_rng(77 + rank)
    lens = rng.integers(16, 193, size=hi - lo).astype(np.int32)
    st = _lib.TokStore(args.tok_dim, "bf16", local, reserve_docs=hi - lo, reserve_tokens=int(lens.sum()))
    g = torch.Generator(device=dev).manual_seed(77 + rank)
    for s in range(0, hi - lo, 100_000):
        ln = lens[s:s + 100_000]
        t = torch.nn.functional.normalize(torch.randn((int(ln.sum()), args.tok_dim), generator=g, device=dev), dim=-1)
        st.add(t.to(torch.bfloat16), ln, normalize=False)
        del t
    sidx, sst = ShardedIndex(idx, N), ShardedTokStore(st, N)
    _, q = bench.make_queries(args.batch, args.dim, dev)
    gq = torch.Generator(device=dev).manual_seed(5)
    qt = torch.nn.functional.normalize(torch.randn((args.batch, args.lq, args.tok_dim), generator=gq, device=dev), dim=-1).to(torch.bfloat16)

    def step():
        s1, i1 = sidx.search(q, args.k1)                       # exact top-500, merged across ranks
        s2 = sst.maxsim(qt, i1, normalize_q=False)             # owners score, all-reduce sums
        return _lib.rank_desc(s2, args.k2, device=local)       # stable top-100 per query

    ms = bench.timed(step, args.steps, args.warmup, dev, world > 1)
    if rank == 0:
        print(json.dumps({"config": "C5 end-to-end Stage1 k=%d -> Stage2 -> top-%d" % (args.k1, args.k2), "docs": N,
                          "dim": args.dim, "tok_dim": args.tok_dim, "batch": args.batch, "n_gpus": world,
                          "ms_per_step": ms / args.steps, "queries_per_s": args.batch * args.steps / (ms / 1e3),
                          "token_store_gb_per_gpu": float(lens.astype(np.int64).sum()) * args.tok_dim * 2 / 1e9}), flush=True)
    bench.finish(0)


if __name__ == "__main__":
    main()
