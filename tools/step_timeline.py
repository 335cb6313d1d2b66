#!/usr/bin/env python
"""GPU time line of one Stage-1 step (TS_DBG_TIMELINE): where the microseconds of a (multi-GPU) search step go.
Run on one GPU or under torchrun; every rank runs `--steps` searches back to back, then the globaltimer stamps
of its LAST step are read (ts_index_debug_timeline) and rank 0 prints one JSON line per rank plus the median over
ranks:  prep (query-prep kernel) | gap | scan (first CTA in .. last CTA out) | gap | select: local sort | push +
publish | wait for the peers' rows | merge.  The stamps cost one store / atomic per kernel.  Development aid."""
import argparse
import json
import os
import statistics
import sys

os.environ["TS_DBG_TIMELINE"] = "1"
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from tristage_rag_b200 import _lib  # noqa: E402
from tristage_rag_b200.dist import ShardedIndex, shard_range  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--samples", type=int, default=5)
    args = ap.parse_args()
    bench.arm_watchdog(300)
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lo, hi = shard_range(args.rows, rank, world)
    idx = _lib.Index(args.dim, "bf16", "ip", local, reserve_rows=hi - lo)
    bench.build_shard(idx, lo, hi, args.dim, dev, seed=1234 + rank)
    sh = ShardedIndex(idx, args.rows)
    _, q = bench.make_queries(args.batch, args.dim, dev)
    names = ["prep", "gap_prep_scan", "scan", "gap_scan_select", "select_sort", "push_publish", "wait_peers", "merge", "step"]
    rows = []
    for _ in range(args.samples):
        for _ in range(args.steps):
            sh.search(q, args.k)
        t = [int(x) for x in idx.debug_timeline()]
        d = lambda a, b: (t[b] - t[a]) / 1e3 if t[a] and t[b] else None      # noqa: E731
        last = 8 if t[8] else (6 if t[6] else 5)
        rows.append([d(0, 1), d(1, 2), d(2, 3), d(3, 4), d(4, 5), d(5, 6), d(6, 7), d(7, 8), d(0, last)])
        if world > 1:
            dist.barrier()
    med = [None if any(r[i] is None for r in rows) else round(statistics.median(r[i] for r in rows), 2) for i in range(len(names))]
    rec = {"rank": rank, **dict(zip(names, med))}
    if world > 1:
        allr = [None] * world
        dist.all_gather_object(allr, rec)
    else:
        allr = [rec]
    if rank == 0:
        for r in allr:
            print(json.dumps(r))
        summary = {"world": world, "rows_per_rank": hi - lo, "batch": args.batch, "k": args.k, "exchange": "one kernel" if getattr(sh, "_p2p", False) else ("nccl" if world > 1 else "none"),
                   "median_over_ranks_us": {n: (None if any(r[n] is None for r in allr) else round(statistics.median(r[n] for r in allr), 2)) for n in names},
                   "max_over_ranks_us": {n: (None if any(r[n] is None for r in allr) else round(max(r[n] for r in allr), 2)) for n in names}}
        print(json.dumps(summary), flush=True)
    bench.finish(0)


if __name__ == "__main__":
    main()
