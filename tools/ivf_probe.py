#!/usr/bin/env python
"""Approximate Stage-1 mode (csrc/ivf.cu) on one GPU: time of the list scan vs batch size, with the
exact scan of the same shard next to it (development aid; one JSON line per configuration, inputs
resident in HBM, CUDA-event timed on the launching stream)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from tristage_rag_b200 import _lib  # noqa: E402
from tristage_rag_b200.ivf import train_centroids  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--nlist", type=int, default=100)       # the reference's defaults (Stage1Config.nlist / nprobe)
    ap.add_argument("--nprobe", type=int, default=10)
    ap.add_argument("--batches", default="1,8,32")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--tag", default="")
    ap.add_argument("--selfcheck", action="store_true",
                    help="device-vs-device consistency: all lists probed must equal the exact scan; probed rows must "
                         "belong to the probed lists")
    args = ap.parse_args()
    bench.arm_watchdog(300)
    dev = torch.device("cuda", 0)
    pk = bench.peaks()
    idx = _lib.Index(args.dim, "bf16", "ip", 0, reserve_rows=args.rows)
    bench.build_shard(idx, 0, args.rows, args.dim, dev, 1234)
    iv = _lib.IVF(idx, args.nlist)
    iv.set_centroids(train_centroids(idx.get_rows(0, min(args.rows, args.nlist * 256)), args.nlist))
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    iv.sync()
    t1.record()
    torch.cuda.synchronize()
    sizes = iv.list_sizes()
    print(json.dumps({"tag": args.tag, "what": "assign+lists", "rows": args.rows, "ms": t0.elapsed_time(t1),
                      "list_min": int(sizes.min()), "list_max": int(sizes.max())}), flush=True)
    ld = (args.dim + 7) // 8 * 8
    if args.selfcheck:
        _, q = bench.make_queries(8, args.dim, dev, seed=3)
        es, ei = idx.search(q, args.k)
        fs, fi = iv.search(q, args.k, args.nlist)                     # every list probed == the exact search
        ps, pi = iv.search(q, args.k, args.nprobe)
        torch.cuda.synchronize()
        a = torch.from_numpy(iv.assignments().astype("int64")).to(dev)
        lists, _ = iv.coarse_host(q.cpu().numpy(), args.nprobe)
        lists = torch.from_numpy(lists.astype("int64")).to(dev)
        ok = pi >= 0
        member = (a[pi.clamp(min=0)][:, :, None] == lists[:, None, :]).any(dim=2) | ~ok
        subset_best = bool((ps[:, 0] <= es[:, 0] + 1e-6).all())
        print(json.dumps({"tag": args.tag, "what": "selfcheck", "all_lists_ids_equal_exact": float((fi == ei).float().mean()),
                          "all_lists_max_score_diff": float((fs - es).abs().max()),
                          "probed_rows_in_probed_lists": float(member.float().mean()), "probe_best_le_exact_best": subset_best,
                          "scores_descending": bool((ps[:, 1:] <= ps[:, :-1]).all())}), flush=True)
    if args.selfcheck:                                   # grid granularity of the list scan at the reference's batch 1
        _, q1 = bench.make_queries(1, args.dim, dev, seed=1)
        for per_sm in (2, 4, 8, 16):
            os.environ["TS_IVF_CTAS_PER_SM"] = str(per_sm)
            fn = lambda: iv.search(q1, args.k, args.nprobe)          # noqa: E731
            bench.timed(fn, 1, 2, dev, False)
            idx.set_profiling(True)
            ms = bench.timed(fn, args.steps, 0, dev, False)
            kms, _n = idx.scan_time_ms()
            idx.set_profiling(False)
            print(json.dumps({"tag": args.tag, "what": "grid sweep B=1", "ctas_per_sm": per_sm, "step_ms": ms / args.steps,
                              "scan_ms": kms}), flush=True)
        os.environ.pop("TS_IVF_CTAS_PER_SM", None)
        _, q32 = bench.make_queries(32, args.dim, dev, seed=32)       # batches: pairs ordered by list (L2 reuse) vs grid order
        for noorder in ("0", "1"):
            os.environ["TS_IVF_NOORDER"] = noorder
            fn = lambda: iv.search(q32, args.k, args.nprobe)           # noqa: E731
            bench.timed(fn, 1, 2, dev, False)
            idx.set_profiling(True)
            ms = bench.timed(fn, args.steps, 0, dev, False)
            kms, _n = idx.scan_time_ms()
            idx.set_profiling(False)
            print(json.dumps({"tag": args.tag, "what": "batch order B=32", "pairs_ordered_by_list": noorder == "0",
                              "step_ms": ms / args.steps, "scan_ms": kms}), flush=True)
        os.environ.pop("TS_IVF_NOORDER", None)
    for B in [int(b) for b in args.batches.split(",")]:
        qh, q = bench.make_queries(B, args.dim, dev, seed=B)
        lists, _ = iv.coarse_host(qh.numpy(), args.nprobe)
        probed_rows = int(sizes[lists].sum())
        for name, fn in (("ivf", lambda: iv.search(q, args.k, args.nprobe)), ("exact", lambda: idx.search(q, args.k))):
            bench.timed(fn, 1, 2, dev, False)
            idx.set_profiling(True)
            ms = bench.timed(fn, args.steps, 0, dev, False)
            kms, n = idx.scan_time_ms()
            idx.set_profiling(False)
            kms = max(kms, 1e-9)
            rows_read = probed_rows if name == "ivf" else args.rows
            gb = rows_read * ld * 2 / (kms / 1e3) / 1e9
            print(json.dumps({"tag": args.tag, "what": name, "rows": args.rows, "dim": args.dim, "B": B, "nlist": args.nlist,
                              "nprobe": args.nprobe, "step_ms": ms / args.steps, "scan_ms": kms, "qps": B * args.steps / (ms / 1e3),
                              "rows_read_per_step": rows_read, "GBps": gb, "hbm_frac": gb / pk["hbm_gbs"]}), flush=True)


if __name__ == "__main__":
    main()
