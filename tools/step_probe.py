#!/usr/bin/env python
"""What one rank of an N-GPU job does per Stage-1 step, measured on ONE GPU: a shard of rows/N rows, B queries,
the local part of the step (query prep + scan + select) timed eagerly, replayed from a CUDA graph, and through
the C-ABI host call (H2D + search + D2H + sync), for the scan-launch variants (TS_FUSE on/off).  The exchange
(all-gather / peer push) is not part of this probe -- it needs the other GPUs.  Development aid."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from tristage_rag_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_250_000)
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--batches", default="1,32")
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--variants", default="TS_FUSE=1,TS_FUSE=0")
    args = ap.parse_args()
    bench.arm_watchdog(300)
    dev = torch.device("cuda", 0)
    pk = bench.peaks()
    idx = _lib.Index(args.dim, "bf16", "ip", 0, reserve_rows=args.rows)
    bench.build_shard(idx, 0, args.rows, args.dim, dev, 1234)
    for B in [int(b) for b in args.batches.split(",")]:
        q_pin, q = bench.make_queries(B, args.dim, dev, seed=B)
        blob = torch.empty(_lib.packed_layout(B, args.k)[1], dtype=torch.uint8, device=dev)
        hs = torch.empty((B, args.k), dtype=torch.float32).pin_memory()
        hi = torch.empty((B, args.k), dtype=torch.int64).pin_memory()
        for var in args.variants.split(","):
            name, val = var.split("=")
            os.environ[name] = val
            fn = lambda: idx.search_packed(q, args.k, blob)      # noqa: E731
            bench.timed(fn, 3, 3, dev, False)
            idx.set_profiling(True)
            ms = bench.timed(fn, args.steps, 0, dev, False)
            scan_ms, _ = idx.scan_time_ms()
            idx.set_profiling(False)
            clean = bench.timed(fn, args.steps, 0, dev, False)      # the same steps without the kernel events between the launches
            rec = {"rows": args.rows, "B": B, "variant": var, "eager_us": clean / args.steps * 1e3, "eager_with_events_us": ms / args.steps * 1e3,
                   "scan_us": scan_ms * 1e3}
            try:
                graph, _ = bench.graph_of(fn, dev)
                rec["graph_us"] = bench.timed(lambda: graph.replay(), args.steps, 5, dev, False) / args.steps * 1e3
            except Exception as e:                                # noqa: BLE001
                rec["graph_error"] = f"{type(e).__name__}: {e}"[:120]
            qn, out = q_pin.numpy(), (hs.numpy(), hi.numpy())
            host = lambda: idx.search_host(qn, args.k, out=out)  # noqa: E731
            rec["host_call_us"] = bench.timed_wall(host, args.steps, 5, dev, False) / args.steps * 1e3
            t0 = time.perf_counter()                              # CPU time to ENQUEUE a step (no sync): launch overhead
            for _ in range(args.steps):
                fn()
            rec["enqueue_us"] = (time.perf_counter() - t0) / args.steps * 1e3
            torch.cuda.synchronize()
            ld = (args.dim + 7) // 8 * 8
            rec["ideal_scan_us"] = args.rows * ld * 2 / (pk["hbm_gbs"] * 1e9) * 1e6
            print(json.dumps({k: (round(v, 1) if isinstance(v, float) else v) for k, v in rec.items()}), flush=True)
            del os.environ[name]


if __name__ == "__main__":
    main()
