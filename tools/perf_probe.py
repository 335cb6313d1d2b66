#!/usr/bin/env python
"""Sweep of scan-kernel time vs batch size / path on one GPU (development aid;
prints one line per configuration, inputs resident in HBM, CUDA-event timed)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from tristage_rag_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--batches", default="1,2,4,8,32,128,256,1024")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--paths", default="stream,umma")
    ap.add_argument("--tag", default="")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16", "fp32"],
                    help="corpus storage dtype (fp32 = the reference's; its umma path reads tf32)")
    ap.add_argument("--selfcheck", action="store_true",
                    help="device-vs-device consistency of the two scan kernels on the same index (4 queries)")
    args = ap.parse_args()
    bench.arm_watchdog(300)
    dev = torch.device("cuda", 0)
    pk = bench.peaks()
    idx = _lib.Index(args.dim, args.dtype, "ip", 0, reserve_rows=args.rows)
    bench.build_shard(idx, 0, args.rows, args.dim, dev, 1234, dtype=args.dtype)
    esz = 4 if args.dtype == "fp32" else 2
    ld = (args.dim + 3) // 4 * 4 if args.dtype == "fp32" else (args.dim + 7) // 8 * 8
    for path in args.paths.split(","):          # all of one path first: a fault in an unvalidated path cannot hide the other's lines
        for B in [int(b) for b in args.batches.split(",")]:
            if path == "stream" and B > 8:
                continue
            _, q = bench.make_queries(B, args.dim, dev, seed=B)
            fn = lambda: idx.search(q, args.k, path=path)  # noqa: E731
            bench.timed(fn, 1, 2, dev, False)
            idx.set_profiling(True)
            ms = bench.timed(fn, args.steps, 0, dev, False)
            kms, n = idx.scan_time_ms()
            idx.set_profiling(False)
            kms = max(kms, 1e-9)
            n_scans = max(n, 1) / args.steps
            gb = args.rows * ld * esz * n_scans / (kms * n_scans / 1e3) / 1e9
            tf = 2.0 * B * args.rows * ld / (kms * n_scans / 1e3) / 1e12
            print(json.dumps({"tag": args.tag, "rows": args.rows, "dim": args.dim, "dtype": args.dtype, "B": B, "path": path, "step_ms": ms / args.steps,
                              "scan_ms_per_launch": kms, "scans_per_step": n_scans,
                              "qps": B * args.steps / (ms / 1e3), "corpus_GBps_per_scan": gb,
                              "hbm_frac": gb / pk["hbm_gbs"], "TFLOPs": tf,
                              "tensor_frac_sustained": tf / pk["bf16_tflops_sustained"]}), flush=True)

    if args.selfcheck:
        _, q = bench.make_queries(4, args.dim, dev, seed=7)
        us, ui = idx.search(q, args.k, path="umma")
        ss, si = idx.search(q, args.k, path="stream")
        torch.cuda.synchronize()
        overlap = sum(len(set(ui[b].tolist()) & set(si[b].tolist())) for b in range(4)) / (4.0 * args.k)
        print(json.dumps({"tag": args.tag, "what": "selfcheck", "dtype": args.dtype, "topk_overlap_umma_vs_stream": overlap,
                          "max_abs_score_diff_rankwise": float((us - ss).abs().max()),
                          "best_id_equal": bool((ui[:, 0] == si[:, 0]).all())}), flush=True)


if __name__ == "__main__":
    main()
