#!/usr/bin/env python
"""Stage-2 MaxSim kernel timing on BASELINE config #4 shapes (development aid)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from tristage_rag_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ndocs", type=int, default=200_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--C", type=int, default=1000)
    ap.add_argument("--Lq", type=int, default=32)
    ap.add_argument("--lo", type=int, default=16)
    ap.add_argument("--hi", type=int, default=180)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    pk = bench.peaks()
    g = torch.Generator(device=dev).manual_seed(77)
    rng = np.random.default_rng(77)
    lens = rng.integers(args.lo, args.hi + 1, size=args.ndocs).astype(np.int32)
    st = _lib.TokStore(args.dim, "bf16", 0, reserve_docs=args.ndocs, reserve_tokens=int(lens.sum()))
    for s in range(0, args.ndocs, 100_000):
        ln = lens[s:s + 100_000]
        t = torch.nn.functional.normalize(torch.randn((int(ln.sum()), args.dim), generator=g, device=dev), dim=-1)
        st.add(t.to(torch.bfloat16), ln, normalize=False)
    q = torch.nn.functional.normalize(torch.randn((args.B, args.Lq, args.dim), generator=g, device=dev), dim=-1).to(torch.bfloat16)
    cand = torch.stack([torch.randperm(args.ndocs, generator=g, device=dev)[:args.C] for _ in range(args.B)])
    fn = lambda: st.maxsim(q, cand, normalize_q=False)  # noqa: E731
    bench.timed(fn, 1, 3, dev, False)
    st.set_profiling(True)
    ms = bench.timed(fn, args.steps, 0, dev, False)
    kms, _ = st.scan_time_ms()
    nbytes = float(lens[cand.cpu().numpy()].astype(np.int64).sum()) * args.dim * 2
    print(json.dumps({"tag": args.tag, "ndocs": args.ndocs, "dim": args.dim, "B": args.B, "C": args.C, "Lq": args.Lq,
                      "Ld": [args.lo, args.hi], "kernel_ms": kms, "step_ms": ms / args.steps,
                      "cand_per_s": args.B * args.C / (kms / 1e3), "GBps": nbytes / (kms / 1e3) / 1e9,
                      "hbm_frac": nbytes / (kms / 1e3) / 1e9 / pk["hbm_gbs"]}), flush=True)


if __name__ == "__main__":
    main()
