#!/usr/bin/env python
"""A/B of an opt-in kernel variant against the hardware-validated default IN ONE PROCESS (the switches are read
at every launch): same inputs, results compared bit for bit on the device, CUDA-event timing of both.  One JSON
line per comparison.  Used by `bench.py` (child process, after the headline) and by the round-2 GPU scripts.

  python tools/variant_ab.py --what s1      # TS_FUSE (one cooperative launch) and TS_SELECT_V1 (first select kernel)
  python tools/variant_ab.py --what pair    # TS_PAIR (cta_group::2 CTA pairs, B >= 129)
  python tools/variant_ab.py --what s2      # TS_S2_V2 (second Stage-2 epilogue), TS_S2_EPI2 (two epilogue warpgroups), both
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from tristage_rag_b200 import _lib  # noqa: E402


def ab(name, switch, fn, steps, dev, extra=None):
    """fn() -> tuple of cuda tensors.  Default first (validated), then with `switch`=1 (several switches: "A+B")."""
    switches = switch.split("+")
    for sw in switches:
        os.environ[sw] = "0"
    ref = fn()
    torch.cuda.synchronize()
    t0 = bench.timed(fn, steps, 3, dev, False) / steps
    rec = {"what": name, "switch": switch, "default_ms": t0}
    try:
        for sw in switches:
            os.environ[sw] = "1"
        got = fn()
        torch.cuda.synchronize()
        rec["bit_equal"] = bool(all(torch.equal(a, b) for a, b in zip(got, ref)))
        if not rec["bit_equal"]:
            rec["max_abs_diff"] = max(float((a.double() - b.double()).abs().max()) for a, b in zip(got, ref))
        rec["variant_ms"] = bench.timed(fn, steps, 3, dev, False) / steps
        rec["speedup"] = t0 / rec["variant_ms"]
    except Exception as e:                                   # noqa: BLE001 -- recorded, the next comparison still runs
        rec["error"] = f"{type(e).__name__}: {e}"[:300]
    finally:
        for sw in switches:
            os.environ[sw] = "0"
    if extra:
        rec.update(extra)
    print(json.dumps(rec), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="s1", choices=["s1", "s2", "pair"])
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--rows", type=int, default=0, help="corpus rows (default: 1.25 M for s1, 4 M for pair)")
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--ndocs", type=int, default=200_000, help="token-store documents (s2)")
    ap.add_argument("--cands", type=int, default=1000, help="candidates per query (s2)")
    ap.add_argument("--queries", type=int, default=64, help="query batch (s2)")
    ap.add_argument("--pair-batches", default="256,1024", help="query batches of the pair comparison (>= 129)")
    args = ap.parse_args()
    bench.arm_watchdog(300)
    dev = torch.device("cuda", 0)
    if args.what == "s1":
        rows, dim, k = args.rows or 1_250_000, args.dim, 100   # the per-GPU shard of the 8-GPU headline run: fixed costs show
        idx = _lib.Index(dim, "bf16", "ip", 0, reserve_rows=rows)
        bench.build_shard(idx, 0, rows, dim, dev, 1234)
        for B in (1, 32):
            _, q = bench.make_queries(B, dim, dev, seed=B)
            fn = lambda: idx.search(q, k)            # noqa: E731
            ab(f"stage1 search B={B}: pre-pass + scan in one cooperative launch", "TS_FUSE", fn, args.steps, dev,
               {"rows": rows, "dim": dim})
            ab(f"stage1 search B={B}: FIRST select kernel (the default is its rewrite)", "TS_SELECT_V1", fn, args.steps, dev,
               {"rows": rows, "dim": dim})
    elif args.what == "pair":
        rows, dim, k = args.rows or 4_000_000, args.dim, 100   # the tensor-bound regime: two query tiles of one slice on a CTA pair
        idx = _lib.Index(dim, "bf16", "ip", 0, reserve_rows=rows)
        bench.build_shard(idx, 0, rows, dim, dev, 1234)
        for B in [int(b) for b in args.pair_batches.split(",")]:
            _, q = bench.make_queries(B, dim, dev, seed=B)
            fn = lambda: idx.search(q, k)            # noqa: E731
            ab(f"stage1 search B={B}: cta_group::2 CTA pairs", "TS_PAIR", fn, min(args.steps, 10), dev,
               {"rows": rows, "dim": dim, "TFLOP_per_search": 2.0 * B * rows * dim / 1e12})
    else:
        ndocs, dim, B, C, Lq = args.ndocs, 128, args.queries, args.cands, 32
        g = torch.Generator(device=dev).manual_seed(77)
        rng = np.random.default_rng(77)
        lens = rng.integers(16, 181, size=ndocs).astype(np.int32)
        st = _lib.TokStore(dim, "bf16", 0, reserve_docs=ndocs, reserve_tokens=int(lens.sum()))
        for s in range(0, ndocs, 100_000):
            ln = lens[s:s + 100_000]
            t = torch.nn.functional.normalize(torch.randn((int(ln.sum()), dim), generator=g, device=dev), dim=-1)
            st.add(t.to(torch.bfloat16), ln, normalize=False)
        q = torch.nn.functional.normalize(torch.randn((B, Lq, dim), generator=g, device=dev), dim=-1).to(torch.bfloat16)
        cand = torch.stack([torch.randperm(ndocs, generator=g, device=dev)[:C] if C <= ndocs
                            else torch.randint(0, ndocs, (C,), generator=g, device=dev) for _ in range(B)])
        for mode, name in ((_lib.TS_S2_MAXSIM, "maxsim"), (_lib.TS_S2_COLBERT, "colbert")):
            fn = lambda: (st.maxsim(q, cand, mode=mode, normalize_q=False),)      # noqa: E731,B023
            for sw, what in (("TS_S2_V2", "second epilogue"), ("TS_S2_EPI2", "two epilogue warpgroups"),
                             ("TS_S2_V2+TS_S2_EPI2", "second epilogue in two warpgroups")):
                if mode == _lib.TS_S2_COLBERT and sw != "TS_S2_V2+TS_S2_EPI2":
                    continue                          # the softmax-sum mode only differs in the finalize step
                ab(f"stage2 {name} {B} q x {C} cand (config #4 shapes): {what}", sw, fn, args.steps, dev,
                   {"ndocs": ndocs, "dim": dim})


if __name__ == "__main__":
    main()
