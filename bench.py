#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 candidate-scoring hot path.

Metric (BASELINE.json): Stage-1 queries/s, exact top-100 over a 10M x 1024 bf16 corpus (configs[2]; fits one
B200: 20.5 GB) with the HBM-roofline fraction of the scan kernel; Stage-2 candidates/s (configs[3]), the other
batch sizes and the configs[4] chain ride along in `roofline.also` (compact: the driver keeps `roofline`).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...     # the CPU arm (oracle port, host cores)

A step = one search call of B queries over the whole (row-sharded) corpus: query prep + scan (fused top-k) +
select (+ exchange of the [B, k] lists + merge for N > 1).  Inputs are synthetic (seeded randn, row-normalised
with the reference formula); 100 rows are planted for each of the first two queries and must come back, and
rank 0 checks three queries of the TIMED index against the CPU oracle over its first 1 M stored rows
(`parity`).  Stage 2: 64 queries x 1000 candidates over a 1 M-doc token store sharded like the corpus, local
MaxSim + all-reduce, two queries checked against the C oracle.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "stage1_queries_per_s_top100_10Mx1024_bf16"
PLANT_QUERIES, PLANT_ROWS = 2, 100


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="tristage", choices=["tristage", "reference"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--path", default="auto")
    ap.add_argument("--no-extra", action="store_true", help="skip B=1 / B=1024 / Stage-2 / C5 side measurements")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the timed workload")
    ap.add_argument("--no-graph", action="store_true", help="N > 1: do not replay the step from a CUDA graph")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while work runs (the timed region of the headline is 60 - 170 ms)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,utilization.gpu,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, uuid: str):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", uuid, f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "50"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        rows = []
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) >= 9:
                try:
                    rows.append((float(c[0]), float(c[1]), float(c[3]), c[5], c[6], c[7], c[8]))
                except ValueError:
                    continue
        os.unlink(self.f.name)
        if not rows:
            return out
        busy = [r for r in rows if r[2] >= 50.0] or rows
        out["sm_mhz"] = statistics.median(r[0] for r in busy)
        out["sm_max_mhz"] = max(r[1] for r in rows)
        out["samples"] = len(busy)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for i, n in enumerate(names):
            if any(r[3 + i].lower().startswith("active") for r in busy):
                out["reasons"].append(n)
        return out


def plant_positions(N, seed=99):
    """global row numbers of the planted rows: [PLANT_QUERIES, PLANT_ROWS], the same on every rank"""
    import numpy as np

    return np.random.default_rng(seed).choice(N, size=PLANT_QUERIES * PLANT_ROWS, replace=False).reshape(PLANT_QUERIES, PLANT_ROWS)


def build_shard(idx, lo, hi, dim, dev, seed, chunk=500_000, dtype="bf16", plant=None):
    """rows [lo, hi) of the synthetic corpus.  plant = (queries [P, dim] on dev, positions [P, R] global): those
    rows become normalize(q + 0.7 * noise / sqrt(dim)) -- known answers for the parity check of the timed index."""
    import torch

    g = torch.Generator(device=dev).manual_seed(seed)
    tdt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[dtype]
    pos_t = torch.from_numpy(plant[1]).to(dev) if plant is not None else None
    for s in range(lo, hi, chunk):
        n = min(chunk, hi - s)
        x = torch.randn((n, dim), generator=g, device=dev, dtype=torch.float32)
        if plant is not None:
            for b in range(pos_t.shape[0]):
                sel = pos_t[b][(pos_t[b] >= s) & (pos_t[b] < s + n)] - s
                if len(sel):
                    noise = torch.randn((len(sel), dim), generator=g, device=dev) / dim ** 0.5
                    x[sel] = plant[0][b][None, :] + 0.7 * noise
        x /= x.norm(dim=1, keepdim=True) + 1e-8          # reference formula, stage1_retriever.py:287-288
        idx.add(x.to(tdt), normalize=False)
        del x


def make_queries(B, dim, dev, seed=4321):
    import torch

    g = torch.Generator(device="cpu").manual_seed(seed)
    q = torch.randn((B, dim), generator=g, dtype=torch.float32)
    q /= q.norm(dim=1, keepdim=True) + 1e-8
    return q.pin_memory(), q.to(dev)


def timed(fn, steps, warmup, dev, dist_on):
    """K steps bracketed by barrier + synchronize; CUDA events; max over ranks (ms)."""
    import torch
    import torch.distributed as dist

    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(dev)
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    if dist_on:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if dist_on:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def timed_wall(fn, steps, warmup, dev, dist_on):
    """end-to-end variant: fn synchronises itself (host result in hand); wall clock, max over ranks."""
    import torch
    import torch.distributed as dist

    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(dev)
    if dist_on:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize(dev)
    ms = (time.perf_counter() - t0) * 1e3
    if dist_on:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def graph_of(step, dev):
    """the step captured once in a CUDA graph (all launches incl. the NCCL exchange): replay has no per-launch host latency"""
    import torch

    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize(dev)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = step()
    return graph, out


def stage1_alg_bytes(n_rows, ld, B, k, L):
    return n_rows * ld * 2 + B * ld * 2 + L * B * k * 8


def workload_config(N, d, k, B, world):
    """The `config` both arms print: the workload only (what runs it is described under `roofline`)."""
    return {"workload": f"exact top-{k} over {N}x{d} bf16 corpus, query batch {B}", "rows": N, "dim": d, "k": k,
            "batch": B, "n_gpus": world}


def r3(x):
    return None if x is None else float(f"{x:.4g}")


def finish(code: int = 0):
    """Leave without running interpreter/NCCL teardown: destroying a process group (or CUDA
    graphs that captured NCCL work) can block for minutes after the result line is out."""
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(code)


_PENDING = {"line": None}      # the result line once the headline is measured: the watchdog prints it rather than nothing


def arm_watchdog(seconds: int):
    """Hard wall-clock limit for the whole run; a wedged collective must not hold the box."""
    import threading

    def _kill():
        if _PENDING["line"] is not None:
            print(f"[bench] watchdog after {seconds}s: printing the measured line without the side measurements", file=sys.stderr, flush=True)
            print(json.dumps(_PENDING["line"]), flush=True)
            os._exit(0)
        print(f"[bench] watchdog: no result after {seconds}s, aborting", file=sys.stderr, flush=True)
        os._exit(3)

    t = threading.Timer(seconds, _kill)
    t.daemon = True
    t.start()


def run_reference(args):
    """CPU arm: the reference's Stage-1 arithmetic (oracle port; FAISS absent) on ALL host cores.  Every step scans
    a bounded sample (500 k of the N rows); ms_per_step is the measured time of that sample step, `value` the rate
    scaled linearly to the full corpus (the scan is O(N))."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)          # torchrun exports OMP_NUM_THREADS=1: the CPU arm uses every core anyway
    from oracle import cpu_baseline

    info = cpu_baseline.host_info()
    sample_rows = min(500_000, args.rows)
    g = torch.Generator().manual_seed(1234)
    X = torch.randn((sample_rows, args.dim), generator=g, dtype=torch.float32)
    X /= X.norm(dim=1, keepdim=True) + 1e-8
    Q = torch.randn((args.batch, args.dim), generator=g, dtype=torch.float32)
    Q /= Q.norm(dim=1, keepdim=True) + 1e-8
    kk = min(args.k, sample_rows)

    def step():
        torch.topk(Q @ X.T, kk, dim=1)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    sample_s = (time.perf_counter() - t0) / max(1, args.steps)
    value = args.batch / (sample_s * (args.rows / sample_rows))
    sample = (f"restated IndexFlatIP (fp32 torch/MKL Q@X.T + topk), each step = {sample_rows} of {args.rows} rows x {args.dim}, "
              f"B={args.batch}, k={args.k}, {threads} threads; value scaled linearly in N; FAISS not installed")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sample_s * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.rows, args.dim, args.k, args.batch, args.gpus),
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample, "ms_per_step_full_corpus": sample_s * 1e3 * args.rows / sample_rows,
                         "host": info},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- parity of the timed workload ---
def check_stage1(idx, lo, n_local, q_dev, D, I, pos, k, rank):
    """D, I = the merged result of the TIMED index for the whole batch.  (1) the planted rows of queries 0/1 come
    back exactly; (2) rank 0: queries 0, 2, 3 against the oracle's exact fp32 scores over its first <= 1 M stored
    rows -- every returned id in that range carries the oracle's score (1e-3 relative), and no row of the range
    that the oracle ranks above our k-th score (beyond the 1e-3 tie band) is missing."""
    import numpy as np

    from oracle import flat_ip

    Dh, Ih = D.cpu().numpy(), I.cpu().numpy()
    out = {"planted_ok": all(set(Ih[b].tolist()) == set(pos[b].tolist()) for b in range(PLANT_QUERIES))}
    if rank != 0:
        return out
    n_chk = int(min(1_000_000, n_local))
    X = idx.get_rows(0, n_chk)                          # the stored (bf16-rounded) rows as fp32
    qs = [0, 2, 3] if Dh.shape[0] >= 4 else [0]
    Qr = flat_ip.round_to(q_dev[qs].float().cpu().numpy(), "bf16")
    S = Qr @ X.T
    rD, rI = flat_ip.topk_desc(S, min(k, n_chk))
    ok, worst = True, 0.0
    for j, b in enumerate(qs):
        inside = (Ih[b] >= lo) & (Ih[b] < lo + n_chk)
        if inside.any():
            ref = S[j, Ih[b][inside] - lo].astype(np.float64)
            err = np.abs(Dh[b][inside] - ref) / np.maximum(np.abs(ref), 1e-30)
            worst = max(worst, float(err.max()))
            ok &= bool((err <= 1e-3).all())
        kth = float(Dh[b, -1])
        band = 1e-3 * abs(kth)
        must = rI[j][rD[j] > kth + band] + lo           # oracle ids clearly above our k-th score
        ok &= bool(np.isin(must, Ih[b]).all())
        ok &= bool((np.diff(Dh[b]) <= 0).all())
    out.update(oracle_ok=ok, oracle_rows=n_chk, oracle_queries=len(qs), max_rel_err=r3(worst))
    return out


# --------------------------------------------------------------------------- side measurements -------------
def side_stage1(sharded, idx, d, k, dev, pk, ld, n_local, world, dist_on):
    """B = 1 (bandwidth-bound) and B = 1024 (tensor-bound) through the same sharded search."""
    out = {}
    for B, steps in ((1, 20), (1024, 5)):
        _, qd = make_queries(B, d, dev, seed=99 + B)
        fn = lambda: sharded.search(qd, k)                   # noqa: E731
        timed(fn, 1, 2, dev, dist_on)
        idx.set_profiling(True)
        ms = timed(fn, steps, 0, dev, dist_on)
        scan_ms, _ = idx.scan_time_ms()
        idx.set_profiling(False)
        rec = {"qps": r3(B * steps / (ms / 1e3)), "ms": r3(ms / steps), "scan_ms": r3(scan_ms)}
        if B == 1:
            gb = stage1_alg_bytes(n_local, ld, B, k, 148) / (scan_ms / 1e3) / 1e9
            rec.update(gbs=r3(gb), frac=r3(gb / pk["hbm_gbs"]))
        else:
            tf = 2.0 * B * n_local * ld / (scan_ms / 1e3) / 1e12
            rec.update(tflops=r3(tf), frac=r3(tf / pk["bf16_tflops_sustained"]), frac_burst=r3(tf / pk["bf16_tflops"]))
        out[f"s1_b{B}"] = rec
    return out


def build_tokstore(n_total, lo, hi, dim, ld_lo, ld_hi, dev, seed, keep_first=0):
    """docs [lo, hi) of a synthetic token store (Ld ~ U[ld_lo, ld_hi], unit tokens, bf16).  Returns the store, the
    GLOBAL length table and (optionally) the bf16 tokens of this rank's first `keep_first` docs for the parity check."""
    import numpy as np
    import torch

    from tristage_rag_b200 import _lib

    lens_all = np.random.default_rng(seed).integers(ld_lo, ld_hi + 1, size=n_total).astype(np.int32)
    lens = lens_all[lo:hi]
    st = _lib.TokStore(dim, "bf16", dev.index, reserve_docs=hi - lo, reserve_tokens=int(lens.sum()))
    g = torch.Generator(device=dev).manual_seed(seed + 1 + lo)
    kept = None
    chunk = 100_000
    for s in range(0, hi - lo, chunk):
        ln = lens[s:s + chunk]
        t = torch.nn.functional.normalize(torch.randn((int(ln.sum()), dim), generator=g, device=dev), dim=-1).to(torch.bfloat16)
        st.add(t, ln, normalize=False)
        if s == 0 and keep_first:
            n0 = min(keep_first, len(ln))
            kept = (t[: int(ln[:n0].sum())].clone(), ln[:n0].copy())
        del t
    return st, lens_all, kept


def side_stage2(dev, pk, rank, world, dist_on, parity=True):
    """BASELINE configs[3]: 64 queries x 1000 candidates (Lq 32, Ld ~ U[16,180], dim 128) over a 1 M-doc token store
    sharded by doc id like the corpus; every rank scores the candidates it owns, one all-reduce sums the [64, 1000]
    matrix (src/stage2_rescorer.py:268-291 is the loop this replaces)."""
    import numpy as np
    import torch

    from tristage_rag_b200.dist import ShardedTokStore, owner_of, shard_range

    ndocs, dim, Bq, C, Lq = 1_000_000, 128, 64, 1000, 32
    lo, hi = shard_range(ndocs, rank, world)
    st, lens_all, kept = build_tokstore(ndocs, lo, hi, dim, 16, 180, dev, seed=77, keep_first=100_000 if parity else 0)
    sst = ShardedTokStore(st, ndocs)
    rng = np.random.default_rng(78)
    cand_h = np.stack([rng.choice(ndocs, size=C, replace=False) for _ in range(Bq)]).astype(np.int64)
    cand = torch.from_numpy(cand_h).to(dev)
    g = torch.Generator(device="cpu").manual_seed(79)
    qt = torch.nn.functional.normalize(torch.randn((Bq, Lq, dim), generator=g), dim=-1).to(torch.bfloat16).to(dev)
    fn_local = lambda: st.maxsim(qt, cand, normalize_q=False)      # noqa: E731
    fn = lambda: sst.maxsim(qt, cand, normalize_q=False)           # noqa: E731
    timed(fn, 1, 3, dev, dist_on)
    steps = 20
    st.set_profiling(True)
    ms = timed(fn, steps, 0, dev, dist_on)
    kms, _ = st.scan_time_ms()
    st.set_profiling(False)
    ms_local = timed(fn_local, steps, 1, dev, dist_on) if world > 1 else ms
    if dist_on:
        import torch.distributed as dist

        t = torch.tensor([kms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        kms = float(t.item())
    tok_bytes = float(lens_all[cand_h].astype(np.int64).sum()) * dim * 2       # algorithmic: real tokens only
    gb = tok_bytes / world / (kms / 1e3) / 1e9                                  # per GPU
    own = np.bincount(owner_of(torch.from_numpy(cand_h.ravel()), ndocs, world).numpy(), minlength=world)
    rec = {"cand_per_s": r3(Bq * C * steps / (ms / 1e3)), "ms": r3(ms / steps), "kernel_ms": r3(kms),
           "gbs_per_gpu": r3(gb), "frac": r3(gb / pk["hbm_gbs"]), "exchange_ms": r3(max(0.0, (ms - ms_local) / steps)),
           "cand_max_over_mean": r3(float(own.max()) / max(1.0, float(own.mean())))}
    scores_t = fn() if parity else None                 # every rank: the exchange is a collective step
    if parity and rank == 0 and kept is not None:
        from oracle import c_oracle

        scores = scores_t.cpu().numpy()
        tok0, ln0 = kept
        off0 = np.concatenate([[0], np.cumsum(ln0.astype(np.int64))])
        tok0 = tok0.float().cpu().numpy()
        ok, n_chk, worst = True, 0, 0.0
        for b in (0, Bq - 1):
            js = np.nonzero((cand_h[b] >= lo) & (cand_h[b] < lo + len(ln0)))[0]
            if not len(js):
                continue
            ids = cand_h[b][js] - lo
            rows = np.concatenate([tok0[off0[i]:off0[i + 1]] for i in ids])
            o = np.concatenate([[0], np.cumsum(ln0[ids].astype(np.int64))])
            ref = c_oracle.maxsim_batch(qt[b].float().cpu().numpy(), rows, o, mode=0)
            err = np.abs(scores[b][js] - ref)
            worst = max(worst, float(err.max()))
            ok &= bool(np.allclose(scores[b][js], ref, rtol=1e-3, atol=2e-4))
            n_chk += len(js)
        rec.update(oracle_ok=ok, oracle_pairs=n_chk, max_abs_err=r3(worst))
    if dist_on:
        import torch.distributed as dist

        dist.barrier()          # rank 0 checked alone: re-align before the next exchange step (bounded waits on the peer plane)
    del st, sst
    return rec


def side_c5(dev, rank, world, dist_on):
    """BASELINE configs[4]: Stage 1 (k = 500, d = 768) -> Stage 2 over those 500 candidates (Ld ~ U[16,192], dim 128)
    -> stable top-100, 64 queries per step (src/retrieval_pipeline.py:358,375 with benchmark/config.yaml:35,45,47).
    5 M docs need 8 GPUs for the token store (~135 GB): the corpus is 625 k docs per GPU, i.e. the full size at N = 8."""
    import torch

    from tristage_rag_b200 import _lib
    from tristage_rag_b200.dist import ShardedIndex, ShardedTokStore, shard_range

    docs, d, k1, k2, B, dim, Lq = 625_000 * world, 768, 500, 100, 64, 128, 32
    lo, hi = shard_range(docs, rank, world)
    idx = _lib.Index(d, "bf16", "ip", dev.index, reserve_rows=hi - lo)
    build_shard(idx, lo, hi, d, dev, seed=4000 + rank)
    st, _, _ = build_tokstore(docs, lo, hi, dim, 16, 192, dev, seed=177)
    sidx, sst = ShardedIndex(idx, docs), ShardedTokStore(st, docs)
    _, q = make_queries(B, d, dev, seed=5)
    g = torch.Generator(device="cpu").manual_seed(6)
    qt = torch.nn.functional.normalize(torch.randn((B, Lq, dim), generator=g), dim=-1).to(torch.bfloat16).to(dev)

    def step():
        s1, i1 = sidx.search(q, k1)                       # exact top-500, merged across ranks
        s2 = sst.maxsim(qt, i1, normalize_q=False)        # owners score, all-reduce sums
        return _lib.rank_desc(s2, k2, device=dev.index)   # stable top-100 per query

    steps = 10
    ms = timed(step, steps, 3, dev, dist_on)
    top_s, top_p = step()
    torch.cuda.synchronize(dev)
    sane = bool((top_s[:, :-1] >= top_s[:, 1:]).all().item()) and bool((top_p >= 0).all().item())
    return {"docs": docs, "qps": r3(B * steps / (ms / 1e3)), "ms": r3(ms / steps), "batch": B, "sorted_ok": sane}


# --------------------------------------------------------------------------- main ---------------------------
def main():
    args = parse()
    arm_watchdog(900)
    if args.impl == "reference":
        run_reference(args)
        finish(0)

    import torch
    import torch.distributed as dist

    from tristage_rag_b200 import _lib
    from tristage_rag_b200.dist import ShardedIndex, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist_on = world > 1
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if dist_on:
        dist.init_process_group("nccl", device_id=dev)
    assert _lib.lib().ts_device_count() >= 1, "no sm_100 device: libtristage has no CPU fallback"

    N, d, B, k = args.rows, args.dim, args.batch, args.k
    lo, hi = shard_range(N, rank, world)
    q_pin, q_dev = make_queries(B, d, dev)
    pos = plant_positions(N)
    idx = _lib.Index(d, "bf16", "ip", local, reserve_rows=hi - lo)
    build_shard(idx, lo, hi, d, dev, seed=1234 + rank, plant=(q_dev[:PLANT_QUERIES], pos) if B >= PLANT_QUERIES else None)
    sharded = ShardedIndex(idx, N)
    torch.cuda.synchronize(dev)

    uuid = str(torch.cuda.get_device_properties(dev).uuid)
    sampler = ClockSampler(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)

    # ---- device-resident timing (value) --------------------------------------
    path = args.path
    p2p = bool(getattr(sharded, "_p2p", False))
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    # the whole exchange is ONE kernel (select + push + wait + merge) when its B CTAs are co-resident and G * k keys fit
    xfuse = p2p and B <= sm_count and k <= 128 and world * k <= 2048 and os.environ.get("TS_XFUSE", "1") != "0"
    step = lambda: sharded.search(q_dev, k, path=path)     # noqa: E731
    timed(step, 2, args.warmup, dev, dist_on)               # warm-up incl. scratch allocation
    # scan-kernel duration: CUDA events around the kernel on its own stream, eager launches, INSIDE the timed region (one
    # pass: a second, event-free pass of the same K steps measured 7 % slower at N = 1 -- after 60 ms more of streaming the
    # GPU sits in sw_power_cap -- while the two event records cost a few us per step)
    idx.set_profiling(True)
    l0 = idx.launches
    ms_eager = timed(step, args.steps, 0, dev, dist_on)
    scan_ms, scan_n = idx.scan_time_ms()
    idx.set_profiling(False)
    # + the post-all-gather merge kernel (or push + wait-merge of the peer-memory exchange)
    launches = idx.launches - l0 + (args.steps * (0 if xfuse else 1) if world > 1 else 0)
    # N > 1: the step (query prep, scan, select, all-gather, merge) is short enough for launch latency to show:
    # replay it from a CUDA graph and keep whichever launch mode the box runs faster
    ms, launch_mode, ms_graph = ms_eager, "eager", None
    if dist_on and not args.no_graph and not p2p:           # the peer exchange takes its step number from the host
        try:
            graph, g_out = graph_of(step, dev)
            ms_graph = timed(lambda: graph.replay(), args.steps, args.warmup, dev, dist_on)
            ref_s, ref_i = step()
            torch.cuda.synchronize(dev)
            assert torch.equal(g_out[1], ref_i) and torch.equal(g_out[0], ref_s), "graph replay diverged from eager"
            if ms_graph < ms_eager:
                ms, launch_mode = ms_graph, "cuda_graph"
        except Exception as e:                               # noqa: BLE001
            print(f"[bench] CUDA-graph capture unavailable ({type(e).__name__}: {e}); eager timing kept", file=sys.stderr)
    value = B * args.steps / (ms / 1e3)
    res_s, res_i = step()
    torch.cuda.synchronize(dev)

    # ---- end-to-end timing (host buffers in, host results out) ---------------
    # inputs and results live in pinned host memory, allocated once like a serving process would
    q_stage = torch.empty_like(q_dev)
    h_scores = torch.empty((B, k), dtype=torch.float32).pin_memory()
    h_ids = torch.empty((B, k), dtype=torch.int64).pin_memory()
    q_np, out_np = q_pin.numpy(), (h_scores.numpy(), h_ids.numpy())

    def e2e_step():
        if world == 1:
            idx.search_host(q_np, k, path=path, out=out_np)  # the C-ABI host call: H2D, search, D2H, sync
        elif p2p:
            sharded.search_host(q_np, k, out=out_np, path=path)   # one C call: H2D, scan, select+push, wait+merge, D2H, sync
        else:
            q_stage.copy_(q_pin, non_blocking=True)
            s, i = sharded.search(q_stage, k, path=path)
            h_scores.copy_(s, non_blocking=True)
            h_ids.copy_(i, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()     # merged result in hand on the host
    e2e_ms = timed_wall(e2e_step, args.steps, min(args.warmup, 3), dev, dist_on)
    e2e_value = B * args.steps / (e2e_ms / 1e3)
    clocks = sampler.stop()

    pk = peaks()
    ld = ((d + 7) // 8) * 8
    n_local = hi - lo
    use_stream = (path == "stream")
    sm = torch.cuda.get_device_properties(dev).multi_processor_count
    L = min(2 * sm, max(1, n_local // 32)) if use_stream else min(sm // ((B + 127) // 128), (n_local + 255) // 256)
    alg_bytes = stage1_alg_bytes(n_local, ld, B, k, L)
    achieved = alg_bytes / (scan_ms / 1e3) / 1e9 if scan_ms > 0 else 0.0
    roof = {"bound": "hbm", "kernel": "s1_stream_kernel" if use_stream else "s1_umma_kernel",
            "achieved": r3(achieved), "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": r3(achieved / pk["hbm_gbs"]),
            "peak_source": pk["source"], "traffic": None, "algorithmic_bytes_per_launch": alg_bytes,
            "kernel_ms": r3(scan_ms), "kernel_launches_timed": scan_n, "frac_of_nominal_8TBs": r3(achieved / 8000.0),
            "launch": launch_mode, "ms_eager": r3(ms_eager / args.steps),
            "ms_graph": r3(ms_graph / args.steps) if ms_graph else None,
            "exchange": ("peer-memory: select + push + wait + merge in one kernel" if xfuse else "peer-memory: select + push kernel, wait + merge kernel" if p2p
                         else ("nccl all-gather + merge kernel" if world > 1 else "none"))}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof) and world == 1 and (N, d) == (10_000_000, 1024):
        with open(prof) as f:                      # NOT measured in this run: one ncu --set full capture of this workload
            tj = json.load(f)
        roof["traffic"] = tj.get(roof["kernel"])
        roof["traffic_source"] = tj.get("source", "static: ncu capture under profiles/")

    line = {
        "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {**workload_config(N, d, k, B, world), "l2": "inputs larger than L2 (shard >= 2.5 GB vs 126 MB), no flush"},
        "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": B * d * 4,
                "d2h_bytes_per_step": B * k * 12, "ms_per_step": r3(e2e_ms / args.steps)},
        "gpu_launches": int(launches), "roofline": roof, "clocks": clocks,
    }
    if world > 1:
        # which data plane carried the [B, k] lists (the driver's NCCL evidence covers the communicator only: with the
        # peer-memory exchange NCCL does the barriers, the max-over-ranks reduction and the set-up agreement, no data)
        line["comm"] = {"stage1_exchange": roof["exchange"], "nranks": world,
                        "nccl_used_for": ("barriers + timing all-reduce + set-up agreement" if p2p
                                          else "all-gather of the packed [B, k] lists + barriers + timing all-reduce"),
                        "stage2_exchange": ("nccl all-reduce(SUM) of the [B, C] scores" if os.environ.get("TS_P2P", "") == "0"
                                            else "scatter fused into the scoring kernel (owners store into every rank's matrix) + wait-take")}
    _PENDING["line"] = line

    # ---- parity of the timed workload -------------------------------------------
    if not args.no_parity and B >= PLANT_QUERIES:
        try:
            par = check_stage1(idx, lo, n_local, q_dev, res_s, res_i, pos, k, rank)
        except Exception as e:                               # noqa: BLE001
            par = {"error": f"{type(e).__name__}: {e}"[:160]}
        par["checked"] = True
        par["ok"] = bool(par.get("planted_ok")) and bool(par.get("oracle_ok", True)) and "error" not in par
        line["parity"] = par
        roof["parity_ok"] = par["ok"]
        if dist_on:
            dist.barrier()      # rank 0 checked alone for seconds: re-align before the next exchange step

    # ---- side measurements: the other batch sizes, Stage 2, the C5 chain ------------
    if not args.no_extra:
        also = {}
        try:
            also.update(side_stage1(sharded, idx, d, k, dev, pk, ld, n_local, world, dist_on))
            del sharded, idx
            torch.cuda.empty_cache()
            also["s2_c4"] = side_stage2(dev, pk, rank, world, dist_on, parity=not args.no_parity)
            torch.cuda.empty_cache()
            also["c5"] = side_c5(dev, rank, world, dist_on)
        except Exception as e:                               # noqa: BLE001 -- never at the expense of the result line
            also["error"] = f"{type(e).__name__}: {e}"[:200]
        roof["also"] = also
        if "s2_c4" in also and "oracle_ok" in also["s2_c4"] and "parity" in line:
            line["parity"]["stage2_ok"] = also["s2_c4"]["oracle_ok"]
            line["parity"]["ok"] = line["parity"]["ok"] and bool(also["s2_c4"]["oracle_ok"])
            roof["parity_ok"] = line["parity"]["ok"]
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import cpu_baseline

        cb = cpu_baseline.stage1_queries_per_s(N, d, B, k, sample_rows=1_000_000, reps=3)
        cb["seconds_per_step_sample"] = r3(cb["seconds_per_step_sample"])
        cb["cpu_model"] = cpu_baseline.host_info()["cpu_model"]
        s2 = cpu_baseline.stage2_candidates_per_s(2000)
        cb["stage2_cand_per_s"] = r3(s2["value"])
        line["cpu_baseline"] = cb
    if dist_on:
        dist.barrier()
    if rank == 0:
        print(json.dumps(line), flush=True)
    finish(0)


if __name__ == "__main__":
    main()
