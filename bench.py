#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 candidate-scoring hot path.

Metric (BASELINE.json): Stage-1 queries/s, exact top-100 over a 10M x 1024 bf16
corpus (configs[2]; fits one B200: 20.5 GB), reported with the HBM-roofline
fraction of the scan kernel; Stage-2 candidates/s (configs[3]) and the other
batch sizes ride along under "extra".

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...     # the CPU arm (oracle port, host cores)

A step = one search call of B queries over the whole (row-sharded) corpus:
query prep + scan (fused top-k) + merge (+ all-gather + merge for N > 1).
Inputs are synthetic (seeded randn, row-normalised with the reference formula).
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="tristage", choices=["tristage", "reference"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--path", default="auto")
    ap.add_argument("--no-extra", action="store_true", help="skip the B=1 / B=1024 / Stage-2 side measurements")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-variants", action="store_true",
                    help="skip the child-process side jobs on unvalidated variants (use under a profiler)")
    ap.add_argument("--graph", action="store_true",
                    help="also time the step replayed from a CUDA graph and report the faster launch mode "
                         "(off by default: measured within 1%% of eager launches on B200)")
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
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while work runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,utilization.gpu,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, uuid: str):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", uuid, f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
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


def build_shard(idx, lo, hi, dim, dev, seed, chunk=500_000, dtype="bf16"):
    import torch

    g = torch.Generator(device=dev).manual_seed(seed)
    tdt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[dtype]
    for s in range(lo, hi, chunk):
        n = min(chunk, hi - s)
        x = torch.randn((n, dim), generator=g, device=dev, dtype=torch.float32)
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


def stage1_alg_bytes(n_rows, ld, B, k, L):
    return n_rows * ld * 2 + B * ld * 2 + L * B * k * 8


def workload_config(N, d, k, B, world):
    """The `config` both arms print: same workload keys for the GPU arm and the CPU reference arm."""
    return {"workload": f"exact top-{k} over {N}x{d} bf16 corpus (row-sharded over {world} GPU), query batch {B}",
            "rows": N, "dim": d, "k": k, "batch": B}


def finish(code: int = 0):
    """Leave without running interpreter/NCCL teardown: destroying a process group (or CUDA
    graphs that captured NCCL work) can block for minutes after the result line is out."""
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(code)


_T0 = time.perf_counter()       # process start: the side jobs stop being launched once the whole run is 7 min old
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
    """CPU arm: the reference's Stage-1 arithmetic (oracle port; FAISS absent) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_baseline

    info = cpu_baseline.host_info()
    sample_rows = 500_000
    vals = []
    r = None
    for _ in range(max(1, min(args.steps, 5)) + min(args.warmup, 1)):
        r = cpu_baseline.stage1_queries_per_s(args.rows, args.dim, args.batch, args.k, sample_rows=sample_rows, reps=1)
        vals.append(r["value"])
    value = max(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": args.batch / value * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {**workload_config(args.rows, args.dim, args.k, args.batch, args.gpus),
                   "path": "cpu: fp32 torch/MKL Q@X.T + topk (restated IndexFlatIP; FAISS not installable here)",
                   "note": "the reference stores fp32; each step scans a 500k-row slice on all host threads and the "
                           "rate is scaled linearly to the full corpus (the scan is O(N)); the CPU arm does not "
                           "shard, so the same number is printed for every --gpus"},
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": r["cores"], "kind": "port", "sample": r["sample"],
                         "host": info},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


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
    idx = _lib.Index(d, "bf16", "ip", local, reserve_rows=hi - lo)
    build_shard(idx, lo, hi, d, dev, seed=1234 + rank)
    sharded = ShardedIndex(idx, N)
    q_pin, q_dev = make_queries(B, d, dev)
    torch.cuda.synchronize(dev)

    uuid = str(torch.cuda.get_device_properties(dev).uuid)
    sampler = ClockSampler(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)

    # ---- device-resident timing (value) --------------------------------------
    path = args.path
    step = lambda: sharded.search(q_dev, k, path=path)     # noqa: E731
    timed(step, 2, args.warmup, dev, dist_on)               # warm-up incl. scratch allocation
    # scan-kernel duration: CUDA events around the kernel on its own stream, eager launches
    idx.set_profiling(True)
    l0 = idx.launches
    ms_eager = timed(step, args.steps, 0, dev, dist_on)
    # + the post-all-gather merge kernel (or push + wait-merge of the peer-memory exchange)
    launches = idx.launches - l0 + (args.steps * (2 if getattr(sharded, "_p2p", False) else 1) if world > 1 else 0)
    scan_ms, scan_n = idx.scan_time_ms()
    idx.set_profiling(False)
    # headline: the same step captured once in a CUDA graph (query prep, threshold pre-pass,
    # scan, select, all-gather, merge) and replayed K times -- no per-launch host latency
    ms, launch_mode = ms_eager, "eager"
    extra_modes = {}
    if args.graph and not getattr(sharded, "_p2p", False):    # the exchange's step counter comes from the host
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(3):
                    step()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                g_out = step()
            replay = lambda: graph.replay()                 # noqa: E731
            ms_graph = timed(replay, args.steps, args.warmup, dev, dist_on)
            ref_s, ref_i = step()
            torch.cuda.synchronize(dev)
            assert torch.equal(g_out[1], ref_i) and torch.equal(g_out[0], ref_s), "graph replay diverged from eager"
            # the two launch modes drive identical kernels; keep whichever the box runs faster
            # (a second eager pass guards against clock drift between the passes)
            ms_eager = min(ms_eager, timed(step, args.steps, 2, dev, dist_on))
            if ms_graph < ms_eager:
                ms, launch_mode = ms_graph, "cuda_graph"
            else:
                ms, launch_mode = ms_eager, "eager"
            extra_modes = {"ms_per_step_cuda_graph": ms_graph / args.steps}
        except Exception as e:                               # noqa: BLE001
            print(f"[bench] CUDA-graph capture unavailable ({type(e).__name__}: {e}); eager timing kept", file=sys.stderr)
    value = B * args.steps / (ms / 1e3)

    # ---- end-to-end timing (host buffers in, host results out) ---------------
    # inputs and results live in pinned host memory, allocated once like a serving process would
    q_stage = torch.empty_like(q_dev)
    h_scores = torch.empty((B, k), dtype=torch.float32).pin_memory()
    h_ids = torch.empty((B, k), dtype=torch.int64).pin_memory()
    q_np, out_np = q_pin.numpy(), (h_scores.numpy(), h_ids.numpy())

    def e2e_step():
        if world == 1:
            idx.search_host(q_np, k, path=path, out=out_np)  # the C-ABI host call: H2D, search, D2H, sync
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
            "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
            "peak_source": pk["source"], "traffic": None, "algorithmic_bytes_per_launch": alg_bytes,
            "kernel_ms": scan_ms, "kernel_launches_timed": scan_n,
            "frac_of_nominal_8TBs": achieved / 8000.0}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof) and world == 1 and (N, d) == (10_000_000, 1024):
        with open(prof) as f:                      # one ncu --set full capture of this exact workload
            roof["traffic"] = json.load(f).get(roof["kernel"])

    line = {
        "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {**workload_config(N, d, k, B, world), "path": roof["kernel"], "parallelism": f"rowshard{world}",
                   "l2": "inputs larger than L2 (shard >= 2.5 GB vs 126 MB), no flush needed",
                   "launch": launch_mode, "ms_per_step_eager": ms_eager / args.steps, **extra_modes,
                   "merge": ("peer-memory exchange" if getattr(sharded, "_p2p", False) else
                             ("nccl all-gather + merge kernel" if world > 1 else "none"))},
        "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": B * d * 4,
                "d2h_bytes_per_step": B * k * 12, "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches), "roofline": roof, "clocks": clocks,
    }

    # ---- side measurements and CPU baseline: rank 0, single GPU only ---------
    if world == 1 and not args.no_extra:
        line["extra"] = extras(idx, d, k, dev, pk, ld, n_local, sm)
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import cpu_baseline

        cb = cpu_baseline.stage1_queries_per_s(N, d, B, k, sample_rows=1_000_000, reps=3)
        cb["host"] = cpu_baseline.host_info()
        line["cpu_baseline"] = cb
    if dist_on:
        dist.barrier()
    _PENDING["line"] = line
    if rank == 0 and world == 1 and not args.no_extra and not args.no_variants:
        try:
            line["extra"]["unvalidated_variants"] = variant_probes()
        except Exception as e:                               # noqa: BLE001 -- never at the expense of the result line
            line["extra"]["unvalidated_variants"] = {"status": f"{type(e).__name__}: {e}"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    finish(0)


def variant_probes():
    """First hardware numbers of what was written after the round-1 GPU budget ended (approximate mode, tf32
    tensor path for fp32 storage, the TS_FUSE / TS_S2_V2 variants, the select-kernel rewrite): each job runs in
    its OWN process with a short timeout, after the headline is measured, so a fault in unvalidated code cannot
    touch this process's CUDA context or the result line.  Device-vs-device self-checks only (no oracle here).
    Every mbarrier wait in the kernels is bounded (a protocol error traps after ~4 s instead of hanging) and a
    fault is confined to the child's context.  The peer-memory exchange (TS_P2P) needs several GPUs and is not here."""
    jobs = {
        "approximate_mode_1Mx768_bf16": [sys.executable, os.path.join(ROOT, "tools", "ivf_probe.py"), "--rows", "1000000", "--dim", "768",
                                         "--batches", "1,32", "--steps", "20", "--selfcheck"],
        "fp32_storage_1Mx768_cuda_core_scan_and_tf32_tensor_path": [sys.executable, os.path.join(ROOT, "tools", "perf_probe.py"), "--rows", "1000000",
                                                 "--dim", "768", "--dtype", "fp32", "--paths", "stream,umma", "--batches", "1,4,32,1024",
                                                 "--steps", "10", "--selfcheck"],
    }
    ab = os.path.join(ROOT, "tools", "variant_ab.py")
    # opt-in variants against the validated default, same process, results compared bit for bit on the device
    jobs["stage1_TS_FUSE_and_select_rewrite_AB_1.25Mx1024"] = [sys.executable, ab, "--what", "s1"]
    jobs["stage2_TS_S2_V2_and_TS_S2_EPI2_AB_config4"] = [sys.executable, ab, "--what", "s2"]
    jobs["stage1_TS_PAIR_AB_4Mx1024"] = [sys.executable, ab, "--what", "pair"]      # last: the least rehearsed protocol
    out = {}
    t_start = time.perf_counter()
    for name, cmd in jobs.items():
        rec = {"status": "not run"}
        now = time.perf_counter()
        if now - t_start > 200 or now - _T0 > 420:           # budget of the side jobs / age of the whole run
            out[name] = {"status": "skipped: side-measurement budget used up"}
            continue
        try:
            p = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT)
            status = None
            try:
                so, se = p.communicate(timeout=75)
            except subprocess.TimeoutExpired:
                p.kill()
                status = "timeout (75 s), killed"
                try:
                    so, se = p.communicate(timeout=10)       # whatever it printed before the limit
                except subprocess.TimeoutExpired:
                    so, se = "", ""
            rows = []
            for ln in so.splitlines():
                if ln.startswith("{"):
                    try:
                        rows.append(json.loads(ln))
                    except ValueError:
                        pass
            rec = {"status": status or ("ok" if p.returncode == 0 else f"exit {p.returncode}"), "lines": rows}
            if rec["status"] != "ok":
                rec["stderr_tail"] = se[-400:]
        except Exception as e:                               # noqa: BLE001
            rec = {"status": f"{type(e).__name__}: {e}"}
        out[name] = rec
    return out


def extras(idx, d, k, dev, pk, ld, n_local, sm):
    """B=1 (stream kernel) and B=1024 (tensor-bound) Stage-1 rates, and Stage-2 MaxSim cand/s (config #4)."""
    import numpy as np
    import torch

    from tristage_rag_b200 import _lib

    out = {}
    for B, steps in ((1, 20), (1024, 5)):
        _, qd = make_queries(B, d, dev, seed=99 + B)
        fn = lambda: idx.search(qd, k)                       # noqa: E731
        timed(fn, 1, 2, dev, False)
        idx.set_profiling(True)
        ms = timed(fn, steps, 0, dev, False)
        scan_ms, _ = idx.scan_time_ms()
        idx.set_profiling(False)
        rec = {"queries_per_s": B * steps / (ms / 1e3), "ms_per_step": ms / steps, "scan_kernel_ms": scan_ms}
        if B == 1:
            L = min(2 * sm, max(1, n_local // 32))
            gb = stage1_alg_bytes(n_local, ld, B, k, L) / (scan_ms / 1e3) / 1e9
            rec.update(bound="hbm", achieved_gbs=gb, frac=gb / pk["hbm_gbs"], frac_of_nominal_8TBs=gb / 8000.0)
        else:
            tf = 2.0 * B * n_local * ld / (scan_ms / 1e3) / 1e12
            rec.update(bound="tensor", achieved_tflops=tf, frac=tf / pk["bf16_tflops_sustained"],
                       frac_of_burst_peak=tf / pk["bf16_tflops"])
        out[f"stage1_B{B}"] = rec

    # Stage 2, config #4: 64 queries x 1000 candidates, Lq 32, Ld ~ U[16,180], dim 128, 1M-doc store
    ndocs, dim, Bq, C, Lq = 1_000_000, 128, 64, 1000, 32
    g = torch.Generator(device=dev).manual_seed(77)
    rng = np.random.default_rng(77)
    lens = rng.integers(16, 181, size=ndocs).astype(np.int32)
    st = _lib.TokStore(dim, "bf16", dev.index, reserve_docs=ndocs, reserve_tokens=int(lens.sum()))
    chunk = 100_000
    for s in range(0, ndocs, chunk):
        ln = lens[s:s + chunk]
        t = torch.randn((int(ln.sum()), dim), generator=g, device=dev)
        t = torch.nn.functional.normalize(t, dim=-1).to(torch.bfloat16)
        st.add(t, ln, normalize=False)
        del t
    qt = torch.nn.functional.normalize(torch.randn((Bq, Lq, dim), generator=g, device=dev), dim=-1).to(torch.bfloat16)
    cand = torch.stack([torch.randperm(ndocs, generator=g, device=dev)[:C] for _ in range(Bq)])
    fn2 = lambda: st.maxsim(qt, cand, normalize_q=False)     # noqa: E731
    timed(fn2, 1, 3, dev, False)
    st.set_profiling(True)
    steps = 20
    ms = timed(fn2, steps, 0, dev, False)
    kms, _ = st.scan_time_ms()
    st.set_profiling(False)
    tok_bytes = float(lens[cand.cpu().numpy()].astype(np.int64).sum()) * dim * 2
    gb = tok_bytes / (kms / 1e3) / 1e9
    out["stage2_maxsim"] = {"candidates_per_s": Bq * C * steps / (ms / 1e3), "ms_per_step": ms / steps,
                            "kernel_ms": kms, "bound": "hbm", "algorithmic_bytes_per_launch": tok_bytes,
                            "achieved_gbs": gb, "frac": gb / pk["hbm_gbs"],
                            "workload": "64 q x 1000 cand, Lq 32, Ld~U[16,180], dim 128, bf16, 1M-doc store"}
    from oracle import cpu_baseline

    out["stage2_cpu_baseline"] = cpu_baseline.stage2_candidates_per_s(2000)
    return out


if __name__ == "__main__":
    main()
