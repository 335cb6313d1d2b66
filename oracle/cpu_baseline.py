"""CPU baseline legs for bench.py (TEST/MEASUREMENT INFRASTRUCTURE, see
oracle/__init__.py): the reference's Stage-1/Stage-2 arithmetic timed on the
host cores, on a bounded sample of the benchmark workload.

Stage 1: restated ``faiss.IndexFlatIP.search`` -- fp32 ``Q @ X.T`` through
torch/MKL on all host threads (FAISS itself calls BLAS sgemm for batches and a
SIMD dot loop for single queries) followed by an exact top-k; labelled
"restatement, FAISS not installed" (faiss-cpu is absent and unpinned,
/root/reference/requirements.txt:10; call site src/stage1_retriever.py:380).
The scan is O(N): the sample is a slice of ``sample_rows`` rows and the rate is
scaled linearly to the full corpus.

Stage 2: the per-candidate loop of ``rescore_candidates``
(/root/reference/src/stage2_rescorer.py:268-276) with a restated
``_maxsim_score`` (:167-183) in torch fp32 -- five small ATen calls and one
``.item()`` per candidate, exactly the reference's dispatch pattern.
"""
from __future__ import annotations

import os
import time


def host_info():
    import torch

    model = "unknown"
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    model = line.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    return {"cpu_count": os.cpu_count(), "torch_threads": torch.get_num_threads(), "cpu_model": model}


def stage1_queries_per_s(n_total: int, dim: int, B: int, k: int, sample_rows: int = 1_000_000,
                         reps: int = 3, seed: int = 1234):
    """q/s of an exact fp32 top-k over n_total rows, measured on sample_rows rows."""
    import torch

    sample_rows = min(sample_rows, n_total)
    g = torch.Generator().manual_seed(seed)
    X = torch.randn((sample_rows, dim), generator=g, dtype=torch.float32)
    X /= X.norm(dim=1, keepdim=True) + 1e-8
    Q = torch.randn((B, dim), generator=g, dtype=torch.float32)
    Q /= Q.norm(dim=1, keepdim=True) + 1e-8
    best = float("inf")
    kk = min(k, sample_rows)
    for _ in range(reps):
        t0 = time.perf_counter()
        S = Q @ X.T
        torch.topk(S, kk, dim=1)
        best = min(best, time.perf_counter() - t0)
    t_full = best * (n_total / sample_rows)
    return {"value": B / t_full, "unit": "queries/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"restated IndexFlatIP (fp32 torch/MKL Q@X.T + topk), {sample_rows} of {n_total} rows x {dim}, "
                      f"B={B}, k={k}, best of {reps}, scaled linearly in N; FAISS not installed",
            "seconds_per_step_sample": best}


def stage2_candidates_per_s(n_cand: int = 2000, Lq: int = 32, dim: int = 128, ld_lo: int = 16, ld_hi: int = 180,
                            seed: int = 77):
    """cand/s of the reference's per-candidate loop (torch fp32, .item() each)."""
    import torch
    import torch.nn.functional as F

    g = torch.Generator().manual_seed(seed)
    q = torch.randn((1, Lq, dim), generator=g)
    lens = torch.randint(ld_lo, ld_hi + 1, (n_cand,), generator=g).tolist()
    docs = [torch.randn((L, dim), generator=g) for L in lens]
    t0 = time.perf_counter()
    acc = 0.0
    for d in docs:
        qn = F.normalize(q, p=2, dim=-1)
        dn = F.normalize(d, p=2, dim=-1)
        sim = torch.matmul(qn.squeeze(0), dn.squeeze(0).T)
        acc += float(torch.mean(torch.max(sim, dim=-1)[0]).item())
    dt = time.perf_counter() - t0
    return {"value": n_cand / dt, "unit": "candidates/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"per-candidate _maxsim_score loop, {n_cand} candidates, Lq={Lq}, Ld~U[{ld_lo},{ld_hi}], dim={dim}, fp32",
            "checksum": acc}
