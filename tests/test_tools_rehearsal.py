"""CPU: the measurement tools that bench.py runs as child processes on the GPU box (tools/ivf_probe.py,
tools/perf_probe.py, tools/variant_ab.py) rehearsed end to end on the emulator build of the kernels, at toy sizes:
their Python (argument handling, tensor shapes, self-checks, JSON lines) must not be what fails on the first hardware
run.  "cuda" tensors are host tensors here and timings are meaningless; test infrastructure only."""
import contextlib
import ctypes as C
import importlib.util
import io
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def fake_cuda(sim, monkeypatch):
    """Run tool code written for cuda tensors over the emulated library: device = cpu, same C-ABI calls."""
    import bench
    from tristage_rag_b200 import _lib

    real_device = torch.device
    monkeypatch.setattr(torch, "device", lambda *a, **kw: real_device("cpu"))
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **kw: None)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **kw: self)

    class Ev:
        def __init__(self, **kw):
            pass

        def record(self):
            pass

        def elapsed_time(self, other):
            return 1.0

    monkeypatch.setattr(torch.cuda, "Event", Ev)

    def timed(fn, steps, warmup, dev, dist_on):
        for _ in range(min(steps + warmup, 2)):
            fn()
        return float(max(steps, 1))

    monkeypatch.setattr(bench, "timed", timed)
    monkeypatch.setattr(bench, "arm_watchdog", lambda s: None)
    monkeypatch.setattr(bench, "make_queries", lambda B, dim, dev, seed=4321: (lambda q: (q, q))(
        torch.nn.functional.normalize(torch.randn((B, dim), generator=torch.Generator().manual_seed(seed)), dim=1)))
    code = {torch.float32: _lib.TS_F32, torch.bfloat16: _lib.TS_BF16, torch.float16: _lib.TS_F16}

    def index_search(self, q, k, normalize_q=False, path="auto"):
        q = q.contiguous()
        s, i = torch.empty((q.shape[0], k)), torch.empty((q.shape[0], k), dtype=torch.int64)
        _lib.check(sim.ts_index_search(self._h, C.c_void_p(q.data_ptr()), code[q.dtype], q.shape[0], int(k),
                                       1 if normalize_q else 0, _lib.PATHS[path], C.c_void_p(s.data_ptr()),
                                       C.c_void_p(i.data_ptr()), None))
        return s, i

    def ivf_search(self, q, k, nprobe, normalize_q=False):
        q = q.contiguous()
        s, i = torch.empty((q.shape[0], k)), torch.empty((q.shape[0], k), dtype=torch.int64)
        _lib.check(sim.ts_ivf_search(self._h, C.c_void_p(q.data_ptr()), code[q.dtype], q.shape[0], int(k), int(nprobe),
                                     1 if normalize_q else 0, C.c_void_p(s.data_ptr()), C.c_void_p(i.data_ptr()), None))
        return s, i

    def maxsim(self, q_tok, cand, q_len=None, n_cand=None, mode=0, normalize_q=True):
        q_tok, cand = q_tok.contiguous(), cand.contiguous()
        out = torch.empty(cand.shape)
        _lib.check(sim.ts_maxsim(self._h, C.c_void_p(q_tok.data_ptr()), code[q_tok.dtype], None, q_tok.shape[0], q_tok.shape[1],
                                 C.c_void_p(cand.data_ptr()), None, cand.shape[1], int(mode), 1 if normalize_q else 0,
                                 C.c_void_p(out.data_ptr()), None))
        return out

    monkeypatch.setattr(_lib.Index, "search", index_search)
    monkeypatch.setattr(_lib.IVF, "search", ivf_search)
    monkeypatch.setattr(_lib.TokStore, "maxsim", maxsim)
    monkeypatch.setenv("HOSTSIM_SM_COUNT", "8")


def run_tool(name, argv, monkeypatch):
    spec = importlib.util.spec_from_file_location(f"tool_{name}", os.path.join(ROOT, "tools", f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    monkeypatch.setattr(sys, "argv", [name] + argv)
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        mod.main()
    return [json.loads(ln) for ln in out.getvalue().splitlines() if ln.startswith("{")]


def test_ivf_probe_runs_end_to_end(fake_cuda, monkeypatch):
    rows = run_tool("ivf_probe", ["--rows", "3000", "--dim", "64", "--nlist", "12", "--nprobe", "3", "--k", "10", "--batches", "1,5",
                                  "--steps", "1", "--selfcheck"], monkeypatch)
    kinds = [r["what"] for r in rows]
    assert kinds.count("ivf") == 2 and kinds.count("exact") == 2 and kinds.count("grid sweep B=1") == 4
    assert kinds.count("batch order B=32") == 2 and "TS_IVF_NOORDER" not in os.environ
    chk = next(r for r in rows if r["what"] == "selfcheck")
    assert chk["all_lists_ids_equal_exact"] > 0.95 and chk["probed_rows_in_probed_lists"] == 1.0
    assert chk["probe_best_le_exact_best"] and chk["scores_descending"]
    assert all(r["rows_read_per_step"] > 0 for r in rows if r["what"] in ("ivf", "exact"))


def test_perf_probe_fp32_both_paths_runs_end_to_end(fake_cuda, monkeypatch):
    rows = run_tool("perf_probe", ["--rows", "2000", "--dim", "64", "--dtype", "fp32", "--paths", "stream,umma", "--batches", "1,4,9",
                                   "--steps", "1", "--k", "10", "--selfcheck"], monkeypatch)
    assert [(r["path"], r["B"]) for r in rows if "path" in r] == [("stream", 1), ("stream", 4), ("umma", 1), ("umma", 4), ("umma", 9)]
    chk = rows[-1]
    assert chk["what"] == "selfcheck" and chk["topk_overlap_umma_vs_stream"] > 0.9 and chk["best_id_equal"]


@pytest.mark.parametrize("what,argv,n", [("s1", ["--rows", "3000", "--dim", "64"], 4), ("pair", ["--rows", "2000", "--dim", "64", "--pair-batches", "130"], 1),
                                         ("s2", ["--ndocs", "120", "--cands", "40", "--queries", "3"], 4)])
def test_variant_ab_runs_end_to_end(fake_cuda, monkeypatch, what, argv, n):
    rows = run_tool("variant_ab", ["--what", what, "--steps", "1"] + argv, monkeypatch)
    assert len(rows) == n, rows
    for r in rows:
        assert "error" not in r, r
        assert r["bit_equal"] is True and r["speedup"] > 0, r
    assert all(os.environ.get(sw, "0") == "0" for r in rows for sw in r["switch"].split("+"))
