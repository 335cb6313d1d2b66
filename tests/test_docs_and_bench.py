"""CPU: the binding stub printed in INTEGRATION.md really drives the C ABI (executed here against the
CPU emulator build of the library, tests/cudasim -- test infrastructure), and the host-side helpers
of bench.py (contract fields, clock-sample parsing, algorithmic bytes) behave as documented."""
import ctypes as C
import json
import os
import re
import subprocess
import sys
import types

import numpy as np
import pytest

from oracle import flat_ip

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_integration_md_binding_stub_runs_against_the_abi(monkeypatch):
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"```python\n# src/_tristage\.py.*?\n(.*?)```", text, flags=re.S)
    assert m, "INTEGRATION.md lost its binding stub"
    code = m.group(1)
    env = dict(os.environ)
    env.pop("CXX", None)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cudasim"), "-j", "8"], env=env, stdout=subprocess.DEVNULL)
    sim = os.path.join(ROOT, "build", "cudasim", "libtristage_cudasim.so")
    real_cdll = C.CDLL
    monkeypatch.setattr(C, "CDLL", lambda name, *a, **kw: real_cdll(sim if name == "libtristage.so" else name, *a, **kw))
    mod = types.ModuleType("_tristage_stub")
    exec(compile(code, "INTEGRATION.md", "exec"), mod.__dict__)
    rng = np.random.default_rng(0)
    X = flat_ip.normalize_rows(rng.standard_normal((700, 48)).astype(np.float32)).astype(np.float32)
    Q = flat_ip.normalize_rows(rng.standard_normal((3, 48)).astype(np.float32)).astype(np.float32)
    idx = mod.IndexFlatIP(48)
    with pytest.raises(ValueError, match=r"No documents indexed\. Call add_documents\(\) first\."):
        idx.search(Q, 5)                                  # TS_ERR_EMPTY carries the reference's message
    idx.add(X[:300])
    idx.add(X[300:])
    assert idx.ntotal == 700
    D, I = idx.search(Q, 10)
    Xr, Qr = flat_ip.round_to(X, "bf16"), flat_ip.round_to(Q, "bf16")
    rD, rI = flat_ip.topk_desc(Qr @ Xr.T, 10)
    assert not flat_ip.check_topk(D, I, lambda b, ids: Xr[ids].astype(np.float64) @ Qr[b].astype(np.float64), rD, rI)
    with pytest.raises(RuntimeError):
        idx.search(Q, 10_000)                             # k beyond TS_MAX_K
    # the IVF half of the stub, driven the way the reference's _create_faiss_index does (:263-273)
    ivf = mod.IndexIVFFlat(mod.IndexFlatIP(48), 48, 6, mod.METRIC_INNER_PRODUCT)
    ivf.train(X)
    ivf.add(X)
    ivf.nprobe = 6                                        # every list probed: the exact result
    D2, I2 = ivf.search(Q, 10)
    assert ivf.is_trained and ivf.ntotal == 700
    assert not flat_ip.check_topk(D2, I2, lambda b, ids: Xr[ids].astype(np.float64) @ Qr[b].astype(np.float64), rD, rI)
    ivf.nprobe = 2
    D3, I3 = ivf.search(Q, 10)
    assert (D3[:, 0] <= D2[:, 0] + 1e-6).all() and (I3 >= -1).all()


def test_bench_contract_helpers():
    sys.path.insert(0, ROOT)
    import bench

    cfg = bench.workload_config(10_000_000, 1024, 100, 32, 8)
    assert cfg["rows"] == 10_000_000 and cfg["batch"] == 32 and cfg["n_gpus"] == 8 and "10000000x1024" in cfg["workload"]
    # both arms print the SAME config for the same arguments (the driver compares them)
    assert cfg == bench.workload_config(10_000_000, 1024, 100, 32, 8)
    pos = bench.plant_positions(10_000_000)
    assert pos.shape == (bench.PLANT_QUERIES, bench.PLANT_ROWS) and len(set(pos.ravel().tolist())) == pos.size
    assert (pos == bench.plant_positions(10_000_000)).all()
    # algorithmic bytes of one search per GPU: shard once + queries + candidate lists
    assert bench.stage1_alg_bytes(1_250_000, 1024, 32, 100, 148) == 1_250_000 * 2048 + 32 * 2048 + 148 * 32 * 100 * 8
    pk = bench.peaks()
    assert pk["hbm_gbs"] > 1000 and pk["bf16_tflops"] >= pk["bf16_tflops_sustained"] > 100


def test_bench_parity_check_accepts_the_oracle_and_rejects_a_wrong_result():
    """bench.check_stage1 (the oracle check of the TIMED index) on a CPU stand-in for the index: the oracle's own
    answer passes, a result with one id swapped for a low-scoring row or one score off by 1 % fails."""
    sys.path.insert(0, ROOT)
    import torch

    import bench

    rng = np.random.default_rng(3)
    N, d, B, k = 5000, 64, 4, 20
    q = flat_ip.normalize_rows(rng.standard_normal((B, d)).astype(np.float32)).astype(np.float32)
    X = flat_ip.normalize_rows(rng.standard_normal((N, d)).astype(np.float32)).astype(np.float32)
    pos = rng.choice(N, bench.PLANT_QUERIES * bench.PLANT_ROWS, replace=False).reshape(bench.PLANT_QUERIES, bench.PLANT_ROWS)
    k = bench.PLANT_ROWS
    for b in range(bench.PLANT_QUERIES):
        X[pos[b]] = flat_ip.normalize_rows(q[b][None, :] + 0.7 * rng.standard_normal((bench.PLANT_ROWS, d)).astype(np.float32) / d ** 0.5)
    Xr, Qr = flat_ip.round_to(X, "bf16"), flat_ip.round_to(q, "bf16")

    class FakeIndex:
        def get_rows(self, start, n):
            return Xr[start:start + n]

    D, I = flat_ip.topk_desc(Qr @ Xr.T, k)
    qd = torch.from_numpy(q).to(torch.bfloat16)
    good = bench.check_stage1(FakeIndex(), 0, N, qd, torch.from_numpy(D), torch.from_numpy(I), pos, k, rank=0)
    assert good["planted_ok"] and good["oracle_ok"] and good["oracle_rows"] == N
    I2 = I.copy()
    I2[2, 5] = int(np.argmin(Qr[2] @ Xr.T))                     # a clearly wrong row among query 2's results
    bad = bench.check_stage1(FakeIndex(), 0, N, qd, torch.from_numpy(D), torch.from_numpy(I2), pos, k, rank=0)
    assert not bad["oracle_ok"]
    D2 = D.copy()
    D2[3, 0] *= 1.01
    assert not bench.check_stage1(FakeIndex(), 0, N, qd, torch.from_numpy(D2), torch.from_numpy(I), pos, k, rank=0)["oracle_ok"]
    I3 = I.copy()
    I3[0, -1] = int(np.setdiff1d(np.arange(N), pos[0])[0])      # a planted row lost
    assert not bench.check_stage1(FakeIndex(), 0, N, qd, torch.from_numpy(D), torch.from_numpy(I3), pos, k, rank=0)["planted_ok"]


def test_clock_sampler_parses_nvidia_smi_rows(tmp_path, monkeypatch):
    sys.path.insert(0, ROOT)
    import bench

    fake = tmp_path / "nvidia-smi"
    rows = ["1965, 1965, 310.5, 3, 0x0000000000000000, Not Active, Not Active, Not Active, Not Active",
            "1620, 1965, 998.1, 100, 0x0000000000000004, Not Active, Not Active, Not Active, Active",
            "1590, 1965, 1001.0, 100, 0x0000000000000004, Not Active, Not Active, Not Active, Active",
            "garbage line"]
    fake.write_text("#!/bin/sh\n" + "".join(f"echo '{r}'\n" for r in rows) + "sleep 30\n")
    fake.chmod(0o755)
    monkeypatch.setenv("PATH", f"{tmp_path}:{os.environ['PATH']}")
    s = bench.ClockSampler("GPU-00000000")
    import time
    time.sleep(0.5)
    out = s.stop()
    assert out["sm_max_mhz"] == 1965.0 and out["sm_mhz"] == 1605.0 and out["samples"] == 2     # median of the busy samples
    assert out["reasons"] == ["sw_power_cap"]


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--rows", "200000", "--dim", "64", "--batch", "4"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["config"]["rows"] == 200000 and line["config"]["batch"] == 4


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree not present")
def test_unmodified_reference_orchestrator_runs_over_the_drop_in_stages():
    """INTEGRATION.md way A end to end: /root/reference/src/retrieval_pipeline.py, unmodified, with the two stage
    modules aliased to this package, reproduces the output of the unmodified reference classes
    (tools/dropin_check.py; the kernels run on the CPU emulator build here)."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "dropin_check.py"), "--emulate"], capture_output=True,
                         text=True, timeout=900)
    assert out.returncode == 0 and "drop-in ok" in out.stdout, (out.stdout[-1500:], out.stderr[-3000:])


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree not present")
def test_unmodified_reference_stage1_runs_over_the_faiss_shaped_module():
    """INTEGRATION.md way C: the reference's own src/stage1_retriever.py with sys.modules["faiss"] set to
    tristage_rag_b200.faiss_compat reproduces the reference's Stage-1 output (flat branch, persistence) and its
    IVF branch agrees with oracle/ivf.py (tools/faiss_shim_check.py; kernels on the CPU emulator build here)."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "faiss_shim_check.py"), "--emulate"], capture_output=True,
                         text=True, timeout=900)
    assert out.returncode == 0 and "faiss shim ok" in out.stdout, (out.stdout[-1500:], out.stderr[-3000:])
