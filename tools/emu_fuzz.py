#!/usr/bin/env python
"""Randomised parity campaign on the CPU emulator build (tests/cudasim): random shapes through the
Stage-1 scans (tensor + stream) and the Stage-2 kernels (tensor, both epilogues, and SIMT) against
the oracle.  Development aid: `python tools/emu_fuzz.py --seconds 600 --seed 1`."""
import argparse
import ctypes as C
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import flat_ip, maxsim  # noqa: E402
from tristage_rag_b200 import _lib  # noqa: E402


def load():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cudasim"), "-j", "8"], stdout=subprocess.DEVNULL)
    L = C.CDLL(os.path.join(ROOT, "build", "cudasim", "libtristage_cudasim.so"))
    for name, (res, args) in _lib.SYMBOLS.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib._lib = L
    _lib._stream_ptr = lambda d: None


def machine(rng, max_sms=148):
    """Random emulated GPU: number of SMs (few SMs = many tiles per CTA) and timing mode."""
    sms = min(int(rng.choice([148, 148, 16, 4, 2])), max_sms)
    seed = int(rng.integers(1, 1 << 30)) if rng.random() < 0.5 else 0
    os.environ["HOSTSIM_SM_COUNT"] = str(sms)
    os.environ["CUDASIM_ASYNC"] = str(seed)
    return f"sms={sms} async={seed}"


def stage1_case(rng):
    variant = str(rng.choice(["", "", "TS_PAIR", "TS_DBG_NOSHARE", "TS_SELECT_V1", "TS_FUSE", "TS_DBG_STATIC", "TS_DBG_NOKTHSTART", "TS_DBG_NOSPREAD"]))
    hw = machine(rng, 16 if variant == "TS_FUSE" else 148)      # a cooperative grid keeps every CTA alive at once
    dtype = rng.choice(["bf16", "fp16", "fp32"], p=[0.6, 0.25, 0.15])
    N = int(rng.choice([1, 3, 50, 255, 256, 257, 1000, 4000, 12000]))
    N += int(rng.integers(0, 7))
    d = int(rng.choice([8, 24, 64, 72, 100, 128, 200, 256]))
    B = int(rng.choice([1, 2, 5, 8, 9, 33, 64, 65, 128, 129, 260]))
    k = int(rng.choice([1, 2, 7, 50, 100, 128, 129, 500, 512]))
    metric = rng.choice(["ip", "cosine"], p=[0.8, 0.2])
    path = "stream" if dtype == "fp32" else rng.choice(["umma", "stream"], p=[0.8, 0.2])
    if path == "stream":
        B = min(B, 9)
        N = min(N, 3000)
    if N * d * ((B + 127) // 128) > 6e6:
        N = max(1, int(6e6 / d / ((B + 127) // 128)))
    X = rng.standard_normal((N, d)).astype(np.float32)
    if metric == "ip":
        X = flat_ip.normalize_rows(X).astype(np.float32)
    else:
        X *= rng.uniform(0.1, 4.0, size=(N, 1)).astype(np.float32)
    flavour = rng.choice(["random", "dups", "planted", "scaled"])
    if flavour == "dups" and N > 4:
        X[rng.integers(0, N, size=max(1, N // 3))] = X[0]
    Q = flat_ip.normalize_rows(rng.standard_normal((B, d)).astype(np.float32)).astype(np.float32)
    if flavour == "planted" and N > 20:
        X[rng.integers(0, N, size=10)] = Q[0] * (1 if metric == "ip" else 2.5)
    idx = _lib.Index(d, dtype, metric, 0)
    for part in np.array_split(X, int(rng.integers(1, 4))):
        if len(part):
            idx.add(part)
    base = int(rng.choice([0, 0, 12345, 5_000_000_000]))
    idx.set_id_base(base)
    if variant:
        os.environ[variant] = "1"
    try:
        D, I = idx.search_host(Q, k, path=path)
    finally:
        if variant:
            os.environ[variant] = "0"
    Xr, Qr = flat_ip.round_to(X, dtype), flat_ip.round_to(Q, dtype)
    inv = (1.0 / (np.linalg.norm(X, axis=1) + 1e-8)).astype(np.float32) if metric == "cosine" else np.ones(N, np.float32)
    rD, rI = flat_ip.topk_desc((Qr @ Xr.T) * inv[None, :], k)
    sc = lambda b, ids: (Xr[ids].astype(np.float64) @ Qr[b].astype(np.float64)) * inv[ids]   # noqa: E731
    I0 = np.where(I >= 0, I - base, -1)
    bad = flat_ip.check_topk(D, I0, sc, rD, rI, rel=1e-3)
    # the 1e-3 RELATIVE rule is meaningless for scores that cancel to ~0 (fp32 accumulation order): drop
    # score complaints whose absolute error is below 2e-6 of the unit-norm scale, keep everything else
    keep = []
    for msg in bad:
        if "off by" in msg:
            g, o = float(msg.split("got ")[1].split(",")[0]), float(msg.split("oracle ")[1].rstrip(")"))
            if abs(g - o) < 2e-6 * max(1.0, float(np.abs(inv).max()) * float(np.linalg.norm(Xr, axis=1).max())):
                continue
        keep.append(msg)
    bad = keep
    return f"S1 N={N} d={d} B={B} k={k} {dtype} {metric} {path} {flavour} base={base} variant={variant} {hw}", bad


def tf32_case(rng):
    """fp32 storage on the tensor path (kind::tf32): against the oracle on tf32-rounded operands."""
    hw = machine(rng)
    N = int(rng.choice([1, 3, 255, 256, 257, 1000, 4000])) + int(rng.integers(0, 7))
    d = int(rng.choice([4, 12, 32, 36, 100, 128, 200, 768]))
    B = int(rng.choice([1, 2, 8, 9, 33, 64, 65, 128, 129, 260]))
    k = int(rng.choice([1, 7, 50, 100, 128, 129, 500, 512]))
    metric = rng.choice(["ip", "cosine"], p=[0.8, 0.2])
    if N * d * ((B + 127) // 128) > 4e6:
        N = max(1, int(4e6 / d / ((B + 127) // 128)))
    X = rng.standard_normal((N, d)).astype(np.float32)
    X = flat_ip.normalize_rows(X).astype(np.float32) if metric == "ip" else X * rng.uniform(0.1, 4.0, size=(N, 1)).astype(np.float32)
    Q = flat_ip.normalize_rows(rng.standard_normal((B, d)).astype(np.float32)).astype(np.float32)
    if N > 20:
        X[rng.integers(0, N, size=10)] = Q[0] * (1 if metric == "ip" else 2.5)
    idx = _lib.Index(d, "fp32", metric, 0)
    for part in np.array_split(X, int(rng.integers(1, 4))):
        if len(part):
            idx.add(part)
    variant = str(rng.choice(["", "", "TS_DBG_NOSHARE", "TS_PAIR"]))      # TS_PAIR must be ignored for fp32
    if variant:
        os.environ[variant] = "1"
    try:
        D, I = idx.search_host(Q, k, path="umma")
    finally:
        if variant:
            os.environ[variant] = "0"
    Xr, Qr = flat_ip.round_to(X, "tf32"), flat_ip.round_to(Q, "tf32")
    inv = (1.0 / (np.linalg.norm(X, axis=1) + 1e-8)).astype(np.float32) if metric == "cosine" else np.ones(N, np.float32)
    rD, rI = flat_ip.topk_desc((Qr @ Xr.T) * inv[None, :], k)
    sc = lambda b, ids: (Xr[ids].astype(np.float64) @ Qr[b].astype(np.float64)) * inv[ids]   # noqa: E731
    bad = flat_ip.check_topk(D, I, sc, rD, rI, rel=1e-3)
    keep = []
    for msg in bad:
        if "off by" in msg:
            g, o = float(msg.split("got ")[1].split(",")[0]), float(msg.split("oracle ")[1].rstrip(")"))
            if abs(g - o) < 2e-6 * max(1.0, float(np.abs(inv).max()) * float(np.linalg.norm(Xr, axis=1).max())):
                continue
        keep.append(msg)
    return f"TF32 N={N} d={d} B={B} k={k} {metric} variant={variant} {hw}", keep


def ivf_case(rng):
    """Approximate mode (csrc/ivf.cu): assignments, coarse probes and the list scan against oracle/ivf.py."""
    from oracle import ivf as oivf

    hw = machine(rng)
    dtype = str(rng.choice(["bf16", "fp16", "fp32"], p=[0.6, 0.2, 0.2]))
    N = int(rng.choice([1, 7, 100, 1000, 3000])) + int(rng.integers(0, 9))
    d = int(rng.choice([4, 8, 24, 64, 100, 136]))
    nlist = int(min(N, rng.choice([1, 2, 7, 33, 100])))
    nprobe = int(rng.integers(1, nlist + 1))
    B = int(rng.choice([1, 2, 5, 9, 40]))
    k = int(rng.choice([1, 7, 100, 128, 129, 512]))
    metric = str(rng.choice(["ip", "cosine"], p=[0.8, 0.2]))
    X = rng.standard_normal((N, d)).astype(np.float32)
    X = flat_ip.normalize_rows(X).astype(np.float32) if metric == "ip" else X * rng.uniform(0.1, 4.0, size=(N, 1)).astype(np.float32)
    if rng.random() < 0.3 and N > 4:
        X[rng.integers(0, N, size=max(1, N // 3))] = X[0]            # duplicates: ties resolve by id
    cent = X[rng.choice(N, size=nlist, replace=False)].astype(np.float32) + 1e-3 * rng.standard_normal((nlist, d)).astype(np.float32)
    Q = flat_ip.normalize_rows(rng.standard_normal((B, d)).astype(np.float32)).astype(np.float32)
    idx = _lib.Index(d, dtype, metric, 0)
    iv = _lib.IVF(idx, nlist)
    iv.set_centroids(cent)
    for part in np.array_split(X, int(rng.integers(1, 4))):
        if len(part):
            idx.add(part)
            if rng.random() < 0.5:
                iv.sync()
    base = int(rng.choice([0, 0, 5_000_000_000]))
    idx.set_id_base(base)
    D, I = iv.search_host(Q, k, nprobe)
    a = iv.assignments()
    Xr, Qr = flat_ip.round_to(X, dtype), flat_ip.round_to(Q, dtype)
    bad = []
    want, margin = oivf.assign_lists(Xr, cent), oivf.assign_margin(Xr, cent)
    scale = np.linalg.norm(Xr, axis=1) * np.linalg.norm(cent, axis=1).max()
    wrong = np.nonzero((a != want) & (margin > 1e-5 * np.maximum(scale, 1.0)))[0]
    if wrong.size:
        bad.append(f"assign rows {wrong[:4].tolist()} got {a[wrong[:4]].tolist()} want {want[wrong[:4]].tolist()}")
    lists, _ = iv.coarse_host(Q, nprobe)
    inv = (1.0 / (np.linalg.norm(X, axis=1) + 1e-8)).astype(np.float32) if metric == "cosine" else np.ones(N, np.float32)
    rD, rI = oivf.ivf_search(Xr * inv[:, None], Qr, a, lists, k)
    sc = lambda b, ids: (Xr[ids].astype(np.float64) @ Qr[b].astype(np.float64)) * inv[ids]   # noqa: E731
    I0 = np.where(I >= 0, I - base, -1)
    for msg in flat_ip.check_topk(D, I0, sc, rD, rI, rel=1e-3):
        if "off by" in msg:
            g, o = float(msg.split("got ")[1].split(",")[0]), float(msg.split("oracle ")[1].rstrip(")"))
            if abs(g - o) < 2e-6 * max(1.0, float(np.abs(inv).max()) * float(np.linalg.norm(Xr, axis=1).max())):
                continue
        bad.append(msg)
    if nprobe == nlist:                                              # every list probed: the exact search
        eD, eI = flat_ip.topk_desc((Qr @ Xr.T) * inv[None, :], k)
        bad += [m for m in flat_ip.check_topk(D, I0, sc, eD, eI, rel=1e-3) if "off by" not in m]
    return f"IVF N={N} d={d} nlist={nlist} nprobe={nprobe} B={B} k={k} {dtype} {metric} base={base} {hw}", bad


def stage2_case(rng):
    hw = machine(rng)
    dtype = rng.choice(["bf16", "fp16", "fp32"], p=[0.7, 0.2, 0.1])
    dim = int(rng.choice([8, 24, 64, 96, 128, 136, 256]))
    Lq = int(rng.choice([1, 2, 7, 31, 32, 33, 64, 65, 128]))
    ndocs = int(rng.integers(1, 120))
    style = rng.choice(["any", "tiny", "long", "mix"])
    lens = {"any": rng.integers(1, 257, size=ndocs), "tiny": rng.integers(1, 9, size=ndocs),
            "long": rng.integers(200, 257, size=ndocs), "mix": rng.choice([1, 8, 9, 63, 64, 65, 255, 256], size=ndocs)}[style]
    tok = rng.standard_normal((int(lens.sum()), dim)).astype(np.float32)
    st = _lib.TokStore(dim, dtype, 0)
    st.add(tok, lens, normalize=True)
    base = int(rng.choice([0, 1000]))
    st.set_id_base(base)
    B, Cn = int(rng.integers(1, 5)), int(rng.choice([1, 31, 32, 33, 64, 100]))
    q = rng.standard_normal((B, Lq, dim)).astype(np.float32)
    cand = rng.integers(-2, ndocs + 2, size=(B, Cn)).astype(np.int64)
    q_len = rng.integers(1, Lq + 1, size=B).astype(np.int32) if rng.random() < 0.5 else None
    n_cand = rng.integers(0, Cn + 1, size=B).astype(np.int32) if rng.random() < 0.5 else None
    mode = int(rng.integers(0, 2))
    v2 = bool(rng.random() < 0.5)
    simt = bool(rng.random() < 0.2)
    os.environ["TS_S2_V2"] = "1" if v2 else "0"
    epi2 = bool(rng.random() < 0.5)
    os.environ["TS_S2_EPI2"] = "1" if epi2 else "0"
    got = st.maxsim_host(q, np.where((cand >= 0) & (cand < ndocs), cand + base, cand if base == 0 else -1), q_len=q_len,
                         n_cand=n_cand, mode=mode | (_lib.TS_S2_FORCE_SIMT if simt else 0))
    off = np.concatenate([[0], np.cumsum(lens)])
    nr = lambda x: flat_ip.round_to(maxsim.l2_normalize_tokens(x), dtype)      # noqa: E731
    ref = np.zeros((B, Cn), np.float32)
    for b in range(B):
        lq = Lq if q_len is None else int(q_len[b])
        for j in range(Cn if n_cand is None else int(n_cand[b])):
            c = int(cand[b, j])
            if 0 <= c < ndocs:
                ref[b, j] = maxsim.score(nr(q[b, :lq]), nr(tok[off[c]:off[c + 1]]), mode, normalize=False)
    ok = np.allclose(got, ref, rtol=1e-3, atol=3e-4)
    return (f"S2 {hw} dim={dim} Lq={Lq} ndocs={ndocs} {style} B={B} C={Cn} {dtype} mode={mode} v2={v2} epi2={epi2} simt={simt} "
            f"q_len={None if q_len is None else q_len.tolist()} n_cand={None if n_cand is None else n_cand.tolist()}"), \
        ([] if ok else [f"max abs err {np.abs(got - ref).max()}"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--only", default="", help="s1 | s2 | ivf | tf32: one kind of case only")
    args = ap.parse_args()
    load()
    rng = np.random.default_rng(args.seed)
    t0, n, fails = time.time(), 0, 0
    while time.time() - t0 < args.seconds:
        r = {"s1": 0.0, "s2": 0.5, "ivf": 0.7, "tf32": 0.9}.get(args.only, rng.random())
        desc, bad = (stage1_case if r < 0.4 else stage2_case if r < 0.65 else ivf_case if r < 0.85 else tf32_case)(rng)
        n += 1
        if bad:
            fails += 1
            print("FAIL", desc, bad[:3], flush=True)
    print(f"{n} cases, {fails} failures, seed {args.seed}")
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
