"""The fused multi-GPU data planes on the CPU emulator with every rank in its OWN process: the receive buffers are
POSIX shared memory mapped into all of them (what NVLink peer mappings are on the hardware), so a kernel that pushes its
rows and then WAITS for its peers' rows runs against real concurrent peers.

  * Stage 1: ts_index_search_sharded with the whole exchange in ONE kernel (select + push + wait + merge; the emulator
    build fuses only under TS_SIM_XFUSE=1) == the two-kernel form == one index over all rows, several steps (both
    parities), exact ties across shards, also under adversarial thread timing (CUDASIM_ASYNC);
  * Stage 2: ts_maxsim_scatter + ts_exchange_wait_take == ts_maxsim over the whole store.

Test infrastructure only (tests/cudasim); the hardware check is tools/dist_check.py."""
import ctypes as C
import os
import sys
from multiprocessing import shared_memory

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def p(a):
    return C.c_void_p(a.ctypes.data)


def _worker(rank, G, names, sizes, seed, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["CUDASIM_EXTERNAL_WAITS"] = "1"
    os.environ["TS_SIM_XFUSE"] = "1"
    if seed:
        os.environ["CUDASIM_ASYNC"] = str(seed + rank)
    for var in ("TS_PAIR", "TS_FUSE", "TS_TF32"):
        os.environ[var] = "0"
    import conftest
    from oracle import flat_ip
    from tristage_rag_b200 import _lib

    conftest._enter_emulation()
    sim = _lib.lib()
    shms = [shared_memory.SharedMemory(name=n) for n in names]
    try:
        # ---------------- Stage 1 ----------------
        B, k, B_max, k_max = 5, 40, 8, 64
        x_bytes = sizes[0]
        bufs = [np.ndarray((x_bytes,), np.uint8, buffer=shms[r].buf) for r in range(G)]
        bases = np.array([b.ctypes.data for b in bufs], np.int64)
        assert all(int(a) % 16 == 0 for a in bases)
        rng = np.random.default_rng(4)                       # the same data in every process
        N, d = 900, 32
        X = flat_ip.normalize_rows(rng.standard_normal((N, d)).astype(np.float32)).astype(np.float32)
        X[N // 2] = X[3]
        X[N - 7] = X[3]                                      # exact ties across shards
        lo, hi = rank * N // G, (rank + 1) * N // G
        mine = _lib.Index(d, "bf16", "ip", 0)
        mine.add(X[lo:hi])
        mine.set_id_base(lo)
        full = _lib.Index(d, "bf16", "ip", 0)
        full.add(X)
        x = _lib.Exchange(0, rank, G, bases, B_max, k_max)
        sim.cudasim_launches.restype = C.c_ulonglong
        per_step = {}
        for step in range(4):
            Q = flat_ip.normalize_rows(rng.standard_normal((B, d)).astype(np.float32)).astype(np.float32)
            os.environ["TS_XFUSE"] = "1" if step != 2 else "0"          # step 2: the two-kernel form, same buffers
            out_s, out_i = np.empty((B, k), np.float32), np.empty((B, k), np.int64)
            l0 = sim.cudasim_launches()
            _lib.check(sim.ts_index_search_sharded(mine._h, x._h, p(Q), _lib.TS_F32, B, k, 0, _lib.PATHS["auto"], p(out_s), p(out_i), None))
            per_step[step] = sim.cudasim_launches() - l0
            Df, If = full.search_host(Q, k)
            assert (out_i == If).all() and (out_s == Df).all(), (rank, step)
        assert per_step[2] == per_step[1] + 1 == per_step[3] + 1, per_step      # the fused form really is one launch fewer
        # ---------------- Stage 2 ----------------
        Bq, Cn, dim, Lq = 4, 37, 32, 8
        n = Bq * Cn
        slot = (n * 4 + 15) // 16 * 16
        flags_off = 2 * slot
        sbufs = [np.ndarray((sizes[1],), np.uint8, buffer=shms[G + r].buf) for r in range(G)]
        sbases = np.array([b.ctypes.data for b in sbufs], np.int64)
        lens = rng.integers(2, 60, size=90)
        tok = rng.standard_normal((int(lens.sum()), dim)).astype(np.float32)
        off = np.concatenate([[0], np.cumsum(lens)])
        dlo, dhi = rank * 90 // G, (rank + 1) * 90 // G
        st = _lib.TokStore(dim, "bf16", 0)
        st.add(tok[off[dlo]:off[dhi]], lens[dlo:dhi], normalize=True)
        st.set_id_base(dlo)
        whole = _lib.TokStore(dim, "bf16", 0)
        whole.add(tok, lens, normalize=True)
        for step in range(4):
            q = rng.standard_normal((Bq, Lq, dim)).astype(np.float32)
            cand = rng.integers(-2, 95, size=(Bq, Cn)).astype(np.int64)
            parity, seq = step & 1, step + 1
            mat_off, f_off = parity * slot, flags_off + parity * G * 4
            _lib.check(sim.ts_maxsim_scatter(st._h, p(q), _lib.TS_F32, None, Bq, Lq, p(cand), None, Cn, 0, _lib.TS_FLAG_NORMALIZE_Q,
                                             p(sbases), G, rank, mat_off, f_off, seq, None))
            out = np.empty((Bq, Cn), np.float32)
            base = sbufs[rank].ctypes.data
            _lib.check(sim.ts_exchange_wait_take(0, C.c_void_p(base + mat_off), C.c_void_p(base + f_off), G, seq, n, p(out), None))
            assert np.array_equal(out, whole.maxsim_host(q, cand)), (rank, step)
        ret[rank] = True
    finally:
        del bufs, sbufs
        for s in shms:
            s.close()


@pytest.mark.parametrize("G,seed", [(2, 0), (3, 0), (2, 11)])
def test_fused_exchange_with_every_rank_in_its_own_process(G, seed):
    sys.path.insert(0, ROOT)
    from tristage_rag_b200 import _lib

    _lib.lib()                                                # the C library gives the buffer size (no GPU call)
    x_bytes = int(_lib.lib().ts_exchange_buffer_bytes(G, 8, 64))
    s_bytes = 2 * ((4 * 37 * 4 + 15) // 16 * 16) + 2 * G * 4 + 16
    try:
        shms = [shared_memory.SharedMemory(create=True, size=x_bytes) for _ in range(G)] + \
               [shared_memory.SharedMemory(create=True, size=s_bytes) for _ in range(G)]
    except OSError as e:                                      # no usable /dev/shm on this box
        pytest.skip(f"POSIX shared memory unavailable: {e}")
    try:
        for s in shms:
            np.ndarray((s.size,), np.uint8, buffer=s.buf)[:] = 0
        ret = mp.get_context("spawn").Manager().dict()
        mp.spawn(_worker, args=(G, [s.name for s in shms], (x_bytes, s_bytes), seed, ret), nprocs=G, join=True)
        assert all(ret.get(r) for r in range(G))
    finally:
        for s in shms:
            s.close()
            s.unlink()
