"""The multi-GPU exchange kernels on ONE GPU (world size 1: the "peer" buffers are this GPU's own memory), so the
round-end `pytest -m gpu` run covers them on hardware; `tools/dist_check.py` is the real multi-GPU check (world sizes 2
and 8, profiles/README.md).  Stage 1: the whole exchange in one kernel (select + push + wait + merge) and the two-kernel
form return the local search bit for bit.  Stage 2: the scatter fused into the scoring kernel + wait-take returns
ts_maxsim's matrix bit for bit and leaves the receive buffer zeroed."""
import ctypes as C

import numpy as np
import pytest

from oracle import flat_ip

pytestmark = pytest.mark.gpu


def test_stage1_exchange_with_one_rank_equals_the_local_search(cuda_device, monkeypatch):
    import torch

    from tristage_rag_b200 import _lib

    rng = np.random.default_rng(3)
    N, d, k = 60_000, 256, 100
    X = flat_ip.normalize_rows(rng.standard_normal((N, d)).astype(np.float32)).astype(np.float32)
    idx = _lib.Index(d, "bf16", "ip", 0)
    idx.add(X)
    idx.set_id_base(1000)
    dev = torch.device("cuda", 0)
    buf = torch.zeros(_lib.Exchange.buffer_bytes(1, 1024, 128), dtype=torch.uint8, device=dev)
    x = _lib.Exchange(0, 0, 1, [buf.data_ptr()], 1024, 128)
    for B in (1, 32, 200):
        q = torch.from_numpy(flat_ip.normalize_rows(rng.standard_normal((B, d)).astype(np.float32)).astype(np.float32)).to(dev)
        s0, i0 = idx.search(q, k)
        for fuse in ("1", "0", "1"):                       # both parities of the buffer get used
            monkeypatch.setenv("TS_XFUSE", fuse)
            s1, i1 = x.search(idx, q, k)
            torch.cuda.synchronize()
            assert torch.equal(i1, i0) and torch.equal(s1, s0), (B, fuse)
        D, I = x.search_host(idx, q.cpu().numpy(), k)
        assert (I == i0.cpu().numpy()).all() and (D == s0.cpu().numpy()).all()


@pytest.mark.parametrize("mode", [0, 1])
def test_stage2_scatter_with_one_rank_equals_maxsim(cuda_device, mode):
    import torch

    from tristage_rag_b200 import _lib

    rng = np.random.default_rng(4)
    dim, ndocs, B, Cn, Lq = 128, 3000, 6, 250, 32
    lens = rng.integers(16, 181, size=ndocs)
    tok = rng.standard_normal((int(lens.sum()), dim)).astype(np.float32)
    st = _lib.TokStore(dim, "bf16", 0)
    st.add(tok, lens, normalize=True)
    assert st.layout == 1
    dev = torch.device("cuda", 0)
    n = B * Cn
    slot = (n * 4 + 15) // 16 * 16
    flags_off = 2 * slot
    buf = torch.zeros(flags_off + 2 * 4 + 16, dtype=torch.uint8, device=dev)
    bases = torch.tensor([buf.data_ptr()], dtype=torch.int64, device=dev)
    for step in range(4):
        q = torch.from_numpy(rng.standard_normal((B, Lq, dim)).astype(np.float32)).to(dev)
        cand = torch.from_numpy(rng.integers(-2, ndocs + 3, size=(B, Cn)).astype(np.int64)).to(dev)
        n_cand = torch.from_numpy(rng.integers(Cn // 2, Cn + 1, size=B).astype(np.int32)).to(dev)
        ref = st.maxsim(q, cand, n_cand=n_cand, mode=mode)
        parity, seq = step & 1, step + 1
        mat_off, f_off = parity * slot, flags_off + parity * 4
        st.maxsim_scatter(q, cand, bases, 1, 0, mat_off, f_off, seq, n_cand=n_cand, mode=mode)
        out = torch.empty((B, Cn), dtype=torch.float32, device=dev)
        _lib.check(_lib.lib().ts_exchange_wait_take(0, C.c_void_p(buf.data_ptr() + mat_off), C.c_void_p(buf.data_ptr() + f_off), 1, seq,
                                                    n, C.c_void_p(out.data_ptr()), _lib._stream_ptr(0)))
        torch.cuda.synchronize()
        assert torch.equal(out, ref), (step, float((out - ref).abs().max()))
        assert not bool(buf[mat_off:mat_off + n * 4].any())
