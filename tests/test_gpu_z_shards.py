"""GPU: shard files through the C ABI (ts_index_save/load/append_file, ts_tokstore_*) -- a saved
shard reloads bit for bit, a host-written file equals a device-written one, any row/doc range
can be appended, a 3-shard save reloads on one GPU (re-sharding), corruption is refused."""
import os

import numpy as np
import pytest

from test_gpu_stage1 import make
from test_shard_files import bf16_bits

from tristage_rag_b200 import _lib
from tristage_rag_b200 import dist as tdist

pytestmark = pytest.mark.gpu


def _payload(path):
    fi = _lib.file_probe(path)
    with open(path, "rb") as f:
        f.seek(fi["payload_offset"])
        return fi, f.read(fi["payload_bytes"])


def test_index_shard_roundtrip_and_host_written_file(cuda_device, tmp_path):
    N, d, B, k = 30000, 200, 8, 50
    X, Q = make(N, d, B, seed=1, planted=10)
    idx = _lib.Index(d, "bf16", "ip", cuda_device)
    idx.add(X)
    idx.set_id_base(1000)
    D, I = idx.search_host(Q, k)
    p = str(tmp_path / "a.tsshard")
    idx.save(p)
    _lib.file_verify(p)
    fi, body = _payload(p)
    assert (fi["kind"], fi["n"], fi["dim"], fi["ld"], fi["id_base"], fi["dtype"]) == (1, N, d, 200, 1000, _lib.TS_BF16)
    # the device cast rounds like the oracle (RNE): the stored corpus is bit-identical to it,
    # so a file written on the host from the same values is the same file
    bits = bf16_bits(X)
    assert body == bits.tobytes()
    p_host = str(tmp_path / "host.tsshard")
    _lib.write_index_file(p_host, bits, "bf16", id_base=1000, dim=d)
    assert open(p_host, "rb").read() == open(p, "rb").read()
    for path in (p, p_host):
        idx2 = _lib.Index.load(path, cuda_device)
        assert (idx2.ntotal, idx2.dim, idx2.dtype, idx2.metric) == (N, d, _lib.TS_BF16, _lib.TS_METRIC_IP)
        D2, I2 = idx2.search_host(Q, k)
        assert (I2 == I).all() and (D2 == D).all() and I.min() >= 1000
    # arbitrary row ranges append in order (the re-sharding primitive)
    part = _lib.Index(d, "bf16", "ip", cuda_device)
    part.append_file(p, 10000, 5000)
    part.append_file(p, 0, 100)
    part.append_file(p, 100, 0)
    assert part.ntotal == 5100
    got = part.get_rows(0, 5100)
    ref = idx.get_rows(0, N)
    assert (got[:5000] == ref[10000:15000]).all() and (got[5000:] == ref[:100]).all()
    with pytest.raises(_lib.TristageError):
        part.append_file(p, N - 10, 11)                            # past the end of the file
    other = _lib.Index(d + 8, "bf16", "ip", cuda_device)
    with pytest.raises(_lib.TristageError):
        other.append_file(p, 0, 10)                                # dim mismatch
    # a flipped payload bit is refused at load
    raw = bytearray(open(p, "rb").read())
    raw[fi["payload_offset"] + 12345] ^= 4
    bad = str(tmp_path / "bad.tsshard")
    open(bad, "wb").write(bytes(raw))
    with pytest.raises(_lib.TristageError, match="checksum"):
        _lib.Index.load(bad, cuda_device)


def test_cosine_index_shard_roundtrip(cuda_device, tmp_path):
    rng = np.random.default_rng(2)
    N, d, k = 5000, 64, 20
    X = (rng.standard_normal((N, d)) * rng.uniform(0.1, 5.0, size=(N, 1))).astype(np.float32)
    Q = rng.standard_normal((4, d)).astype(np.float32)
    idx = _lib.Index(d, "fp16", "cosine", cuda_device)
    idx.add(X)
    D, I = idx.search_host(Q, k, normalize_q=True)
    p = str(tmp_path / "c.tsshard")
    idx.save(p)
    _lib.file_verify(p)
    assert _lib.file_probe(p)["table_bytes"] == N * 4
    idx2 = _lib.Index.load(p, cuda_device)
    assert idx2.metric == _lib.TS_METRIC_COSINE and idx2.dtype == _lib.TS_F16
    D2, I2 = idx2.search_host(Q, k, normalize_q=True)
    assert (I2 == I).all() and (D2 == D).all()
    half = _lib.Index(d, "fp16", "cosine", cuda_device)
    half.append_file(p, 2500, 2500)                                # inverse norms travel with the rows
    half.set_id_base(2500)
    D3, I3 = half.search_host(Q, k, normalize_q=True)
    keep = [[(s, i) for s, i in zip(D[b], I[b]) if i >= 2500] for b in range(4)]
    for b in range(4):
        n = len(keep[b])
        assert [i for _, i in keep[b]] == I3[b, :n].tolist() and [s for s, _ in keep[b]] == D3[b, :n].tolist()


def test_three_shard_save_reloads_on_one_gpu(cuda_device, tmp_path):
    """What three ranks save, one rank loads: ShardedIndex.load re-partitions by row range."""
    N, d, B, k = 20001, 128, 16, 100
    X, Q = make(N, d, B, seed=9, planted=20)
    full = _lib.Index(d, "bf16", "ip", cuda_device)
    full.add(X)
    D, I = full.search_host(Q, k)
    directory = str(tmp_path / "corpus")
    os.makedirs(directory)
    for r in range(3):
        lo, hi = tdist.shard_range(N, r, 3)
        sh = _lib.Index(d, "bf16", "ip", cuda_device)
        sh.add(X[lo:hi])
        sh.set_id_base(lo)
        sh.save(os.path.join(directory, tdist.shard_file_name("index", r, 3)))
    tdist.write_manifest(directory, "index", N, 3)
    one = tdist.ShardedIndex.load(directory, cuda_device)          # no process group: world size 1
    assert one.local.ntotal == N and (one.lo, one.hi) == (0, N)
    D2, I2 = one.local.search_host(Q, k)
    assert (I2 == I).all() and (D2 == D).all()
    one.save(str(tmp_path / "again"))                              # and back out as a 1-shard corpus
    man = tdist.read_manifest(str(tmp_path / "again"), "index")
    assert man["world_size"] == 1 and man["shards"][0]["hi"] == N
    _lib.file_verify(os.path.join(str(tmp_path / "again"), man["shards"][0]["file"]))


def test_tokstore_shard_roundtrip_ranges_and_resharding(cuda_device, tmp_path):
    rng = np.random.default_rng(8)
    dim, ndocs = 128, 600
    lens = rng.integers(1, 200, size=ndocs)
    tok = rng.standard_normal((int(lens.sum()), dim)).astype(np.float32)
    st = _lib.TokStore(dim, "bf16", cuda_device)
    st.add(tok, lens, normalize=True)
    q = rng.standard_normal((3, 32, dim)).astype(np.float32)
    cand = rng.integers(0, ndocs, size=(3, 64)).astype(np.int64)
    ref = st.maxsim_host(q, cand)
    p = str(tmp_path / "tok.tsshard")
    st.save(p)
    _lib.file_verify(p)
    fi = _lib.file_probe(p)
    assert (fi["kind"], fi["n"], fi["ntokens"], fi["dim"]) == (2, ndocs, int(lens.sum()), dim)
    st2 = _lib.TokStore.load(p, device=cuda_device)
    assert (st2.ndocs, st2.ntokens, st2.dim, st2.dtype) == (ndocs, int(lens.sum()), dim, _lib.TS_BF16)
    assert np.array_equal(st2.maxsim_host(q, cand), ref)
    # docs [200, 450) only: scores of owned candidates are unchanged, the rest score 0
    part = _lib.TokStore(dim, "bf16", cuda_device)
    part.append_file(p, 200, 250)
    part.set_id_base(200)
    got = part.maxsim_host(q, cand)
    own = (cand >= 200) & (cand < 450)
    assert own.any() and np.array_equal(got[own], ref[own]) and (got[~own] == 0).all()
    # two saved token shards -> one store (ShardedTokStore.load, world size 1)
    directory = str(tmp_path / "tokens")
    os.makedirs(directory)
    off = np.concatenate([[0], np.cumsum(lens)])
    for r in range(2):
        lo, hi = tdist.shard_range(ndocs, r, 2)
        sh = _lib.TokStore(dim, "bf16", cuda_device)
        sh.add(tok[off[lo]:off[hi]], lens[lo:hi], normalize=True)
        sh.set_id_base(lo)
        sh.save(os.path.join(directory, tdist.shard_file_name("tokstore", r, 2)))
    tdist.write_manifest(directory, "tokstore", ndocs, 2)
    merged = tdist.ShardedTokStore.load(directory, cuda_device)
    assert merged.local.ndocs == ndocs and merged.local.ntokens == int(lens.sum())
    assert np.array_equal(merged.local.maxsim_host(q, cand), ref)
    with pytest.raises(ValueError):
        _lib.TokStore.load(p, dim=64, device=cuda_device)           # caller's expectation is checked
