"""CPU: the shard-file format (include/tristage.h "shard files") and the re-sharding plan of
dist.py.  Everything here runs without a GPU: ts_file_probe / ts_file_verify /
ts_file_write_*_host are host-only entry points of the C ABI, and the per-rank index in the
world-size-2 gloo test is a file-backed stand-in scored by the oracle (test infrastructure)."""
import os
import socket
import struct

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import flat_ip
from tristage_rag_b200 import _lib
from tristage_rag_b200 import dist as tdist

M64 = (1 << 64) - 1
P1, P2, P3, P4, P5 = (11400714785074694791, 14029467366897019727, 1609587929392839161, 9650029242287828579,
                      2870177450012600261)


def _rotl(x, r):
    return ((x << r) | (x >> (64 - r))) & M64


def _round(acc, inp):
    return (_rotl((acc + inp * P2) & M64, 31) * P1) & M64


def _merge(h, v):
    return (((h ^ _round(0, v)) * P1) + P4) & M64


def xxh64(data: bytes, seed: int = 0) -> int:
    """Independent pure-Python XXH64 (the published algorithm) to pin the C implementation."""
    n, i = len(data), 0
    if n >= 32:
        v = [(seed + P1 + P2) & M64, (seed + P2) & M64, seed, (seed - P1) & M64]
        while i + 32 <= n:
            for j in range(4):
                v[j] = _round(v[j], struct.unpack_from("<Q", data, i + 8 * j)[0])
            i += 32
        h = (_rotl(v[0], 1) + _rotl(v[1], 7) + _rotl(v[2], 12) + _rotl(v[3], 18)) & M64
        for j in range(4):
            h = _merge(h, v[j])
    else:
        h = (seed + P5) & M64
    h = (h + n) & M64
    while i + 8 <= n:
        h = (_rotl(h ^ _round(0, struct.unpack_from("<Q", data, i)[0]), 27) * P1 + P4) & M64
        i += 8
    if i + 4 <= n:
        h = (_rotl(h ^ (struct.unpack_from("<I", data, i)[0] * P1 & M64), 23) * P2 + P3) & M64
        i += 4
    while i < n:
        h = (_rotl(h ^ (data[i] * P5 & M64), 11) * P1) & M64
        i += 1
    h ^= h >> 33
    h = (h * P2) & M64
    h ^= h >> 29
    h = (h * P3) & M64
    h ^= h >> 32
    return h


def bf16_bits(x: np.ndarray) -> np.ndarray:
    """float32 -> bf16 bit patterns (round to nearest even, like the device cast)."""
    r = flat_ip.round_to(np.ascontiguousarray(x, np.float32), "bf16")
    return (r.view(np.uint32) >> 16).astype(np.uint16)


def bf16_to_f32(u: np.ndarray) -> np.ndarray:
    return (u.astype(np.uint32) << 16).view(np.float32)


def test_xxh64_known_answers():
    assert xxh64(b"") == 0xEF46DB3751D8E999                      # published test vector
    assert xxh64(b"", 1) != xxh64(b"")
    rng = np.random.default_rng(0)
    for n in (1, 3, 4, 7, 8, 31, 32, 33, 63, 64, 100, 1000):
        b = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
        assert xxh64(b) == xxh64(bytes(b)) and xxh64(b) != xxh64(b + b"\0")
    assert len({xxh64(bytes([i])) for i in range(256)}) == 256


@pytest.mark.parametrize("n,dim", [(0, 64), (1, 8), (1000, 60), (4099, 768)])
def test_index_file_layout_hash_and_probe(tmp_path, n, dim):
    rng = np.random.default_rng(n + dim)
    ld = (dim + 7) // 8 * 8
    rows = np.zeros((n, ld), np.uint16)
    rows[:, :dim] = bf16_bits(rng.standard_normal((n, dim)))
    p = str(tmp_path / "x.tsshard")
    _lib.write_index_file(p, rows, "bf16", id_base=12345, dim=dim)
    assert not os.path.exists(p + ".tmp")
    fi = _lib.file_probe(p)
    assert (fi["kind"], fi["version"], fi["dim"], fi["ld"], fi["dtype"], fi["metric"]) == (1, 2, dim, ld, _lib.TS_BF16, 0)
    assert (fi["n"], fi["nrows"], fi["id_base"]) == (n, n, 12345)
    assert fi["payload_offset"] % 4096 == 0 and fi["payload_bytes"] == n * ld * 2 and fi["table_bytes"] == 0
    raw = open(p, "rb").read()
    assert raw[:8] == b"TSSHARD2"
    payload = raw[fi["payload_offset"]: fi["payload_offset"] + fi["payload_bytes"]]
    assert payload == rows.tobytes()
    assert fi["payload_hash"] == xxh64(payload) and fi["table_hash"] == xxh64(b"")
    _lib.file_verify(p)


def test_cosine_index_file_carries_inverse_norms(tmp_path):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((257, 16)).astype(np.float32)
    inv = (1.0 / (np.linalg.norm(x, axis=1) + 1e-8)).astype(np.float32)
    p = str(tmp_path / "c.tsshard")
    _lib.write_index_file(p, x, "fp32", metric="cosine", inv_norm=inv)
    fi = _lib.file_probe(p)
    assert fi["metric"] == _lib.TS_METRIC_COSINE and fi["table_bytes"] == 257 * 4 and fi["dtype"] == _lib.TS_F32
    raw = open(p, "rb").read()
    assert raw[fi["table_offset"]: fi["table_offset"] + fi["table_bytes"]] == inv.tobytes()
    assert fi["table_hash"] == xxh64(inv.tobytes())
    _lib.file_verify(p)
    with pytest.raises(_lib.TristageError):                       # cosine without norms is refused
        _lib.write_index_file(str(tmp_path / "bad"), x, "fp32", metric="cosine")


def test_tokstore_file_pads_docs_to_8_rows(tmp_path):
    rng = np.random.default_rng(4)
    dim = 32
    lens = np.array([1, 8, 9, 16, 255, 256, 3], np.int32)
    tok = bf16_bits(rng.standard_normal((int(lens.sum()), dim)))
    p = str(tmp_path / "t.tsshard")
    _lib.write_tokstore_file(p, tok, lens, "bf16", id_base=40)
    fi = _lib.file_probe(p)
    pad = (lens + 7) // 8 * 8
    assert (fi["kind"], fi["n"], fi["ntokens"], fi["nrows"], fi["id_base"]) == (2, 7, int(lens.sum()), int(pad.sum()), 40)
    raw = open(p, "rb").read()
    t0 = fi["table_offset"]
    off = np.frombuffer(raw, np.int64, 7, t0)
    ln = np.frombuffer(raw, np.int32, 7, t0 + 7 * 8)
    assert off.tolist() == np.concatenate([[0], np.cumsum(pad)[:-1]]).tolist() and ln.tolist() == lens.tolist()
    assert fi["table_hash"] == xxh64(raw[t0:t0 + fi["table_bytes"]])     # 84 bytes: exercises the 8/4-byte tails
    body = np.frombuffer(raw, np.uint16, fi["nrows"] * dim, fi["payload_offset"]).reshape(-1, dim)
    src = 0
    for o, L, P in zip(off, lens, pad):
        assert (body[o:o + L] == tok[src:src + L]).all() and (body[o + L:o + P] == 0).all()
        src += L
    _lib.file_verify(p)
    import ctypes as C
    for bad in (0, 257):                                          # token counts outside 1..256 are refused
        ln = np.array([bad], np.int32)
        rc = _lib.lib().ts_file_write_tokstore_host(str(tmp_path / "bad").encode(), dim, _lib.TS_BF16, 1, 0,
                                                    C.c_void_p(ln.ctypes.data), C.c_void_p(tok.ctypes.data))
        assert rc == -1 and not os.path.exists(str(tmp_path / "bad"))


def test_corruption_truncation_and_foreign_files_are_detected(tmp_path):
    rows = bf16_bits(np.random.default_rng(5).standard_normal((3000, 64)))
    p = str(tmp_path / "x.tsshard")
    _lib.write_index_file(p, rows, "bf16")
    good = open(p, "rb").read()
    fi = _lib.file_probe(p)

    def write(name, data):
        q = str(tmp_path / name)
        open(q, "wb").write(data)
        return q

    flipped = bytearray(good)
    flipped[fi["payload_offset"] + 777] ^= 0x10
    q = write("flipped", bytes(flipped))
    _lib.file_probe(q)                                            # header is fine ...
    with pytest.raises(_lib.TristageError, match="checksum"):     # ... the payload is not
        _lib.file_verify(q)
    with pytest.raises(_lib.TristageError, match="truncated"):
        _lib.file_probe(write("short", good[:-100]))
    with pytest.raises(_lib.TristageError, match="not a tristage shard"):
        _lib.file_probe(write("foreign", b"FAISSIDX" + good[8:]))
    with pytest.raises(_lib.TristageError, match="not a tristage shard"):
        _lib.file_probe(write("tiny", b"abc"))
    with pytest.raises(_lib.TristageError):
        _lib.file_probe(str(tmp_path / "missing"))
    hdr = bytearray(good)
    struct.pack_into("<q", hdr, 32, 2999)                         # n no longer matches the payload
    with pytest.raises(_lib.TristageError):
        _lib.file_probe(write("lied", bytes(hdr)))


def test_plan_reshard_covers_every_range_exactly():
    for n in (0, 1, 7, 41, 1000, 10_000_000):
        for w_save in (1, 2, 3, 8):
            shards = [{"file": f"f{r}", "lo": tdist.shard_range(n, r, w_save)[0], "hi": tdist.shard_range(n, r, w_save)[1]}
                      for r in range(w_save)]
            for w_load in (1, 2, 4, 5, 8):
                seen = 0
                for r in range(w_load):
                    lo, hi = tdist.shard_range(n, r, w_load)
                    at = lo
                    for f, first, cnt in tdist.plan_reshard(shards, lo, hi):
                        sh = shards[int(f[1:])]
                        assert cnt > 0 and sh["lo"] + first == at and first + cnt <= sh["hi"] - sh["lo"]
                        at += cnt
                    assert at == hi
                    seen += hi - lo
                assert seen == n
    with pytest.raises(ValueError):
        tdist.plan_reshard([{"file": "a", "lo": 0, "hi": 5}, {"file": "b", "lo": 6, "hi": 9}], 0, 9)
    with pytest.raises(ValueError):
        tdist.plan_reshard([{"file": "a", "lo": 0, "hi": 5}], 0, 9)


# ----------------------------------------------------------- gloo, world size 2 --
class _FileLocalIndex:
    """File-backed stand-in for _lib.Index on the CPU: bf16 rows kept as bit patterns, saved with the
    host writer, re-loaded by reading the payload at the probed offsets, scored by the oracle."""

    def __init__(self, dim):
        self.dim, self.bits, self.base = dim, np.zeros((0, dim), np.uint16), 0

    def set_id_base(self, b):
        self.base = b

    def add_bits(self, bits):
        self.bits = np.concatenate([self.bits, bits])

    def save(self, path):
        _lib.write_index_file(path, self.bits, "bf16", id_base=self.base, dim=self.dim)

    def append_file(self, path, first, n):
        fi = _lib.file_probe(path)
        assert fi["dim"] == self.dim and first + n <= fi["n"]
        body = np.memmap(path, np.uint16, "r", fi["payload_offset"], (fi["n"], fi["ld"]))
        self.add_bits(np.array(body[first:first + n, : self.dim]))

    def search(self, q, k):
        D, I = flat_ip.topk_desc(q.numpy() @ bf16_to_f32(self.bits).T, k)
        return torch.from_numpy(D), torch.from_numpy(np.where(I >= 0, I + self.base, -1))


def _numpy_merge(all_s, all_i):
    G, B, k = all_s.shape
    s = all_s.permute(1, 0, 2).reshape(B, G * k).numpy()
    i = all_i.permute(1, 0, 2).reshape(B, G * k).numpy()
    outD = np.full((B, k), flat_ip.LOWEST_F32, np.float32)
    outI = np.full((B, k), -1, np.int64)
    for b in range(B):
        ok = i[b] >= 0
        order = np.lexsort((i[b][ok], -s[b][ok].astype(np.float64)))[:k]
        outD[b, : len(order)] = s[b][ok][order]
        outI[b, : len(order)] = i[b][ok][order]
    return torch.from_numpy(outD), torch.from_numpy(outI)


def _corpus(n_total, dim=24):
    rng = np.random.default_rng(11)
    X = bf16_bits(flat_ip.normalize_rows(rng.standard_normal((n_total, dim)).astype(np.float32)))
    Q = flat_ip.normalize_rows(rng.standard_normal((3, dim)).astype(np.float32)).astype(np.float32)
    return X, Q


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, directory, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        X, Q = _corpus(n_total)
        make_local = lambda info, reserve: _FileLocalIndex(info["dim"])      # noqa: E731
        # load the 3-rank save under world size 2, search, save again as a 2-rank corpus
        idx = tdist.ShardedIndex.load(os.path.join(directory, "w3"), group=None, make_local=make_local,
                                      merge_fn=_numpy_merge)
        lo, hi = tdist.shard_range(n_total, rank, world)
        assert (idx.local.bits == X[lo:hi]).all() and idx.local.base == lo
        D, I = idx.search(torch.from_numpy(Q), 9)
        rD, rI = flat_ip.topk_desc(Q @ bf16_to_f32(X).T, 9)
        assert (I.numpy() == rI).all() and np.allclose(D.numpy(), rD)
        idx.save(os.path.join(directory, "w2"))
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_sharded_save_load_reshards_across_world_sizes(tmp_path):
    n_total = 103
    X, Q = _corpus(n_total)
    d3 = str(tmp_path / "w3")
    os.makedirs(d3)
    for r in range(3):                                             # what three ranks would have written
        lo, hi = tdist.shard_range(n_total, r, 3)
        _lib.write_index_file(os.path.join(d3, tdist.shard_file_name("index", r, 3)), X[lo:hi], "bf16", id_base=lo)
    tdist.write_manifest(d3, "index", n_total, 3)
    world, port = 2, _free_port()
    ret = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, port, n_total, str(tmp_path), ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))
    # the 2-rank save the workers wrote loads again on ONE process (world size 1) bit for bit
    man = tdist.read_manifest(str(tmp_path / "w2"), "index")
    assert man["world_size"] == 2 and [s["lo"] for s in man["shards"]] == [0, 52]
    for s in man["shards"]:
        _lib.file_verify(os.path.join(str(tmp_path / "w2"), s["file"]))
        assert _lib.file_probe(os.path.join(str(tmp_path / "w2"), s["file"]))["id_base"] == s["lo"]
    one = tdist.ShardedIndex.load(str(tmp_path / "w2"), make_local=lambda info, reserve: _FileLocalIndex(info["dim"]),
                                  merge_fn=_numpy_merge)
    assert (one.local.bits == X).all() and (one.lo, one.hi) == (0, n_total)
    with pytest.raises(FileNotFoundError):
        tdist.read_manifest(str(tmp_path / "w2"), "tokstore")
