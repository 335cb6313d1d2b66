"""CPU: the HOST side of the C ABI (csrc/api.cu + csrc/shard_file.cu) compiled against a stand-in
CUDA runtime (tests/hostsim/: "device" memory is host memory, ingest launchers restated on the
CPU, every scan / scoring launcher refuses) -- argument checks, growth, staging, shard files and
the Python binding above them, without a GPU.  Test infrastructure only: the simulated library
lives in build/hostsim/, is loaded explicitly here, and is never the product path."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from test_shard_files import bf16_bits, bf16_to_f32

from oracle import flat_ip, maxsim
from tristage_rag_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM_DIR = os.path.join(ROOT, "build", "hostsim")


@pytest.fixture(scope="module")
def sim_built():
    env = dict(os.environ)
    env.pop("CXX", None)
    out = subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "hostsim")], env=env, stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True)
    assert out.returncode == 0, out.stdout
    return SIM_DIR


@pytest.fixture()
def sim(sim_built, monkeypatch):
    """The Python binding (_lib.Index / _lib.TokStore ...) running over the simulated library."""
    L = C.CDLL(os.path.join(sim_built, "libtristage_hostsim.so"))
    assert L.hostsim_is_simulation() == 1
    for name, (res, args) in _lib.SYMBOLS.items():
        if name.startswith(("ts_bm25_", "ts_hybrid_", "ts_ivf_")):
            continue                 # kernels + their host code live in one file: covered by tests/cudasim instead
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    monkeypatch.setattr(_lib, "_lib", L)
    monkeypatch.setattr(_lib, "_stream_ptr", lambda device: None)
    yield L
    assert C.c_longlong.in_dll(L, "hostsim_live_pinned").value == 0


def test_host_side_under_address_sanitizer(sim_built, tmp_path):
    """tests/hostsim/hostsim_main.cc: growth, staged adds, save / load / range append through the
    double-buffered copy loops (48 KB staging in this build), error paths, no leaked allocation."""
    out = subprocess.run([os.path.join(sim_built, "hostsim_asan"), str(tmp_path)], stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True, timeout=300)
    assert out.returncode == 0 and "hostsim ok" in out.stdout, out.stdout[-3000:]


def test_simulated_library_has_no_compute_path(sim):
    idx = _lib.Index(32, "bf16")
    idx.add(np.ones((4, 32), np.float32))
    with pytest.raises(_lib.TristageError, match="not simulated"):
        idx.search_host(np.ones((1, 32), np.float32), 2)
    st = _lib.TokStore(16, "bf16")
    st.add(np.ones((3, 16), np.float32), [3])
    with pytest.raises(_lib.TristageError, match="not simulated"):
        st.maxsim_host(np.ones((1, 2, 16), np.float32), np.zeros((1, 1), np.int64))


def test_index_ingest_bookkeeping_matches_the_oracle_rounding(sim, tmp_path):
    rng = np.random.default_rng(0)
    d = 100                                              # ld = 104: four zero pad columns per row
    X = rng.standard_normal((5000, d)).astype(np.float32)
    idx = _lib.Index(d, "bf16", "ip", 0)
    for part in np.array_split(X, 7):                    # 7 adds: capacity doubles 1024 -> 8192
        idx.add(part, normalize=True)
    assert idx.ntotal == 5000
    want = flat_ip.round_to(flat_ip.normalize_rows(X), "bf16")
    got = idx.get_rows(0, 5000)
    # x/(|x|+1e-8) in fp32 then RNE to bf16: the ingest formula of the reference (:285-288)
    assert np.abs(got - want).max() <= 2 ** -8 and (got == want).mean() > 0.999
    p = str(tmp_path / "i.tsshard")
    idx.set_id_base(77)
    idx.save(p)
    _lib.file_verify(p)
    fi = _lib.file_probe(p)
    assert (fi["n"], fi["dim"], fi["ld"], fi["id_base"]) == (5000, d, 104, 77)
    body = np.memmap(p, np.uint16, "r", fi["payload_offset"], (5000, 104))
    assert (body[:, d:] == 0).all() and (bf16_to_f32(np.array(body[:, :d])) == got).all()
    back = _lib.Index.load(p, 0)
    assert (back.ntotal, back.dim, back.dtype, back.metric) == (5000, d, _lib.TS_BF16, _lib.TS_METRIC_IP)
    assert (back.get_rows(0, 5000) == got).all()
    part = _lib.Index(d, "bf16", "ip", 0)
    part.append_file(p, 4000, 1000)
    part.append_file(p, 0, 10)
    assert (part.get_rows(0, 1010) == np.concatenate([got[4000:], got[:10]])).all()
    # a damaged file must not load through a PARTIAL range either (the re-sharding path): one flipped payload
    # byte far from the requested rows is found before anything is appended
    raw = bytearray(open(p, "rb").read())
    raw[fi["payload_offset"] + 104 * 2 * 2500 + 11] ^= 0x40
    bad = str(tmp_path / "damaged.tsshard")
    open(bad, "wb").write(bytes(raw))
    with pytest.raises(_lib.TristageError, match="checksum"):
        part.append_file(bad, 4000, 1000)
    assert part.ntotal == 1010
    idx.reset()
    assert idx.ntotal == 0
    with pytest.raises(_lib.TristageError):
        idx.get_rows(0, 1)


def test_tokstore_ingest_and_reshard_bookkeeping(sim, tmp_path):
    rng = np.random.default_rng(1)
    dim, ndocs = 40, 900
    lens = rng.integers(1, 257, size=ndocs)
    tok = rng.standard_normal((int(lens.sum()), dim)).astype(np.float32)
    st = _lib.TokStore(dim, "bf16", 0)
    cut = int(lens[:300].sum())
    st.add(tok[:cut], lens[:300], normalize=True)
    st.add(tok[cut:], lens[300:], normalize=True)
    assert (st.ndocs, st.ntokens) == (ndocs, int(lens.sum()))
    p = str(tmp_path / "t.tsshard")
    st.save(p)
    _lib.file_verify(p)
    fi = _lib.file_probe(p)
    pad = (lens + 7) // 8 * 8
    assert fi["nrows"] == int(pad.sum()) and fi["ntokens"] == int(lens.sum())
    off = np.memmap(p, np.int64, "r", fi["table_offset"], (ndocs,))
    body = np.memmap(p, np.uint16, "r", fi["payload_offset"], (fi["nrows"], dim))
    want = flat_ip.round_to(maxsim.l2_normalize_tokens(tok), "bf16")
    src = np.concatenate([[0], np.cumsum(lens)])
    for dno in (0, 1, 299, 300, 450, ndocs - 1):
        rows = bf16_to_f32(np.array(body[off[dno]: off[dno] + pad[dno]]))
        assert np.abs(rows[: lens[dno]] - want[src[dno]: src[dno + 1]]).max() <= 2 ** -8
        assert (rows[lens[dno]:] == 0).all()
    # the same docs written by the host writer from the stored bits: identical file
    unp = np.concatenate([np.array(body[off[i]: off[i] + lens[i]]) for i in range(ndocs)])
    p_host = str(tmp_path / "host.tsshard")
    _lib.write_tokstore_file(p_host, unp, lens, "bf16")
    assert open(p_host, "rb").read() == open(p, "rb").read()
    # three pieces appended to a fresh store == the original store
    st3 = _lib.TokStore(dim, "bf16", 0)
    for lo, n in ((0, 123), (123, 700), (823, 77)):
        st3.append_file(p, lo, n)
    p3 = str(tmp_path / "t3.tsshard")
    st3.save(p3)
    assert open(p3, "rb").read() == open(p, "rb").read()
    loaded = _lib.TokStore.load(p)
    assert (loaded.dim, loaded.dtype, loaded.ndocs) == (dim, _lib.TS_BF16, ndocs)
    with pytest.raises(ValueError):
        _lib.TokStore.load(p, dim=64)
    with pytest.raises(_lib.TristageError):
        st.add(np.ones((300, dim), np.float32), [300])       # more than TS_S2_MAX_LD tokens


def test_sharded_wrappers_save_and_reshard_over_the_binding(sim, tmp_path):
    """dist.ShardedIndex / ShardedTokStore save + load (world size 1) over the real binding."""
    from tristage_rag_b200 import dist as tdist

    rng = np.random.default_rng(2)
    N, d = 1003, 32
    X = flat_ip.normalize_rows(rng.standard_normal((N, d)).astype(np.float32)).astype(np.float32)
    directory = str(tmp_path / "corpus")
    os.makedirs(directory)
    for r in range(4):
        lo, hi = tdist.shard_range(N, r, 4)
        sh = _lib.Index(d, "bf16", "ip", 0)
        sh.add(X[lo:hi])
        sh.set_id_base(lo)
        sh.save(os.path.join(directory, tdist.shard_file_name("index", r, 4)))
    tdist.write_manifest(directory, "index", N, 4)
    one = tdist.ShardedIndex.load(directory, 0)
    assert one.local.ntotal == N
    assert (one.local.get_rows(0, N) == flat_ip.round_to(X, "bf16")).all()
    one.save(str(tmp_path / "again"))
    fi = _lib.file_probe(os.path.join(str(tmp_path / "again"), tdist.shard_file_name("index", 0, 1)))
    assert fi["n"] == N and fi["id_base"] == 0
    assert (bf16_bits(X) == np.memmap(os.path.join(str(tmp_path / "again"), tdist.shard_file_name("index", 0, 1)),
                                      np.uint16, "r", fi["payload_offset"], (N, d))).all()
