"""CPU: the index arithmetic of the Stage-2 epilogue (tests/s2_epilogue_model.py mirrors
csrc/s2_maxsim.cu) -- packing, per-quarter drain, masks of the padded last unit, quarter ranges
in finalize -- equals a direct max over every doc's columns, for the validated layout and for
the opt-in V2 epilogue, on ragged, tiny, maximal and adversarial doc-length mixes."""
import zlib

import numpy as np
import pytest

from s2_epilogue_model import TILE_N, direct, epilogue_v1, epilogue_v2, pack_tiles

CASES = {
    "config4_mix": lambda rng: rng.integers(16, 181, size=300),
    "short_docs": lambda rng: rng.integers(1, 41, size=400),
    "one_token_docs": lambda rng: np.ones(100, np.int64),
    "eight_token_docs": lambda rng: np.full(70, 8),
    "max_docs": lambda rng: np.full(9, 256),
    "exact_quarters": lambda rng: np.array([64, 64, 64, 64, 128, 128, 192, 64, 56, 8, 8, 57, 63, 65, 1]),
    "any_length": lambda rng: rng.integers(1, 257, size=300),
    "straddlers": lambda rng: np.array([60, 10, 120, 7, 59, 130, 3, 3, 3, 250, 5, 1, 63, 1, 64, 1, 127, 129]),
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("lq", [1, 7, 32, 33, 64, 100, 128])
def test_epilogue_models_equal_direct_max(name, lq):
    rng = np.random.default_rng(zlib.crc32(f"{name}-{lq}".encode()))
    lens = [int(x) for x in CASES[name](rng)]
    tiles = pack_tiles(lens)
    assert sum(len(t["docs"]) for t in tiles) == len(lens)
    for t in tiles:
        assert 0 < t["used"] <= TILE_N and t["used"] % 8 == 0 and len(t["docs"]) <= 32
        S = rng.standard_normal((128, TILE_N)).astype(np.float32)
        if lq <= 32:                                   # rep4: the 32 query tokens sit in all four lane quarters
            S[32:64] = S[64:96] = S[96:128] = S[:32]
        S[:, t["used"]:] = 99.0                        # columns no doc owns must never be read into a score
        for col, L in t["docs"]:                       # pad columns hold garbage that beats every real score
            S[:, col + L: col + ((L + 7) & ~7)] = 77.0
        want = direct(S, t, lq)
        assert np.array_equal(epilogue_v1(S, t, lq), want)
        assert np.array_equal(epilogue_v2(S, t, lq), want)


def test_first_doc_per_quarter_meta():
    t = pack_tiles([60, 10, 120, 7])[0]                # columns [0,64) [64,80) [80,200) [200,208)
    assert t["qf"] == [1, 2, 2] and t["used"] == 208
    t = pack_tiles([40])[0]
    assert t["qf"] == [-1, -1, -1]
    t = pack_tiles([256])[0]
    assert t["qf"] == [0, 0, 0]
    t = pack_tiles([64, 64, 64, 64])[0]
    assert t["qf"] == [1, 2, 3]
