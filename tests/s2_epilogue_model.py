"""Executable model (numpy) of the Stage-2 kernel's tile packing and epilogue index logic
(csrc/s2_maxsim.cu): greedy packing of 8-row padded docs into 256-column tiles, the per-quarter
drain of the TMEM accumulator and the finalize step -- both the validated layout (V1: "-inf"
initialisation + blind 4-way max) and the opt-in V2 (first doc per quarter from the tile meta,
64-column groups, 8-column units, quarter ranges in finalize).  Test infrastructure: lets the
index arithmetic be checked on the CPU against a direct max over each doc's columns."""
import numpy as np

TILE_N, TILE_M = 256, 128
NEG = -np.inf


def pack_tiles(lens):
    """Producer: docs (token counts) -> tiles [{used, docs:[(col, len)], qf:[qf1,qf2,qf3]}]."""
    tiles, cur, cols, qf = [], [], 0, [-1, -1, -1]
    for L in lens:
        pad = (L + 7) & ~7
        if cols + pad > TILE_N:
            tiles.append({"used": cols, "docs": cur, "qf": qf})
            cur, cols, qf = [], 0, [-1, -1, -1]
        for q in range(3):
            if qf[q] < 0 and cols + pad > 64 * (q + 1):
                qf[q] = len(cur)
        cur.append((cols, L))
        cols += pad
    if cur:
        tiles.append({"used": cols, "docs": cur, "qf": qf})
    return tiles


def epilogue_v2(S, tile, lq):
    """S: [128, 256] accumulator (lane = query token, replicated in the 4 quarters when lq <= 32)."""
    used, docs = tile["used"], tile["docs"]
    nd, rep4 = len(docs), lq <= 32
    mvals = np.full((nd, TILE_M), np.nan, np.float32)            # NaN = never written
    for quarter in range(4):
        c_lo = quarter * 64 if rep4 else 0
        c_hi = min(used, c_lo + 64) if rep4 else used
        active = True if rep4 else quarter * 32 < lq
        d = (0 if quarter == 0 else tile["qf"][quarter - 1]) if rep4 else 0
        if not (active and d >= 0 and c_lo < c_hi):
            continue
        rows = slice(quarter * 32, quarter * 32 + 32)
        seg_end = docs[d][0] + docs[d][1]
        seg_next = docs[d][0] + ((docs[d][1] + 7) & ~7)
        best = np.full(32, NEG, np.float32)
        for g0 in range(c_lo, c_hi, 64):
            for u in range(8):
                cu = g0 + u * 8
                if cu < c_hi:
                    assert u < 4 or g0 + 32 < c_hi, "second tcgen05.ld must have been issued"
                    if cu >= seg_next:
                        mvals[d, rows] = best
                        d += 1
                        best = np.full(32, NEG, np.float32)
                        seg_end = docs[d][0] + docs[d][1]
                        seg_next = docs[d][0] + ((docs[d][1] + 7) & ~7)
                    v = S[rows, cu:cu + 8].copy()
                    if cu + 8 > seg_end:
                        for j in range(8):
                            if not cu + j < seg_end:
                                v[:, j] = NEG
                    best = np.maximum(best, v.max(axis=1))
        mvals[d, rows] = best
    out = np.zeros(nd, np.float32)
    for d in range(nd):
        col, L = docs[d]
        q_lo = col >> 6 if rep4 else 0
        q_hi = (col + ((L + 7) & ~7) - 1) >> 6 if rep4 else 0
        m = np.empty(lq, np.float32)
        for i in range(lq):
            v = mvals[d, (q_lo * 32 if rep4 else 0) + i]
            for qq in range(q_lo + 1, q_hi + 1):
                v = max(v, mvals[d, qq * 32 + i])
            m[i] = v
        assert not np.isnan(m).any(), "finalize read a value no warp wrote"
        out[d] = m.mean(dtype=np.float32)
    return out


def epilogue_v1(S, tile, lq):
    used, docs = tile["used"], tile["docs"]
    nd, rep4 = len(docs), lq <= 32
    mvals = np.full((nd, TILE_M), np.nan, np.float32)
    for quarter in range(4):
        c_lo = quarter * 64 if rep4 else 0
        c_hi = min(used, c_lo + 64) if rep4 else used
        if not (True if rep4 else quarter * 32 < lq):
            continue
        rows = slice(quarter * 32, quarter * 32 + 32)
        d = -1
        for dd in range(nd):
            col = docs[dd][0]
            nxt = col + ((docs[dd][1] + 7) & ~7)
            if nxt <= c_lo or col >= c_hi:
                mvals[dd, rows] = NEG
            elif d < 0:
                d = dd
        if d < 0:
            continue
        seg_end = docs[d][0] + docs[d][1]
        seg_next = docs[d][0] + ((docs[d][1] + 7) & ~7)
        best = np.full(32, NEG, np.float32)
        for c0 in range(c_lo, c_hi, 32):
            for u in range(4):
                cu = c0 + u * 8
                if cu < c_hi:
                    if cu >= seg_next:
                        mvals[d, rows] = best
                        d += 1
                        best = np.full(32, NEG, np.float32)
                        seg_end = docs[d][0] + docs[d][1]
                        seg_next = docs[d][0] + ((docs[d][1] + 7) & ~7)
                    for j in range(8):
                        if cu + j < seg_end:
                            best = np.maximum(best, S[rows, cu + j])
        mvals[d, rows] = best
    out = np.zeros(nd, np.float32)
    for d in range(nd):
        m = np.empty(lq, np.float32)
        for i in range(lq):
            v = mvals[d, i]
            if rep4:
                v = max(max(v, mvals[d, 32 + i]), max(mvals[d, 64 + i], mvals[d, 96 + i]))
            m[i] = v
        assert not np.isnan(m).any()
        out[d] = m.mean(dtype=np.float32)
    return out


def direct(S, tile, lq):
    return np.array([S[:lq, col:col + L].max(axis=1).mean(dtype=np.float32) for col, L in tile["docs"]], np.float32)
