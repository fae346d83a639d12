"""Oracle-free checks used by the full-size GPU tests (tests/test_gpu_full_size.py); exercised on the CPU against the
oracle in tests/test_fullsize_checks_cpu.py so that a wrong helper is found without a GPU."""
import numpy as np


def check_state_against_panel(xt, first, M, k, a, d, rng, sample=4000):
    """(a, d) entering site k checked against the bit columns themselves (xt = columns [first, k)) on a sample of
    adjacent pairs: a[i-1] and a[i] agree on [d[i], k); if d[i] > 0 they differ at d[i]-1 with the 0 above the 1
    (sorted by reversed prefix), and if d[i] == 0 they are in index order (stable from the identity start).  Pairs
    whose match starts at or before `first` are only required to agree on all of [first, k)."""
    assert np.array_equal(np.sort(a), np.arange(M)), "a is not a permutation"
    assert d[0] == k + 1 and d[M] == k + 1 and d[1:M].min() >= 0 and d[1:M].max() <= k
    idx = np.unique(np.concatenate([rng.integers(1, M, size=sample), [1, M - 1]]))
    up, lo, dd = a[idx - 1].astype(np.int64), a[idx].astype(np.int64), d[idx].astype(np.int64)
    agree_from = np.full(idx.size, k, np.int64)  # smallest s such that the pair agrees on [s, k)
    alive = np.ones(idx.size, bool)
    bit_above = np.zeros(idx.size, np.uint8)
    for s in range(k - 1, first - 1, -1):  # walk back from the cut; pairs drop out at their first difference
        col = xt[s - first]
        bu = (col[up >> 5] >> (up & 31).astype(np.uint32)) & 1
        bl = (col[lo >> 5] >> (lo & 31).astype(np.uint32)) & 1
        diff = alive & (bu != bl)
        bit_above[diff] = bu[diff]
        alive &= ~diff
        agree_from[alive] = s
        if not alive.any():
            break
    dropped = ~alive
    assert np.array_equal(agree_from[dropped], dd[dropped]), "divergence is not the start of the match with the neighbour above"
    assert (bit_above[dropped] == 0).all(), "pair out of order (1 above 0 at the first difference)"
    assert (dd[alive] <= first).all(), "pair agrees further back than its divergence says"
    if first == 0:
        assert (up[alive] < lo[alive]).all(), "tied pair not in start order"
    return int(dropped.sum()), int(alive.sum())


_P3LEN = np.array([c if c < 64 else ((c - 64) << 6) if c < 96 else ((c - 96) << 11) for c in range(128)], np.int64)


def column_offsets(yz, M, cols):
    """byte offset in a pack3 stream at which each of the given columns starts (pbwtCore.c:232-238: run lengths)."""
    ends = np.cumsum(_P3LEN[yz & 0x7F])  # symbols decoded after each byte
    return [0 if c == 0 else int(np.searchsorted(ends, np.int64(c) * M, side="left")) + 1 for c in cols]


def split_panel(x, M, Q):
    """first M haplotypes = reference panel, last Q = queries"""
    bits = np.unpackbits(np.ascontiguousarray(x).view(np.uint8), axis=1, bitorder="little")

    def repack(b):
        n, m = b.shape
        wpc = 4 * ((m + 127) // 128)
        pad = np.zeros((n, wpc * 32), np.uint8)
        pad[:, :m] = b
        return np.ascontiguousarray(np.packbits(pad, axis=1, bitorder="little")).view("<u4").reshape(n, wpc)

    return repack(bits[:, :M]), repack(bits[:, M:M + Q])


