"""CPU test of the block characterisation behind -longWithin on the GPU (match_scan.cuh):
matchLongWithin2 (pbwtMatch.c:85-113) cuts state k into blocks at every position with d > k - L and reports a block
iff it holds both alleles and does not run to M.  Equivalently: iff it contains an interior head -- a position h with
y[h] != y[h-1] and d[h] <= k - L -- and ends before M; and the number of records is zeros x ones of the block."""
import numpy as np
import pytest


def blocks_reference(d, y, M, thr):
    """(start, end, records) of every block the reference reports at this state."""
    out, start, n0, n1 = [], 0, 0, 0
    for i in range(M):
        if d[i] > thr:
            if n0 and n1:
                recs = sum(1 for ia in range(start, i) for ib in range(ia + 1, i) if y[ib] != y[ia])
                out.append((start, i, recs))
            n0 = n1 = 0
            start = i
        if y[i] == 0:
            n0 += 1
        else:
            n1 += 1
    return out


def blocks_by_interior_heads(d, y, M, thr):
    out = []
    for h in range(1, M):
        if y[h] != y[h - 1] and d[h] <= thr:  # interior head
            q = h
            while q > 0 and d[q] <= thr:
                q -= 1
            if any(y[p] == y[h] for p in range(q, h)):  # not the block's first interior head
                continue
            r = h + 1
            while r < M and d[r] <= thr:
                r += 1
            if r < M:
                z = sum(1 for p in range(q, r) if y[p] == 0)
                out.append((q, r, z * (r - q - z)))
    return out


@pytest.mark.parametrize("M,dens,seed", [(5, 0.5, 1), (40, 0.5, 2), (40, 0.1, 3), (120, 0.3, 4)])
def test_reported_blocks_are_those_with_an_interior_head(M, dens, seed):
    rng = np.random.default_rng(seed)
    for _ in range(300):
        k = int(rng.integers(1, 30))
        L = int(rng.integers(1, k + 2))
        d = rng.integers(0, k + 1, size=M + 1)
        d[0] = k + 1
        d[M] = k + 1
        y = (rng.random(M) < dens).astype(int)
        assert blocks_reference(d, y, M, k - L) == blocks_by_interior_heads(d, y, M, k - L)
