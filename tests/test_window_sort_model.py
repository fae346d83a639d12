"""CPU test of the identity behind pass 1 as a sort (pbwt_b200/csrc/window_sort.cu): sweeping a
window of sites with pbwtCursorForwardsAD from a fresh cursor (pbwtCore.c:485-508) gives

    order      = the start order, stably sorted by the window's alleles read backwards (last site most significant),
    divergence = 1 + the most recent site at which the two neighbours differ, 0 for haplotypes the window cannot
                 separate,

so the window's node can be computed by a radix sort and a neighbour comparison instead of a site-by-site sweep.
The sweep is the oracle's; the sort is a few lines of numpy."""
import numpy as np
import pytest

import oracle_lib as O


def node_by_sorting(hap, a_start):
    """hap: uint8 [M, n] alleles of the window's sites; a_start: int [M] start order -> (order, divergence[1:M])."""
    M, n = hap.shape
    order = np.asarray(a_start, np.int64)
    for k in range(n):  # LSD radix sort, one stable pass per site (np.argsort(kind="stable") = the stable partition)
        order = order[np.argsort(hap[order, k], kind="stable")]
    d = np.zeros(M, np.int64)
    for i in range(1, M):
        diff = np.nonzero(hap[order[i - 1]] != hap[order[i]])[0]
        d[i] = diff[-1] + 1 if diff.size else 0
    return order, d


@pytest.mark.parametrize("M,n,dens,seed", [(2, 1, 0.5, 1), (9, 7, 0.5, 2), (64, 33, 0.3, 3), (300, 90, 0.05, 4), (257, 16, 0.5, 5),
                                           (500, 200, 0.01, 6)])
def test_window_node_is_a_sort(M, n, dens, seed):
    rng = np.random.default_rng(seed)
    hap = (rng.random((M, n)) < dens).astype(np.uint8)
    if M > 8:
        hap[: M // 4] = hap[0]  # ties that survive the window
    a0 = rng.permutation(M).astype(np.int32)
    orc = O.port()
    _, a_end, d_end = orc.build(O.pack_panel(hap), M, n, a0, with_d=True)  # the sweep (fresh cursor: d = 0)
    order, d = node_by_sorting(hap, a0)
    assert (order == a_end).all()
    assert (d[1:] == d_end[1:M]).all()
    assert d_end[0] == n + 1 and d_end[M] == n + 1  # the sentinels are not part of the node
