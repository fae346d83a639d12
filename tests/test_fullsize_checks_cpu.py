"""The helpers of the full-size GPU tests, exercised on the CPU with the oracle standing in for the device: a correct
state passes, a corrupted one is caught."""
import numpy as np
import pytest

import oracle_lib as O
from fullsize_checks import check_state_against_panel, column_offsets, split_panel


def test_state_check_accepts_the_oracle_and_rejects_corruption():
    M, N = 700, 900
    x = O.generate_panel(M, N, seed=77, L=50, mu=3.0)
    a_tr, d_tr, _, _ = O.port().trace(x, M, N)
    rng = np.random.default_rng(1)
    for k, first in ((N, 0), (N, 500), (400, 0), (400, 250)):
        dropped, alive = check_state_against_panel(x[first:k], first, M, k, a_tr[k], d_tr[k], rng, sample=M)
        assert dropped + alive > 0
    a, d = a_tr[N].copy(), d_tr[N].copy()
    i = int(np.argmax(d[1:M])) + 1
    d[i] -= 1  # claims a longer match than there is
    with pytest.raises(AssertionError):
        check_state_against_panel(x, 0, M, N, a, d, rng, sample=4 * M)
    d = d_tr[N].copy()
    j = int(np.argmin(np.where(d[1:M] > 0, d[1:M], N))) + 1
    d[j] += 1  # claims a shorter one
    with pytest.raises(AssertionError):
        check_state_against_panel(x, 0, M, N, a, d, rng, sample=4 * M)
    a[[i - 1, i]] = a[[i, i - 1]]  # a swapped pair
    with pytest.raises(AssertionError):
        check_state_against_panel(x, 0, M, N, a, d_tr[N], rng, sample=4 * M)


def test_column_offsets_of_a_pack3_stream():
    M, N = 5000, 60
    x = O.generate_panel(M, N, seed=5, L=20, mu=4.0)
    orc = O.port()
    yz, _, _ = orc.build(x, M, N)
    sizes = [orc.build(x[:k], M, k)[0].size for k in (0, 1, 17, 59, 60)]
    assert column_offsets(yz, M, [0, 1, 17, 59, 60]) == sizes


def test_split_panel():
    M, Q, N = 130, 40, 9
    x = O.generate_panel(M + Q, N, seed=9, L=4, mu=20.0)
    hap = O.unpack_panel(x, M + Q)
    px, qx = split_panel(x, M, Q)
    assert np.array_equal(O.unpack_panel(px, M), hap[:M]) and np.array_equal(O.unpack_panel(qx, Q), hap[M:])
