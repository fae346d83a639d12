"""CPU fuzz of the oracle pin: the from-scratch port (oracle/pbwt_oracle.c) against the unmodified reference compiled
in place (oracle/_ref/libpbwtref.so, driver oracle/ref_driver.c) on panels drawn by hypothesis -- copies of a few
founder haplotypes with sparse flips, so that ties, long allele runs, monomorphic sites and deep match blocks all
occur -- for every function of the hot path the GPU parity tests later check against the port:
build (pbwtCore.c:573-585), the AD / A sweeps of a stored stream (:527-557), matchMaximalWithin (pbwtMatch.c:115-142),
matchLongWithin2 (:85-113), pbwtBuildReverse (pbwtCore.c:151-191) and matchSequencesSweep (pbwtMatch.c:363-443).
Skipped where the compiled reference is absent (it is built by __graft_entry__.build() wherever /root/reference exists
and travels with the repo snapshot)."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import oracle_lib as O

pytestmark = pytest.mark.skipif(O.reference() is None, reason="compiled reference (oracle/_ref) not present")


def founder_panel(rng, M, N, founders, flip):
    """M haplotypes, each a copy of one of `founders` random haplotypes that switches founder now and then, plus flips."""
    F = (rng.random((founders, N)) < rng.choice([0.05, 0.3, 0.5])).astype(np.uint8)
    hap = np.empty((M, N), np.uint8)
    for h in range(M):
        src = rng.integers(founders)
        for k in range(N):
            if rng.random() < 0.05:
                src = rng.integers(founders)
            hap[h, k] = F[src, k]
    hap ^= (rng.random((M, N)) < flip).astype(np.uint8)
    return hap


@settings(max_examples=60, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(M=st.integers(2, 70), N=st.integers(0, 48), founders=st.integers(1, 5), flip=st.sampled_from([0.0, 0.01, 0.1]),
       seed=st.integers(0, 2**31 - 1), permuted=st.booleans())
def test_port_equals_reference_on_founder_panels(M, N, founders, flip, seed, permuted):
    rng = np.random.default_rng(seed)
    hap = founder_panel(rng, M, N, founders, flip)
    x = O.pack_panel(hap) if N else np.zeros((0, O.words_per_col(M)), np.uint32)
    a0 = rng.permutation(M).astype(np.int32) if permuted else None
    ref, orc = O.reference(), O.port()
    r, o = ref.build(x, M, N, a0, True), orc.build(x, M, N, a0, True)
    assert r[0].tobytes() == o[0].tobytes() and (r[1] == o[1]).all() and (r[2] == o[2]).all()
    yz = r[0]
    assert (ref.build(x, M, N, a0, False)[0] == yz).all()  # the stream does not depend on carrying d
    (mr, nr), (mo, no) = ref.max_within(yz, M, N, a0), orc.max_within(yz, M, N, a0)
    assert nr == no and (mr == mo).all()
    for u, v in zip(ref.sweep_ad(yz, M, N, a0), orc.sweep_ad(yz, M, N, a0)):
        assert (u == v).all()
    assert (ref.sweep_a(yz, M, N, a0) == orc.sweep_a(yz, M, N, a0)).all()
    if N:
        for L in {1, max(1, N // 3), N}:
            (lr, nlr), (lo, nlo) = ref.long_within(yz, M, N, L, a0), orc.long_within(yz, M, N, L, a0)
            assert nlr == nlo and (lr == lo).all()
        for u, v in zip(ref.build_reverse(yz, M, N, a0), orc.build_reverse(yz, M, N, a0)):
            assert u.tobytes() == v.tobytes()


@settings(max_examples=40, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(M=st.integers(2, 60), Q=st.integers(1, 12), N=st.integers(1, 40), founders=st.integers(1, 4),
       flip=st.sampled_from([0.0, 0.02, 0.2]), seed=st.integers(0, 2**31 - 1))
def test_port_equals_reference_query_sweep(M, Q, N, founders, flip, seed):
    rng = np.random.default_rng(seed)
    hap = founder_panel(rng, M + Q, N, founders, flip)
    hp, hq = hap[:M], hap[M:].copy()
    if Q > 1:
        hq[0] = hp[rng.integers(M)]  # a query that is a panel haplotype
    ref, orc = O.reference(), O.port()
    pyz = ref.build(O.pack_panel(hp), M, N, with_d=False)[0]
    qyz = ref.build(O.pack_panel(hq), Q, N, with_d=False)[0]
    (rr, rn), (oo, on) = ref.match_sweep(pyz, M, N, qyz, Q), orc.match_sweep(pyz, M, N, qyz, Q)
    assert len(rr) == len(oo) and (rr == oo).all()
    assert len(rn) == len(on) and (rn == on).all()
