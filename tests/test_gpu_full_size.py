"""BASELINE.json configurations at their FULL sizes on the GPU, checked through size-independent properties
(the CPU oracle would take minutes to hours at these sizes; small-size parity is in test_gpu_parity.py):

  * aFend is a permutation, d sentinels / ranges are right;
  * the emitted pack3 stream, swept again from scratch by the stored-stream route (device unpack3 +
    ForwardsReadAD), reproduces aFend and d bit for bit -- encode -> decode -> sweep round trip;
  * match records are well formed, in callback order, and a random sample is verified against the panel itself:
    the two haplotypes agree on [start, end), differ at start-1 and (for end < N) at end -- i.e. each sampled
    record is a real, maximal match (the reference's own -check criterion, pbwtMatch.c:33-44).
Timings are printed for the record (they are not bench values)."""
import time

import numpy as np
import pytest

import pbwt_b200 as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = P.Engine(0)
    yield e
    e.close()


def hap_bits(panel, h, k0, k1):
    """alleles of haplotype h at sites [k0, k1) from the device panel (downloaded column range)."""
    x = panel.to_host(k0, k1 - k0)
    return (x[:, h >> 5] >> np.uint32(h & 31)) & 1


def check_records_are_maximal_matches(panel, rec, N, rng, sample=60, query_panel=None):
    for r in rec[rng.choice(len(rec), size=min(sample, len(rec)), replace=False)]:
        ai, bi, s, e = int(r["ai"]), int(r["bi"]), int(r["start"]), int(r["end"])
        lo, hi = max(s - 1, 0), min(e + 1, N)
        xa = hap_bits(query_panel if query_panel is not None else panel, ai, lo, hi)
        xb = hap_bits(panel, bi, lo, hi)
        body = slice(s - lo, e - lo)
        assert (xa[body] == xb[body]).all(), "not a match"
        if s > 0:
            assert xa[0] != xb[0], "extends backwards"
        if e < N:
            assert xa[-1] != xb[-1], "extends forwards"


def test_config2_10k_x_100k_build_maxwithin(eng):
    M, N = 10000, 100000
    panel = eng.panel_synthetic(M, N, seed=0x5EED0002, L=2000, mu=2.0)
    t0 = time.time()
    res = eng.build_device(panel, P.WANT_YZ | P.WANT_MAXWITHIN)
    eng.synchronize()
    dt = time.time() - t0
    out = res.fetch()
    res.free()
    print(f"\nconfig2: {M}x{N} build+maxWithin {dt*1e3:.1f} ms (first call), {len(out['matches'])} records, yz {out['yz'].size} B")
    assert sorted(out["a_end"].tolist()) == list(range(M))
    assert out["d_end"][0] == N + 1 and out["d_end"][M] == N + 1 and 0 <= out["d_end"][1:M].min() and out["d_end"][1:M].max() <= N
    again = eng.max_within(out["yz"], M, N)  # stored-stream route, window-parallel in position space
    assert (again["a_end"] == out["a_end"]).all() and (again["d_end"] == out["d_end"]).all()
    assert len(again["matches"]) == len(out["matches"]) and (again["matches"] == out["matches"]).all()
    m = out["matches"]
    assert ((m["start"] <= m["end"]) & (m["end"] <= N) & (m["ai"] != m["bi"]) & (m["ai"] >= 0) & (m["ai"] < M) & (m["bi"] < M)).all()
    assert (np.diff(m["end"]) >= 0).all()
    check_records_are_maximal_matches(panel, m, N, np.random.default_rng(2))
    panel.free()


def test_config3_100k_x_200k_build_pack3(eng):
    M, N = 100000, 200000
    panel = eng.panel_synthetic(M, N, seed=0x5EED0003, L=1000, mu=8.0)
    eng.synchronize()
    t0 = time.time()
    res = eng.build_device(panel, P.WANT_YZ)
    eng.synchronize()
    dt = time.time() - t0
    out = res.fetch()
    res.free()
    print(f"\nconfig3: {M}x{N} build+pack3 {dt*1e3:.0f} ms = {M*N/dt:.3e} hap-sites/s, yz {out['yz'].size} B")
    assert np.array_equal(np.sort(out["a_end"]), np.arange(M))
    r2 = eng.sweep_yz_device(out["yz"], M, N, P.WANT_D)  # decode + sweep the emitted stream
    again = r2.fetch()
    r2.free()
    assert (again["a_end"] == out["a_end"]).all()
    assert again["d_end"][0] == N + 1 and again["d_end"][M] == N + 1
    panel.free()


def test_config4_1m_x_100k_build_with_d(eng):
    M, N = 1000000, 100000
    panel = eng.panel_synthetic(M, N, seed=0x5EED0004, L=1000, mu=8.0)
    eng.synchronize()
    t0 = time.time()
    res = eng.build_device(panel, P.WANT_YZ | P.WANT_D)
    eng.synchronize()
    dt = time.time() - t0
    out = res.fetch()
    res.free()
    panel.free()
    print(f"\nconfig4: {M}x{N} build(a,d)+pack3 on one GPU {dt:.2f} s = {M*N/dt:.3e} hap-sites/s, yz {out['yz'].size} B")
    assert np.array_equal(np.sort(out["a_end"]), np.arange(M))
    assert out["d_end"][0] == N + 1 and out["d_end"][M] == N + 1 and out["d_end"][1:M].max() <= N
    r2 = eng.sweep_yz_device(out["yz"], M, N, P.WANT_D)
    again = r2.fetch()
    r2.free()
    assert (again["a_end"] == out["a_end"]).all() and (again["d_end"] == out["d_end"]).all()


def test_config5_matchdynamic_1000_queries_vs_100k_panel(eng):
    M, Q, N = 100000, 1000, 20000  # BASELINE gives no N for this config; 20,000 sites keep the run short
    both = eng.panel_synthetic(M + Q, N, seed=0x5EED0005, L=1000, mu=8.0)
    x = both.to_host()
    both.free()
    # split the (M+Q)-haplotype panel: first M haplotypes = reference panel, last Q = queries
    bits = np.unpackbits(np.ascontiguousarray(x).view(np.uint8), axis=1, bitorder="little")
    def repack(b):
        n, m = b.shape
        wpc = P.words_per_col(m)
        pad = np.zeros((n, wpc * 32), np.uint8)
        pad[:, :m] = b
        return np.ascontiguousarray(np.packbits(pad, axis=1, bitorder="little")).view("<u4").reshape(n, wpc)
    px, qx = repack(bits[:, :M]), repack(bits[:, M:M + Q])
    del bits, x
    panel, queries = eng.panel_from_host(px, M), eng.panel_from_host(qx, Q)
    rp = eng.build_device(panel, P.WANT_YZ); pyz = rp.fetch()["yz"]; rp.free()
    rq = eng.build_device(queries, P.WANT_YZ); qyz = rq.fetch()["yz"]; rq.free()
    t0 = time.time()
    rec, nom = eng.match_sweep(pyz, M, N, qyz, Q)
    dt = time.time() - t0
    print(f"\nconfig5: {Q} queries vs {M}x{N} panel: {dt:.2f} s = {M*N/dt:.3e} panel hap-sites/s, {len(rec)} records, {len(nom)} no-match")
    assert len(rec) > Q and ((rec["ai"] >= 0) & (rec["ai"] < Q) & (rec["bi"] >= 0) & (rec["bi"] < M) & (rec["start"] <= rec["end"]) & (rec["end"] <= N)).all()
    assert (np.diff(rec["end"]) >= 0).all()  # callback order: site ascending
    nz = rec[rec["start"] < rec["end"]]
    check_records_are_maximal_matches(panel, nz, N, np.random.default_rng(5), query_panel=queries)
    # every query has at least one match open at N (the final flush reports it)
    assert set(rec["ai"][rec["end"] == N].tolist()) == set(range(Q))
    panel.free()
    queries.free()
