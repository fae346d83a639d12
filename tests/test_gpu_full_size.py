"""BASELINE.json configurations at their FULL sizes on the GPU, pinned to the CPU oracle.

The checker is `oracle_lib.reference()` -- the unmodified reference compiled in place into oracle/_ref (it travels
to the GPU box as a prebuilt .so) -- and only falls back to the from-scratch port when _ref is absent.  What is
compared, per config:

  * C2 (10,000 x 100,000, build + maxWithin): EVERYTHING at full size, bit for bit: pack3 bytes, aFend, d, and all
    ~18.5 M match records in callback order (the reference needs ~10 s for it).
  * C3 (100,000 x 200,000, build + pack3): the pack3 stream and aFend of the whole panel against the reference's
    a-only build (pbwtCursorWriteForwards; ~45 s of CPU), plus the first 5,000 sites built separately with forced
    windows (so that the window sort / combine scan of the wide-panel path face the oracle with d as well).
  * C4 (1,000,000 x 100,000, build with d): the first 2,000 sites bit-exact (a truncated build is a prefix of the
    full one: pack3 prefix + the cursor state at the cut through pbwtgpu_cursor_at), and segments that straddle the
    window boundaries swept by the oracle FROM the GPU's state entering the segment (pack3 bytes of the segment and
    the state leaving it must agree) -- the step function is pinned where the window-parallel scheme stitches.
  * C5 (1,000 queries vs 100,000 x 100,000): every record and every "no match" event of matchSequencesSweep
    against the reference (~35 s of CPU).
On top of that the size-independent properties: the final (a, d) is verified against the PANEL ITSELF on a sample of
adjacent pairs (they agree on [d, N), differ at d-1 with 0 above 1, or are in index order when d = 0 -- over all
pairs this characterises the PBWT order and divergence completely), and sampled match records are real maximal
matches (the reference's own -check criterion, pbwtMatch.c:33-44).
Timings are printed for the record (they are not bench values)."""
import time

import numpy as np
import pytest

import oracle_lib as O
import pbwt_b200 as P
from fullsize_checks import check_state_against_panel, column_offsets, split_panel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = P.Engine(0)
    yield e
    e.close()


def checker():
    """The compiled reference when oracle/_ref travelled to this box, else the pinned port."""
    ref = O.reference()
    return (ref, "reference") if ref is not None else (O.port(), "port")


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


def test_config2_10k_x_100k_build_maxwithin_everything_vs_reference(eng):
    M, N = 10000, 100000
    chk, kind = checker()
    panel = eng.panel_synthetic(M, N, seed=0x5EED0002, L=2000, mu=2.0)
    t0 = time.time()
    res = eng.build_device(panel, P.WANT_YZ | P.WANT_MAXWITHIN)
    eng.synchronize()
    dt = time.time() - t0
    out = res.fetch()
    res.free()
    x = panel.to_host()
    t0 = time.time()
    yz, a_end, d_end = chk.build(x, M, N, with_d=True)
    rec, n = chk.max_within(yz, M, N, cap=len(out["matches"]) + 16)
    print(f"\nconfig2: {M}x{N} build+maxWithin {dt*1e3:.1f} ms (first call), {len(out['matches'])} records, yz {out['yz'].size} B; "
          f"{kind} on the CPU: {time.time()-t0:.1f} s")
    assert out["yz"].tobytes() == yz.tobytes(), "pack3 stream differs"
    assert np.array_equal(out["a_end"], a_end) and np.array_equal(out["d_end"], d_end)
    assert n == len(out["matches"]) and np.array_equal(out["matches"], rec[:n]), "match records differ"
    # the stored-stream route (device unpack3 + ForwardsReadAD, windows combined in position space) on the same stream
    again = eng.max_within(out["yz"], M, N)
    assert np.array_equal(again["a_end"], a_end) and np.array_equal(again["d_end"], d_end)
    assert np.array_equal(again["matches"], rec[:n])
    check_state_against_panel(x[N - 30000:], N - 30000, M, N, out["a_end"], out["d_end"], np.random.default_rng(2), sample=2000)
    check_records_are_maximal_matches(panel, out["matches"], N, np.random.default_rng(2))
    panel.free()


def test_config3_100k_x_200k_build_pack3_whole_stream_vs_reference(eng):
    M, N = 100000, 200000
    chk, kind = checker()
    panel = eng.panel_synthetic(M, N, seed=0x5EED0003, L=1000, mu=8.0)
    eng.synchronize()
    t0 = time.time()
    res = eng.build_device(panel, P.WANT_YZ)
    eng.synchronize()
    dt = time.time() - t0
    out = res.fetch()
    res.free()
    print(f"\nconfig3: {M}x{N} build+pack3 {dt*1e3:.0f} ms = {M*N/dt:.3e} hap-sites/s, yz {out['yz'].size} B")
    x = panel.to_host()
    panel.free()
    t0 = time.time()
    yz, a_end, _ = chk.build(x, M, N, with_d=False)  # pbwtCursorWriteForwards over all 200,000 sites
    print(f"config3: {kind} a-only build of the whole panel on the CPU: {time.time()-t0:.1f} s")
    assert out["yz"].tobytes() == yz.tobytes(), "pack3 stream differs"
    assert np.array_equal(out["a_end"], a_end)
    # with d, and with windows forced onto a short prefix: window sort, combine scan and pass 2 of the wide path
    n5 = 5000
    eng.set_option("windows", 12)
    try:
        got = eng.build(x[:n5], M, n5, with_d=True)
    finally:
        eng.set_option("windows", 0)
    yz5, a5, d5 = chk.build(x[:n5], M, n5, with_d=True)
    assert got["yz"].tobytes() == yz5.tobytes() and np.array_equal(got["a_end"], a5) and np.array_equal(got["d_end"], d5)
    assert out["yz"][: yz5.size].tobytes() == yz5.tobytes(), "the truncated build is not a prefix of the full one"
    # d of the full panel: the stored-stream route over the emitted bytes, then against the panel itself
    r2 = eng.sweep_yz_device(out["yz"], M, N, P.WANT_D)
    again = r2.fetch()
    r2.free()
    assert np.array_equal(again["a_end"], a_end)
    print("config3: pairs checked against the panel (differ inside the last 30,000 sites / agree on all of them):",
          check_state_against_panel(x[N - 30000:], N - 30000, M, N, again["a_end"], again["d_end"], np.random.default_rng(3), sample=3000))


def test_config4_1m_x_100k_build_with_d_prefix_and_stitches_vs_oracle(eng):
    M, N = 1000000, 100000
    chk, kind = checker()
    panel = eng.panel_synthetic(M, N, seed=0x5EED0004, L=1000, mu=8.0)
    eng.synchronize()
    t0 = time.time()
    res = eng.build_device(panel, P.WANT_YZ | P.WANT_D)
    eng.synchronize()
    dt = time.time() - t0
    out = res.fetch()
    res.free()
    print(f"\nconfig4: {M}x{N} build(a,d)+pack3 on one GPU {dt:.2f} s = {M*N/dt:.3e} hap-sites/s, yz {out['yz'].size} B")
    assert np.array_equal(np.sort(out["a_end"]), np.arange(M))
    # (1) prefix: the reference's first 2,000 sites
    n2 = 2000
    x2 = panel.to_host(0, n2)
    t0 = time.time()
    yz2, a2, d2 = chk.build(x2, M, n2, with_d=True)
    print(f"config4: {kind} build of the first {n2} sites on the CPU: {time.time()-t0:.1f} s")
    assert out["yz"][: yz2.size].tobytes() == yz2.tobytes(), "pack3 prefix differs"
    _, a_cut, d_cut, _ = eng.cursor_at(out["yz"], M, N, n2)
    assert np.array_equal(a_cut, a2) and np.array_equal(d_cut, d2), "cursor state at the cut differs"
    # (2) stitches: the one-GPU plan sweeps this panel in 3 windows; segments across N/3, N/2 and 2N/3 are swept by
    # the oracle port (pinned to the reference) from the GPU's state entering them
    seg = 300
    for k0 in (N // 3 - seg // 2, N // 2 - seg // 2, 2 * N // 3 - seg // 2):
        _, a0, d0, _ = eng.cursor_at(out["yz"], M, N, k0)
        xs = panel.to_host(k0, seg)
        yzs, a1, d1, _ = O.sweep_from(xs, M, seg, k0, N, a0, d0)
        o0, o1 = column_offsets(out["yz"], M, [k0, k0 + seg])
        assert out["yz"][o0:o1].tobytes() == yzs.tobytes(), f"pack3 bytes of sites [{k0}, {k0+seg}) differ"
        _, ag, dg, _ = eng.cursor_at(out["yz"], M, N, k0 + seg)
        assert np.array_equal(ag, a1) and np.array_equal(dg, d1), f"state leaving the segment at {k0} differs"
    # (3) the final state against the panel itself (last 6,000 sites are enough for almost every pair: deeper pairs
    # are only required to agree there) and against the stored-stream route
    r2 = eng.sweep_yz_device(out["yz"], M, N, P.WANT_D)
    again = r2.fetch()
    r2.free()
    assert np.array_equal(again["a_end"], out["a_end"]) and np.array_equal(again["d_end"], out["d_end"])
    tail = 8000
    xt = panel.to_host(N - tail, tail)
    panel.free()
    print("config4: pairs checked against the panel (differ inside the last 8,000 sites / agree on all of them):",
          check_state_against_panel(xt, N - tail, M, N, out["a_end"], out["d_end"], np.random.default_rng(4), sample=3000))


def test_config5_matchdynamic_1000_queries_vs_100k_x_100k_every_record_vs_reference(eng):
    M, Q, N = 100000, 1000, 100000  # BASELINE gives no N for this config; SURVEY 8 says assume 100,000 sites
    chk, kind = checker()
    both = eng.panel_synthetic(M + Q, N, seed=0x5EED0005, L=1000, mu=8.0)
    x = both.to_host()
    both.free()
    px, qx = split_panel(x, M, Q)
    del x
    panel, queries = eng.panel_from_host(px, M), eng.panel_from_host(qx, Q)
    rp = eng.build_device(panel, P.WANT_YZ); pyz = rp.fetch()["yz"]; rp.free()
    rq = eng.build_device(queries, P.WANT_YZ); qyz = rq.fetch()["yz"]; rq.free()
    eng.synchronize()
    t0 = time.time()
    rec, nom = eng.match_sweep(pyz, M, N, qyz, Q)
    dt = time.time() - t0
    print(f"\nconfig5: {Q} queries vs {M}x{N} panel: {dt:.2f} s = {M*N/dt:.3e} panel hap-sites/s, {len(rec)} records, {len(nom)} no-match")
    t0 = time.time()
    want, wnom = chk.match_sweep(pyz, M, N, qyz, Q)
    print(f"config5: {kind} matchSequencesSweep on the CPU: {time.time()-t0:.1f} s")
    assert len(rec) == len(want) and np.array_equal(rec, want), "match records differ"
    assert len(nom) == len(wnom) and np.array_equal(nom, wnom), '"no match" events differ'
    nz = rec[rec["start"] < rec["end"]]
    check_records_are_maximal_matches(panel, nz, N, np.random.default_rng(5), query_panel=queries)
    panel.free()
    queries.free()
