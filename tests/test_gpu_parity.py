"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded inputs and against the committed golden fixtures.  Bit-exact: a[], d[], pack3 bytes and
match records in callback order."""
import os

import numpy as np
import pytest

import oracle_lib as O
import pbwt_b200

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def eng():
    e = pbwt_b200.Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def orc():
    return O.port()


def rec4(rec):
    return np.stack([rec["ai"], rec["bi"], rec["start"], rec["end"]], 1)


def check_build(eng, orc, x, M, N, a_start=None, matches=True):
    yz, a_end, d_end = orc.build(x, M, N, a_start, with_d=True)
    got = eng.build(x, M, N, a_start=a_start, with_d=True, with_matches=matches and N > 0 and M > 1)
    assert got["yz"].tobytes() == yz.tobytes()
    assert (got["a_end"] == a_end).all()
    assert (got["d_end"] == d_end).all()
    if matches and N > 0 and M > 1:
        rec, n = orc.max_within(yz, M, N, a_start)
        assert len(got["matches"]) == n
        assert (got["matches"] == rec).all()
    # a-only build (pbwtCursorWriteForwards) must give the same stream and aFend
    got2 = eng.build(x, M, N, a_start=a_start, with_d=False)
    assert got2["yz"].tobytes() == yz.tobytes() and (got2["a_end"] == a_end).all()
    # stored-stream route: device unpack3 + ForwardsReadAD [+ maxWithin] must agree as well
    if N > 0:
        got3 = eng.max_within(yz, M, N, a_start=a_start, with_matches=matches and M > 1)
        assert (got3["a_end"] == a_end).all() and (got3["d_end"] == d_end).all()
        if matches and M > 1:
            assert len(got3["matches"]) == n and (got3["matches"] == rec).all()
    return yz


@pytest.mark.parametrize("name", ["merge1", "merge2", "omni", "mosaic300x400"])
def test_golden_panels(eng, name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    M, N = int(g["M"]), int(g["N"])
    got = eng.build(g["xbits"], M, N, with_d=True, with_matches=True)
    assert got["yz"].tobytes() == g["yz"].tobytes()
    assert (got["a_end"] == g["a_end"]).all() and (got["d_end"] == g["d_end"]).all()
    assert (rec4(got["matches"]) == g["matches"]).all()
    got = eng.max_within(g["yz"], M, N)
    assert (got["a_end"] == g["a_end"]).all() and (got["d_end"] == g["d_end"]).all()
    assert (rec4(got["matches"]) == g["matches"]).all()


@pytest.mark.parametrize("M", [2, 3, 31, 32, 33, 63, 64, 65, 127, 128, 129, 1000, 1024, 1025, 2752, 4097, 10000, 13312])
def test_shapes_random(eng, orc, M):
    rng = np.random.default_rng(M)
    N = 38 if M < 5000 else 12
    for dens in (0.5, 0.03):
        hap = (rng.random((M, N)) < dens).astype(np.uint8)
        check_build(eng, orc, O.pack_panel(hap), M, N)


@pytest.mark.parametrize("N", [0, 1, 2, 5, 38])
def test_site_counts(eng, orc, N):
    rng = np.random.default_rng(100 + N)
    M = 200
    hap = (rng.random((M, N)) < 0.3).astype(np.uint8)
    check_build(eng, orc, O.pack_panel(hap) if N else np.zeros((0, O.words_per_col(M)), np.uint32), M, N)


def test_special_columns(eng, orc):
    M = 700
    cols = [np.zeros(M), np.ones(M), np.eye(1, M, 0)[0], np.eye(1, M, M - 1)[0], np.arange(M) % 2, (np.arange(M) // 3) % 2,
            np.zeros(M), np.arange(M) % 2, np.ones(M)]
    rng = np.random.default_rng(5)
    cols += [(rng.random(M) < 0.5) for _ in range(6)]
    hap = np.array(cols, np.uint8).T.copy()
    check_build(eng, orc, O.pack_panel(hap), M, hap.shape[1])


def test_non_identity_start_and_duplicates(eng, orc):
    rng = np.random.default_rng(9)
    M, N = 600, 60
    base = (rng.random((40, N)) < 0.4).astype(np.uint8)
    hap = base[rng.integers(0, 40, M)]  # blocks of identical haplotypes: O(B^2) reports at k == N
    a0 = rng.permutation(M).astype(np.int32)
    check_build(eng, orc, O.pack_panel(hap), M, N, a_start=a0)


def test_long_runs_pack3_thresholds(eng, orc):
    """Runs around 64 / 2048 / 63488 inside real columns (a-only path, large M)."""
    M = 20000
    cols = []
    for n in (63, 64, 65, 2047, 2048, 2049, 19000):
        c = np.zeros(M, np.uint8)
        c[0] = 1
        c[n + 1:n + 3] = 1
        cols.append(c)
    hap = np.array(cols, np.uint8).T.copy()
    x = O.pack_panel(hap)
    yz, a_end, _ = orc.build(x, M, hap.shape[1], with_d=False)
    got = eng.build(x, M, hap.shape[1], with_d=False)
    assert got["yz"].tobytes() == yz.tobytes() and (got["a_end"] == a_end).all()


def test_pack3_codec_vectors(eng):
    g = np.load(os.path.join(GOLD, "pack3_vectors.npz"))
    for key in g.files:
        n = int(key[1:])
        y = np.concatenate([[1], np.zeros(n, np.uint8), [1, 1]]).astype(np.uint8)[None, :]
        enc = eng.pack3(y)
        assert enc.tobytes() == g[key].tobytes(), key
        dec, zc = eng.unpack3(enc, y.shape[1], 1)
        assert (dec == y).all() and zc[0] == n, key


def test_pack3_roundtrip_random(eng, orc):
    rng = np.random.default_rng(3)
    for M, N, dens in [(1, 5, 0.5), (37, 50, 0.5), (5000, 20, 0.01), (70003, 3, 0.0005), (200003, 2, 0.5)]:
        y = (rng.random((N, M)) < dens).astype(np.uint8)
        enc = eng.pack3(y)
        ref = np.concatenate([orc.pack3(y[k]) for k in range(N)])
        assert enc.tobytes() == ref.tobytes(), (M, N)
        dec, zc = eng.unpack3(enc, M, N)
        assert (dec == y).all() and (zc == (y == 0).sum(1)).all(), (M, N)


def test_mosaic_panel_medium(eng, orc):
    """A realistic low-diversity panel at a size the oracle finishes in seconds."""
    M, N = 3000, 1500
    x = O.generate_panel(M, N, seed=0x5EED0002, L=300, mu=2.0)
    check_build(eng, orc, x, M, N)


def test_generator_matches_cpu(eng):
    for M, N, L, mu in [(500, 300, 50, 3.0), (10000, 40, 2000, 2.0), (33, 7, 5, 1.0)]:
        p = eng.panel_synthetic(M, N, seed=77, L=L, mu=mu)
        x = p.to_host()
        p.free()
        assert (x == O.generate_panel(M, N, seed=77, L=L, mu=mu)).all(), (M, N)


def test_full_size_properties(eng):
    """BASELINE config 2 shape (10,000 x N_prop sites): size-independent properties instead of the oracle:
    aFend is a permutation, the stream decodes to exactly M*N symbols whose per-column zero counts match the
    panel's popcounts, sweeping the emitted stream reproduces a/d, and match records are in range."""
    M, N = 10000, 4000
    p = eng.panel_synthetic(M, N, seed=0x5EED0002, L=2000, mu=2.0)
    res = eng.build_device(p, pbwt_b200.WANT_YZ | pbwt_b200.WANT_MAXWITHIN)
    out = res.fetch()
    res.free()
    x = p.to_host()
    p.free()
    assert sorted(out["a_end"].tolist()) == list(range(M))
    assert out["d_end"][0] == N + 1 and out["d_end"][M] == N + 1 and out["d_end"][1:M].max() <= N
    y, zc = eng.unpack3(out["yz"], M, N)
    ones = np.unpackbits(np.ascontiguousarray(x).view(np.uint8), axis=1, bitorder="little").sum(1)
    assert (zc == M - ones).all()
    again = eng.max_within(out["yz"], M, N)
    assert (again["a_end"] == out["a_end"]).all() and (again["d_end"] == out["d_end"]).all()
    assert len(again["matches"]) == len(out["matches"]) and (again["matches"] == out["matches"]).all()
    m = out["matches"]
    assert ((m["start"] <= m["end"]) & (m["end"] <= N) & (m["ai"] != m["bi"]) & (m["ai"] < M) & (m["bi"] < M)).all()
    assert (np.diff(m["end"]) >= 0).all()  # callback order: k ascending


@pytest.mark.parametrize("no_sort", [0, 1])
@pytest.mark.parametrize("M,N,W,dens", [(8, 10, 3, 0.3), (33, 40, 5, 0.5), (200, 64, 64, 0.2), (1000, 300, 7, 0.05),
                                        (1025, 120, 16, 0.5), (2752, 38, 4, 0.3), (5000, 200, 148, 0.02),
                                        (12288, 400, 6, 0.3), (12289, 90, 4, 0.3), (9000, 1000, 11, 0.003)])
def test_window_parallel_forced(eng, orc, M, N, W, dens, no_sort):
    """Force the window-parallel path (pass 1 + combine scan + pass 2) at small sizes: results must equal the
    sequential reference sweep bit for bit, including match-record order across window boundaries.  Pass 1 runs
    as the window sort (window_sort.cu) or, with no_sort_pass1, as a sweep; with no_sort the combine scan also runs as
    sweeps over pseudo-sites (windows.cu) instead of the radix sort + range maxima of combine_sort.cu."""
    rng = np.random.default_rng(M * 7 + N)
    hap = (rng.random((M, N)) < dens).astype(np.uint8)
    hap[: M // 3] = hap[0]  # a block of identical haplotypes: ties that survive every window
    x = O.pack_panel(hap)
    a0 = rng.permutation(M).astype(np.int32)
    eng.set_option("windows", W)
    eng.set_option("no_sort_pass1", no_sort)
    eng.set_option("no_sort_combine", no_sort)
    try:
        check_build(eng, orc, x, M, N, matches=M * N < 2000000)
        check_build(eng, orc, x, M, N, a_start=a0, matches=M * N < 2000000)
    finally:
        eng.set_option("windows", 0)
        eng.set_option("no_sort_pass1", 0)
        eng.set_option("no_sort_combine", 0)


@pytest.mark.parametrize("M,N,W", [(2000, 2000, 37), (13312, 1500, 21), (4097, 3000, 148)])
def test_window_parallel_mosaic(eng, orc, M, N, W):
    """Mosaic panels: long shared segments, so tie groups survive several windows and the combine's range maxima span
    whole blocks of positions (sparse-table path of combine_sort.cu)."""
    x = O.generate_panel(M, N, seed=0x5EED0003, L=300, mu=2.0)
    eng.set_option("windows", W)
    try:
        check_build(eng, orc, x, M, N, matches=M <= 4097)
    finally:
        eng.set_option("windows", 0)


def test_small_record_pool_rerun(eng, orc):
    """A deliberately tiny match-record pool must trigger the exact-size rerun, not drop records."""
    M, N = 500, 100
    rng = np.random.default_rng(77)
    hap = (rng.random((M, N)) < 0.3).astype(np.uint8)
    x = O.pack_panel(hap)
    eng.set_option("match_capacity", 1)
    try:
        check_build(eng, orc, x, M, N)
    finally:
        eng.set_option("match_capacity", 0)


def _split_streams(orc, hp, hq):
    M, N = hp.shape
    Q = hq.shape[0]
    pyz = orc.build(O.pack_panel(hp), M, N, with_d=False)[0]
    qyz = orc.build(O.pack_panel(hq), Q, N, with_d=False)[0]
    return pyz, qyz


def test_query_sweep_golden(eng):
    g = np.load(os.path.join(GOLD, "query_sweep.npz"))
    M, Q, N = int(g["M"]), int(g["Q"]), int(g["N"])
    rec, nom = eng.match_sweep(g["pyz"], M, N, g["qyz"], Q)
    assert (rec4(rec) == g["matches"]).all()
    assert (np.stack([nom["query"], nom["allele"], nom["site"]], 1) == g["nomatch"]).all()


@pytest.mark.parametrize("M,Q,N,dp,dq", [(60, 7, 80, 0.03, 0.3), (5, 9, 30, 0.5, 0.5), (200, 1, 50, 0.2, 0.2), (1000, 300, 120, 0.1, 0.1),
                                         (33, 64, 40, 0.5, 0.02), (3000, 1000, 60, 0.05, 0.05)])
def test_query_sweep_random(eng, orc, M, Q, N, dp, dq):
    """matchSequencesSweep parity incl. 'no match' events, a query identical to a panel haplotype, Q=1, Q>M."""
    rng = np.random.default_rng(M + Q)
    hp = (rng.random((M, N)) < dp).astype(np.uint8)
    hq = (rng.random((Q, N)) < dq).astype(np.uint8)
    hq[0] = hp[M // 2]
    pyz, qyz = _split_streams(orc, hp, hq)
    want_rec, want_nom = orc.match_sweep(pyz, M, N, qyz, Q)
    rec, nom = eng.match_sweep(pyz, M, N, qyz, Q)
    assert len(rec) == len(want_rec) and (rec == want_rec).all()
    assert len(nom) == len(want_nom) and (nom == want_nom).all()


def test_query_sweep_mosaic(eng, orc):
    M, Q, N = 4000, 500, 1500
    x = O.generate_panel(M + Q, N, seed=0x5EED0005, L=300, mu=3.0)
    hap = O.unpack_panel(x, M + Q)
    pyz, qyz = _split_streams(orc, hap[:M], hap[M:])
    want_rec, want_nom = orc.match_sweep(pyz, M, N, qyz, Q)
    rec, nom = eng.match_sweep(pyz, M, N, qyz, Q)
    assert len(rec) == len(want_rec) and (rec == want_rec).all() and (nom == want_nom).all()


@pytest.mark.parametrize("M,N,W,dens", [(8, 10, 1, 0.3), (33, 40, 3, 0.5), (1000, 60, 1, 0.2), (3000, 200, 5, 0.05),
                                        (5000, 64, 37, 0.3), (2752, 38, 2, 0.3)])
@pytest.mark.parametrize("no_cluster", [0, 1])
def test_global_memory_kernel_forced(eng, orc, M, N, W, dens, no_cluster):
    """The wide-panel kernel (state in global memory; groups of CTAs as thread-block clusters, or as a cooperative
    grid synchronised through global memory) forced onto narrow panels, with and without window parallelism:
    same bits as the reference sweep."""
    rng = np.random.default_rng(M * 3 + N)
    hap = (rng.random((M, N)) < dens).astype(np.uint8)
    hap[: M // 4] = hap[0]
    x = O.pack_panel(hap)
    a0 = rng.permutation(M).astype(np.int32)
    eng.set_option("force_global", 1)
    eng.set_option("no_cluster", no_cluster)
    eng.set_option("windows", W)
    try:
        check_build(eng, orc, x, M, N)
        check_build(eng, orc, x, M, N, a_start=a0)
    finally:
        eng.set_option("force_global", 0)
        eng.set_option("no_cluster", 0)
        eng.set_option("windows", 0)


@pytest.mark.parametrize("M,N,dens", [(13313, 20, 0.3), (40000, 30, 0.02), (70003, 12, 0.4), (200003, 6, 0.1)])
def test_wide_panels(eng, orc, M, N, dens):
    """M beyond the shared-memory limit (BASELINE configs 3-5 are 1e5..1e6 haplotypes)."""
    rng = np.random.default_rng(M)
    hap = (rng.random((M, N)) < dens).astype(np.uint8)
    check_build(eng, orc, O.pack_panel(hap), M, N, matches=False)


def test_wide_panel_windowed_mosaic(eng, orc):
    """Wide panel, window-parallel, with maxWithin (records from per-CTA streams put into callback order)."""
    M, N = 30000, 1200
    x = O.generate_panel(M, N, seed=0x5EED0003, L=300, mu=40.0)
    check_build(eng, orc, x, M, N)


@pytest.mark.parametrize("M,N,dens,W,no_cluster", [(13313, 48, 0.5, 0, 0), (20000, 40, 0.5, 3, 0), (20000, 40, 0.5, 3, 1),
                                                   (50021, 64, 0.08, 1, 0), (100000, 40, 0.3, 0, 0), (100000, 40, 0.3, 0, 1),
                                                   (600000, 26, 0.5, 0, 0)])
def test_wide_panels_max_within(eng, orc, M, N, dens, W, no_cluster):
    """matchMaximalWithin on panels wider than shared memory: random panels with a block of duplicates (big
    blocks written by the whole CTA, several candidate rounds) through the build and the stored-stream route."""
    rng = np.random.default_rng(M + 7)
    hap = (rng.random((M, N)) < dens).astype(np.uint8)
    hap[100:400] = hap[100]
    hap[M - 200:] = hap[M - 1]
    eng.set_option("windows", W)
    eng.set_option("no_cluster", no_cluster)
    try:
        check_build(eng, orc, O.pack_panel(hap), M, N)
    finally:
        eng.set_option("windows", 0)
        eng.set_option("no_cluster", 0)


@pytest.mark.parametrize("M,N,W,no_sort", [(50000, 2000, 7, 0), (125000, 600, 2, 0), (125000, 600, 2, 1), (140000, 300, 2, 0)])
def test_wide_panel_window_sort(eng, orc, M, N, W, no_sort):
    """Pass 1 of wide panels as a cluster sort (window_sort_wide.cu; ids in distributed shared memory) up to its
    shared-memory limit (~127k haplotypes), the sweep beyond it or with no_sort_pass1: same bits either way."""
    x = O.generate_panel(M, N, seed=0x5EED0031, L=250, mu=M / 50.0)
    eng.set_option("windows", W)
    eng.set_option("no_sort_pass1", no_sort)
    try:
        check_build(eng, orc, x, M, N, matches=False)
    finally:
        eng.set_option("windows", 0)
        eng.set_option("no_sort_pass1", 0)


def test_wide_panel_small_record_pool(eng, orc):
    """The record pool overflows on the first attempt and the sweep is rerun with the exact size."""
    M, N = 16000, 60
    x = O.generate_panel(M, N, seed=0x5EED0013, L=20, mu=300.0)
    eng.set_option("match_capacity", 10000)
    try:
        check_build(eng, orc, x, M, N)
    finally:
        eng.set_option("match_capacity", 0)


def test_blocking_sync_option(eng, orc):
    """Option blocking_sync: the library's host waits sleep on an event (cudaEventBlockingSync) instead of spinning --
    for hosts that run more waiting threads than cores.  Same results."""
    rng = np.random.default_rng(77)
    M, N = 1500, 300
    x = O.pack_panel((rng.random((M, N)) < 0.1).astype(np.uint8))
    eng.set_option("blocking_sync", 1)
    eng.set_option("windows", 7)
    try:
        check_build(eng, orc, x, M, N)
    finally:
        eng.set_option("blocking_sync", 0)
        eng.set_option("windows", 0)


def test_record_pool_overflow_is_paid_once(orc):
    """A panel denser in matches than the first attempt's pool allows (4,200 identical haplotypes: 17.6 M records at N,
    the pool starts at 4,096 per haplotype) reruns pass 2 with the exact size -- and the context remembers the density,
    so that the same call does not pay for the rerun again."""
    M, N = 4200, 8
    x = O.pack_panel(np.zeros((M, N), np.uint8))
    e = pbwt_b200.Engine(0)
    try:
        e.profile(True)
        first = e.build(x, M, N, with_d=True, with_matches=True)
        n1 = e.last_timings()["sweep"][1]
        second = e.build(x, M, N, with_d=True, with_matches=True)
        n2 = e.last_timings()["sweep"][1]
        yz = orc.build(x, M, N, with_d=False)[0]
        rec, n = orc.max_within(yz, M, N)
        assert n == M * (M - 1) and len(first["matches"]) == n and (first["matches"] == rec).all() and (second["matches"] == rec).all()
        assert n1 == 2 * n2 and n2 >= 1  # two launches of pass 2 the first time, one afterwards
    finally:
        e.close()


@pytest.mark.parametrize("force_global", [0, 1])
@pytest.mark.parametrize("M,N,R,W,dens,with_start", [(200, 90, 3, 2, 0.3, False), (1000, 240, 4, 5, 0.05, True), (33, 64, 8, 1, 0.5, False),
                                                      (3000, 400, 2, 7, 0.02, True), (1000, 480, 8, 9, 0.1, True), (200, 300, 6, 12, 0.4, False),
                                                      (500, 4000, 3, 10, 0.03, True), (300, 1500, 2, 8, 0.2, False)])
def test_site_shards_sequential(eng, orc, M, N, R, W, dens, with_start, force_global):
    """The multi-GPU protocol (sites partitioned into R shards: prepare -> gather aggregates -> finish) run
    shard by shard on one GPU: concatenated streams / records and the last shard's a, d equal the sequential sweep."""
    import torch

    rng = np.random.default_rng(M + R)
    hap = (rng.random((M, N)) < dens).astype(np.uint8)
    hap[: M // 3] = hap[1]
    x = O.pack_panel(hap)
    a0 = rng.permutation(M).astype(np.int32) if with_start else None
    yz, a_end, d_end = orc.build(x, M, N, a0, with_d=True)
    rec, n = orc.max_within(yz, M, N, a0)
    flags = pbwt_b200.WANT_YZ | pbwt_b200.WANT_MAXWITHIN
    bounds = [N * r // R for r in range(R + 1)]
    eng.set_option("windows", W)
    eng.set_option("force_global", force_global)  # 1: the wide-panel kernel (what a 1M-haplotype shard runs)
    try:
        panels = [eng.panel_from_host(x[bounds[r]:bounds[r + 1]], M) for r in range(R)]
        shards = [eng.shard_prepare(panels[r], bounds[r], N, flags, a_start=a0) for r in range(R)]
        ai = eng.shard_aggregate_ints(M)
        aggs = torch.zeros((R, ai), dtype=torch.int32, device="cuda")
        for r in range(R):
            shards[r].export(aggs[r].data_ptr())
        outs = []
        for r in range(R):
            res = shards[r].finish(aggs.data_ptr(), r, flags)
            outs.append(res.fetch())
            res.free()
        for s in shards:
            s.free()
        for p in panels:
            p.free()
    finally:
        eng.set_option("windows", 0)
        eng.set_option("force_global", 0)
    assert np.concatenate([o["yz"] for o in outs]).tobytes() == yz.tobytes()
    got = np.concatenate([o["matches"] for o in outs])
    assert len(got) == n and (got == rec).all()
    assert (outs[-1]["a_end"] == a_end).all() and (outs[-1]["d_end"] == d_end).all()
    assert all((o["a_end"] == -1).all() for o in outs[:-1])


def _stats_hist(rec, N):
    """-stats (pbwtMatch.c:130-131): one count per reporter = run of records with the same (ai, end), at its longest length"""
    hist = np.zeros(N + 1, np.int64)
    if len(rec):
        key = rec["ai"].astype(np.int64) * (N + 2) + rec["end"]
        starts = np.flatnonzero(np.r_[True, key[1:] != key[:-1]])
        best = np.maximum.reduceat((rec["end"] - rec["start"]).astype(np.int64), starts)
        np.add.at(hist, best, 1)
    return hist


@pytest.mark.parametrize("M,N,W,dens,force_global", [(33, 40, 1, 0.5, 0), (1000, 300, 7, 0.05, 0), (5000, 200, 37, 0.02, 0),
                                                      (3000, 200, 5, 0.05, 1), (2752, 38, 2, 0.3, 0)])
def test_stats_histogram_consumed_on_device(eng, orc, M, N, W, dens, force_global):
    """WANT_STATS: the matches never leave the device; the histogram equals the one the reference's -stats builds."""
    rng = np.random.default_rng(M + N)
    hap = (rng.random((M, N)) < dens).astype(np.uint8)
    hap[: M // 4] = hap[0]
    yz = orc.build(O.pack_panel(hap), M, N, with_d=False)[0]
    rec, n = orc.max_within(yz, M, N)
    eng.set_option("windows", W)
    eng.set_option("force_global", force_global)
    try:
        hist = eng.max_within_stats(yz, M, N)
        panel = eng.panel_from_host(O.pack_panel(hap), M)
        r = eng.build_device(panel, pbwt_b200.WANT_MAXWITHIN | pbwt_b200.WANT_STATS)
        hist2 = r.fetch_hist()
        assert r.nmatch == 0
        r.free()
        panel.free()
    finally:
        eng.set_option("windows", 0)
        eng.set_option("force_global", 0)
    want = _stats_hist(rec[:n], N)
    assert np.array_equal(hist, want) and np.array_equal(hist2, want)


@pytest.mark.parametrize("opt,val", [("no_fused_pack3", 1), ("pack3_window_capacity", 48), ("pack3_window_capacity", 100000)])
@pytest.mark.parametrize("M,N,W,dens", [(33, 40, 3, 0.5), (1000, 300, 7, 0.05), (1025, 120, 16, 0.5), (5000, 200, 37, 0.02), (12289, 90, 4, 0.3)])
def test_pack3_fused_and_fallback_paths(eng, orc, M, N, W, dens, opt, val):
    """The pack3 stream comes from the encoder fused into the sweep (every warp encodes the runs that start among its
    positions); a window that outgrows its output region (capacity forced tiny here) sends the build back to the
    stand-alone encoder over the sorted columns, which no_fused_pack3 selects directly.  All three must agree with the
    reference byte for byte -- dense columns make warps overflow their 32-byte staging row (the re-encode path)."""
    rng = np.random.default_rng(M + N + W)
    hap = (rng.random((M, N)) < dens).astype(np.uint8)
    hap[:, N // 2] = 0  # a monomorphic column: one run of length M
    hap[M // 2:, N // 3] = 1
    x = O.pack_panel(hap)
    eng.set_option("windows", W)
    eng.set_option(opt, val)
    try:
        check_build(eng, orc, x, M, N, matches=True)
        got = eng.build(x, M, N, with_d=False, with_matches=False)
        assert got["yz"].tobytes() == orc.build(x, M, N, with_d=False)[0].tobytes()
    finally:
        eng.set_option("windows", 0)
        eng.set_option(opt, 0)


@pytest.mark.parametrize("M,N,W,dens,force_global", [(8, 10, 1, 0.3, 0), (1000, 300, 7, 0.05, 0), (2752, 38, 2, 0.3, 0), (5000, 200, 37, 0.02, 0),
                                                      (3000, 200, 5, 0.05, 1)])
def test_grouped_records_expand_to_the_records(eng, orc, M, N, W, dens, force_global):
    """The grouped stream ({ai, start, end, n} + n x bi per run of records that share ai, start, end), built on the device,
    expands to exactly the reference's callback sequence -- through the result object and through the one-shot calls."""
    rng = np.random.default_rng(M + N)
    hap = (rng.random((M, N)) < dens).astype(np.uint8)
    hap[: M // 4] = hap[0]
    x = O.pack_panel(hap)
    yz = orc.build(x, M, N, with_d=False)[0]
    rec, n = orc.max_within(yz, M, N)
    eng.set_option("windows", W)
    eng.set_option("force_global", force_global)
    try:
        panel = eng.panel_from_host(x, M)
        r = eng.build_device(panel, pbwt_b200.WANT_MAXWITHIN)
        g1 = r.fetch_groups()
        r.free()
        panel.free()
        g2 = eng.max_within_grouped(yz, M, N)
    finally:
        eng.set_option("windows", 0)
        eng.set_option("force_global", 0)
    for g in (g1, g2):
        got = pbwt_b200.expand_groups(g)
        assert len(got) == n and (got == rec[:n]).all()
    assert g1.size <= 5 * n  # (a group has at least one record)


def test_smem_limit_without_matches(eng, orc):
    """M between the match kernel's limit (12,544) and the shared-memory limit (13,312): a/d/pack3 only."""
    rng = np.random.default_rng(13)
    M, N = 13312, 10
    hap = (rng.random((M, N)) < 0.3).astype(np.uint8)
    check_build(eng, orc, O.pack_panel(hap), M, N, matches=False)


@pytest.mark.parametrize("M,Q,N,dp,dq", [(60, 7, 80, 0.03, 0.3), (1000, 300, 120, 0.1, 0.1), (3000, 1000, 60, 0.05, 0.05),
                                         (5000, 2500, 50, 0.2, 0.2), (300, 8192, 40, 0.4, 0.4), (2000, 700, 90, 0.004, 0.01)])
@pytest.mark.parametrize("no_cluster", [0, 1])
def test_query_sweep_wide_kernel_forced(eng, orc, M, Q, N, dp, dq, no_cluster):
    """The wide-panel query sweep (state in global memory) forced onto narrow panels."""
    rng = np.random.default_rng(M + Q + 1)
    hp = (rng.random((M, N)) < dp).astype(np.uint8)
    hq = (rng.random((Q, N)) < dq).astype(np.uint8)
    hq[0] = hp[M // 2]
    pyz, qyz = _split_streams(orc, hp, hq)
    want_rec, want_nom = orc.match_sweep(pyz, M, N, qyz, Q)
    eng.set_option("force_global", 1)
    eng.set_option("no_cluster", no_cluster)
    try:
        rec, nom = eng.match_sweep(pyz, M, N, qyz, Q)
    finally:
        eng.set_option("force_global", 0)
        eng.set_option("no_cluster", 0)
    assert len(rec) == len(want_rec) and (rec == want_rec).all()
    assert len(nom) == len(want_nom) and (nom == want_nom).all()


@pytest.mark.parametrize("M,Q,N,W,dp,dq,no_cluster,starts", [
    (60, 7, 80, 5, 0.03, 0.3, 0, False), (60, 7, 80, 80, 0.03, 0.3, 0, True), (1000, 300, 120, 7, 0.1, 0.1, 0, True),
    (1000, 300, 121, 7, 0.1, 0.1, 1, False), (3000, 1000, 60, 4, 0.05, 0.05, 0, False), (300, 8192, 40, 3, 0.4, 0.4, 0, False),
    (2000, 700, 90, 9, 0.004, 0.01, 0, True), (200, 1, 50, 6, 0.2, 0.2, 0, False), (5, 9, 30, 4, 0.5, 0.5, 0, True),
    (2500, 64, 300, 37, 0.02, 0.02, 0, False)])
def test_query_sweep_window_parallel(eng, orc, M, Q, N, W, dp, dq, no_cluster, starts):
    """The window-parallel query sweep forced onto small panels: clamped pass 1 + chain + pass 2 (query_chain.cu) must
    reproduce matchSequencesSweep record for record, 'no match' events included, also from non-identity start orders."""
    rng = np.random.default_rng(M + Q + W)
    hp = (rng.random((M, N)) < dp).astype(np.uint8)
    hq = (rng.random((Q, N)) < dq).astype(np.uint8)
    hq[0] = hp[M // 2]
    for k in rng.choice(N, size=max(1, N // 20), replace=False):  # monomorphic panel columns: 'no match' events
        hp[:, k] = rng.integers(0, 2)
    pa = rng.permutation(M).astype(np.int32) if starts else None
    qa = rng.permutation(Q).astype(np.int32) if starts else None
    pyz = orc.build(O.pack_panel(hp), M, N, a_start=pa, with_d=False)[0]
    qyz = orc.build(O.pack_panel(hq), Q, N, a_start=qa, with_d=False)[0]
    want_rec, want_nom = orc.match_sweep(pyz, M, N, qyz, Q, pa_start=pa, qa_start=qa)
    eng.set_option("force_global", 1)
    eng.set_option("no_cluster", no_cluster)
    eng.set_option("windows", W)
    try:
        rec, nom = eng.match_sweep(pyz, M, N, qyz, Q, pa_start=pa, qa_start=qa)
    finally:
        eng.set_option("force_global", 0)
        eng.set_option("no_cluster", 0)
        eng.set_option("windows", 0)
    assert len(rec) == len(want_rec) and (rec == want_rec).all()
    assert len(nom) == len(want_nom) and (nom == want_nom).all()


def test_query_sweep_window_parallel_mosaic_small_pool(eng, orc):
    """Mosaic panel with matches that span many windows; the record pool overflows on the first attempt (rerun of pass 2
    from the kept chain states)."""
    M, Q, N = 4000, 500, 1500
    x = O.generate_panel(M + Q, N, seed=0x5EED0005, L=300, mu=3.0)
    hap = O.unpack_panel(x, M + Q)
    pyz, qyz = _split_streams(orc, hap[:M], hap[M:])
    want_rec, want_nom = orc.match_sweep(pyz, M, N, qyz, Q)
    eng.set_option("force_global", 1)
    eng.set_option("windows", 23)
    eng.set_option("match_capacity", 1000)
    try:
        rec, nom = eng.match_sweep(pyz, M, N, qyz, Q)
    finally:
        eng.set_option("force_global", 0)
        eng.set_option("windows", 0)
        eng.set_option("match_capacity", 0)
    assert len(rec) == len(want_rec) and (rec == want_rec).all() and (nom == want_nom).all()


def test_query_sweep_wide_panel_default_windows(eng, orc):
    """A panel wider than shared memory and long enough for the default window plan (no options)."""
    M, Q, N = 40000, 300, 3000
    x = O.generate_panel(M + Q, N, seed=0x5EED0025, L=400, mu=6.0)
    hap = O.unpack_panel(x, M + Q)
    pyz, qyz = _split_streams(orc, hap[:M], hap[M:])
    want_rec, want_nom = orc.match_sweep(pyz, M, N, qyz, Q)
    rec, nom = eng.match_sweep(pyz, M, N, qyz, Q)
    assert len(rec) == len(want_rec) and (rec == want_rec).all() and (nom == want_nom).all()


@pytest.mark.parametrize("M,Q,N,R,W,dp,dq,starts", [(60, 7, 80, 3, 2, 0.03, 0.3, False), (1000, 300, 240, 4, 5, 0.1, 0.1, True),
                                                     (33, 64, 64, 8, 1, 0.5, 0.02, False), (2000, 700, 400, 2, 7, 0.004, 0.01, True),
                                                     (200, 1, 90, 3, 4, 0.2, 0.2, False), (500, 40, 2000, 5, 6, 0.03, 0.03, True)])
def test_query_sweep_site_shards_sequential(eng, orc, M, Q, N, R, W, dp, dq, starts):
    """The multi-GPU query sweep (pbwtgpu_qshard_*: sites partitioned into R shards, two all-gathers of aggregates, the
    query state handed from shard to shard) run shard by shard on one GPU: the shards' records and 'no match' events,
    concatenated in site order, are matchSequencesSweep's."""
    import torch

    rng = np.random.default_rng(M + Q + R)
    hp = (rng.random((M, N)) < dp).astype(np.uint8)
    hq = (rng.random((Q, N)) < dq).astype(np.uint8)
    hq[0] = hp[M // 2]
    for k in rng.choice(N, size=max(1, N // 20), replace=False):
        hp[:, k] = rng.integers(0, 2)
    pa = rng.permutation(M).astype(np.int32) if starts else None
    qa = rng.permutation(Q).astype(np.int32) if starts else None
    pyz = orc.build(O.pack_panel(hp), M, N, a_start=pa, with_d=False)[0]
    qyz = orc.build(O.pack_panel(hq), Q, N, a_start=qa, with_d=False)[0]
    want_rec, want_nom = orc.match_sweep(pyz, M, N, qyz, Q, pa_start=pa, qa_start=qa)
    bounds = [N * r // R for r in range(R + 1)]
    eng.set_option("windows", W)
    try:
        shards = [eng.query_shard(pyz, M, N, qyz, Q, bounds[r], bounds[r + 1] - bounds[r], pa_start=pa, qa_start=qa) for r in range(R)]
        qaggs = torch.zeros((R, eng.shard_aggregate_ints(Q)), dtype=torch.int32, device="cuda")
        paggs = torch.zeros((R, eng.shard_aggregate_ints(M)), dtype=torch.int32, device="cuda")
        state = torch.zeros((R + 1, 2, Q), dtype=torch.int32, device="cuda")
        for r in range(R):
            shards[r].export_queries(qaggs[r].data_ptr())
        for r in range(R):
            shards[r].prepare(qaggs.data_ptr(), r, paggs[r].data_ptr())
        outs = []
        for r in range(R):
            shards[r].chain(paggs.data_ptr(), r, state[r].data_ptr() if r else None, state[r + 1].data_ptr())
            outs.append(shards[r].finish())
        for sh in shards:
            sh.free()
    finally:
        eng.set_option("windows", 0)
    rec = np.concatenate([o[0] for o in outs])
    nom = np.concatenate([o[1] for o in outs])
    assert len(rec) == len(want_rec) and (rec == want_rec).all()
    assert len(nom) == len(want_nom) and (nom == want_nom).all()


def test_query_sweep_many_queries_narrow_panel(eng, orc):
    """More queries than the shared-memory kernel holds: routed to the wide kernel without forcing."""
    M, Q, N = 900, 3000, 70
    x = O.generate_panel(M + Q, N, seed=0x5EED0015, L=30, mu=5.0)
    hap = O.unpack_panel(x, M + Q)
    pyz, qyz = _split_streams(orc, hap[:M], hap[M:])
    want_rec, want_nom = orc.match_sweep(pyz, M, N, qyz, Q)
    rec, nom = eng.match_sweep(pyz, M, N, qyz, Q)
    assert len(rec) == len(want_rec) and (rec == want_rec).all() and (nom == want_nom).all()


@pytest.mark.parametrize("M,Q,N,L,mu,no_cluster", [(40000, 300, 400, 200, 6.0, 0), (40000, 300, 400, 200, 6.0, 1),
                                                   (100000, 2000, 150, 50, 2000.0, 0), (100000, 2000, 150, 50, 2000.0, 1)])
def test_query_sweep_wide_panel(eng, orc, M, Q, N, L, mu, no_cluster):
    """BASELINE config 5's shape class: a panel wider than shared memory, queries spread over the group's CTAs."""
    x = O.generate_panel(M + Q, N, seed=0x5EED0005, L=L, mu=mu)
    hap = O.unpack_panel(x, M + Q)
    pyz, qyz = _split_streams(orc, hap[:M], hap[M:])
    want_rec, want_nom = orc.match_sweep(pyz, M, N, qyz, Q)
    eng.set_option("no_cluster", no_cluster)
    try:
        rec, nom = eng.match_sweep(pyz, M, N, qyz, Q)
    finally:
        eng.set_option("no_cluster", 0)
    assert len(rec) == len(want_rec) and (rec == want_rec).all() and (nom == want_nom).all()


# ---- -longWithin L (matchLongWithin2, pbwtMatch.c:85-113) ----------------------------------------------------

def test_long_within_goldens(eng):
    """Records printed by the reference CLI for -longWithin L on the golden panels."""
    g = np.load(os.path.join(GOLD, "long_within.npz"))
    for key in [k for k in g.files if not k.endswith("_md5")]:
        name, L = key.rsplit("_L", 1)
        p = np.load(os.path.join(GOLD, name + ".npz"))
        M, N = int(p["M"]), int(p["N"])
        got = eng.long_within(p["yz"], M, N, int(L))
        assert (got["a_end"] == p["a_end"]).all() and (got["d_end"] == p["d_end"]).all()
        assert len(got["matches"]) == len(g[key]) and (rec4(got["matches"]) == g[key]).all(), key


def check_long(eng, orc, hap, L, a_start=None):
    M, N = hap.shape
    x = O.pack_panel(hap) if N else np.zeros((0, O.words_per_col(M)), np.uint32)
    yz = orc.build(x, M, N, a_start, with_d=False)[0]
    want, n = orc.long_within(yz, M, N, L, a_start)
    got = eng.long_within(yz, M, N, L, a_start=a_start)  # stored-stream route (the CLI's)
    assert len(got["matches"]) == n and (got["matches"] == want).all()
    got2 = eng.build(x, M, N, a_start=a_start, long_L=L)  # build route: gather mode, same scan
    assert got2["yz"].tobytes() == yz.tobytes()
    assert len(got2["matches"]) == n and (got2["matches"] == want).all()
    return n


@pytest.mark.parametrize("M,N,L,W,dens", [(2, 5, 1, 1, 0.5), (8, 0, 3, 1, 0.3), (9, 1, 1, 1, 0.5), (8, 10, 1, 1, 0.3), (33, 40, 2, 3, 0.5), (500, 64, 4, 5, 0.3), (1000, 90, 6, 1, 0.1),
                                          (3000, 300, 40, 9, 0.05), (10000, 200, 14, 0, 0.5), (12544, 100, 13, 0, 0.4)])
def test_long_within_random(eng, orc, M, N, L, W, dens):
    rng = np.random.default_rng(M + L)
    hap = (rng.random((M, N)) < dens).astype(np.uint8)
    if M > 40:
        hap[3:23] = hap[3]  # a block of identical haplotypes (never separated: reported only when both alleles ... none)
        hap[23:33, : N // 2] = hap[3, : N // 2]
    a0 = rng.permutation(M).astype(np.int32) if M % 2 else None
    eng.set_option("windows", W)
    try:
        check_long(eng, orc, hap, L, a0)
    finally:
        eng.set_option("windows", 0)


def test_long_within_mosaic(eng, orc):
    M, N = 4000, 2000
    hap = O.unpack_panel(O.generate_panel(M, N, seed=0x5EED0021, L=200, mu=20.0), M)
    n = check_long(eng, orc, hap, 200)
    assert n > 1000


@pytest.mark.parametrize("M,N,L,W,dens,no_cluster", [(33, 40, 2, 3, 0.5, 0), (3000, 200, 30, 5, 0.05, 0), (3000, 200, 30, 5, 0.05, 1),
                                                     (5000, 64, 12, 37, 0.3, 0)])
def test_long_within_wide_kernel_forced(eng, orc, M, N, L, W, dens, no_cluster):
    rng = np.random.default_rng(M + L + 1)
    hap = (rng.random((M, N)) < dens).astype(np.uint8)
    eng.set_option("force_global", 1)
    eng.set_option("no_cluster", no_cluster)
    eng.set_option("windows", W)
    try:
        check_long(eng, orc, hap, L)
    finally:
        eng.set_option("force_global", 0)
        eng.set_option("no_cluster", 0)
        eng.set_option("windows", 0)


@pytest.mark.parametrize("M,N,L,no_cluster", [(20000, 60, 16, 0), (100000, 48, 18, 0), (100000, 48, 18, 1)])
def test_long_within_wide_panels(eng, orc, M, N, L, no_cluster):
    rng = np.random.default_rng(M + 3)
    hap = (rng.random((M, N)) < 0.5).astype(np.uint8)
    eng.set_option("no_cluster", no_cluster)
    try:
        n = check_long(eng, orc, hap, L)
    finally:
        eng.set_option("no_cluster", 0)
    assert n > 0


# ---- reverse PBWT (pbwtBuildReverse, pbwtCore.c:151-191) -------------------------------------------------------

@pytest.mark.parametrize("name", ["merge1", "merge2", "omni", "mosaic300x400"])
def test_build_reverse_goldens(eng, name):
    g, r = np.load(os.path.join(GOLD, name + ".npz")), np.load(os.path.join(GOLD, "reverse.npz"))
    M, N = int(g["M"]), int(g["N"])
    a_fend, zz, a_rend = eng.build_reverse(g["yz"], M, N)
    assert (a_fend == g["a_end"]).all() and (a_rend == r[name + "_a_rend"]).all()
    assert zz.tobytes() == r[name + "_zz"].tobytes()


@pytest.mark.parametrize("M,N,W,dens,force_global,no_cluster", [
    (2, 3, 1, 0.5, 0, 0), (8, 0, 1, 0.3, 0, 0), (33, 40, 3, 0.5, 0, 0), (1000, 300, 7, 0.05, 0, 0), (5000, 200, 37, 0.3, 0, 0),
    (13312, 30, 2, 0.4, 0, 0), (20000, 24, 1, 0.5, 0, 0), (1000, 300, 7, 0.05, 1, 0), (3000, 120, 5, 0.2, 1, 1), (40000, 40, 3, 0.1, 0, 0),
    (40000, 40, 3, 0.1, 0, 1)])
def test_build_reverse_random(eng, orc, M, N, W, dens, force_global, no_cluster):
    """Forward sweep of the stored stream with every column un-sorted on the device, then the build over the sites
    in reverse order from aFend: aFend, the reverse stream zz and aRend equal the reference's."""
    rng = np.random.default_rng(M + N + 17)
    hap = (rng.random((M, N)) < dens).astype(np.uint8)
    a0 = rng.permutation(M).astype(np.int32) if M % 2 else None
    x = O.pack_panel(hap) if N else np.zeros((0, O.words_per_col(M)), np.uint32)
    yz = orc.build(x, M, N, a0, with_d=False)[0]
    want_fend, want_zz, want_rend = orc.build_reverse(yz, M, N, a0)
    eng.set_option("windows", W)
    eng.set_option("force_global", force_global)
    eng.set_option("no_cluster", no_cluster)
    try:
        a_fend, zz, a_rend = eng.build_reverse(yz, M, N, a_start=a0)
    finally:
        eng.set_option("windows", 0)
        eng.set_option("force_global", 0)
        eng.set_option("no_cluster", 0)
    assert (a_fend == want_fend).all() and (a_rend == want_rend).all()
    assert zz.tobytes() == want_zz.tobytes()


def test_cursor_at_matches_trace(eng, orc):
    """pbwtgpu_cursor_at: (y, a, d, c) of the cursor entering site k of a stored stream = the reference's per-site
    trace (state after k x pbwtCursorForwardsReadAD), at the ends and inside / across window boundaries."""
    M, N = 700, 900
    x = O.generate_panel(M, N, seed=0x5EED0041, L=120, mu=5.0)
    a0 = np.random.default_rng(3).permutation(M).astype(np.int32)
    yz = orc.build(x, M, N, a0, with_d=False)[0]
    a_tr, d_tr, y_tr, c_tr = orc.trace(x, M, N, a0)
    for k in (0, 1, 255, 256, 257, 600, N - 1, N):
        y, a, d, c = eng.cursor_at(yz, M, N, k, a_start=a0)
        assert (a == a_tr[k]).all() and (d == d_tr[k]).all(), k
        if k < N:
            assert (y == y_tr[k]).all() and c == c_tr[k], k


def test_error_codes(eng, orc):
    """Bad arguments and unsupported shapes come back as error codes with a message (the host shim turns them into the
    reference's die()); nothing falls back to a CPU path."""
    M, N = 64, 20
    hap = (np.random.default_rng(1).random((M, N)) < 0.3).astype(np.uint8)
    yz = orc.build(O.pack_panel(hap), M, N, with_d=False)[0]
    with pytest.raises(pbwt_b200.PbwtGpuError, match="invalid argument"):
        eng.long_within(yz, M, N, 0)  # L must be >= 1 (pbwtMatch.c:153)
    with pytest.raises(pbwt_b200.PbwtGpuError, match="invalid argument"):
        eng.cursor_at(yz, M, N, N + 1)
    with pytest.raises(pbwt_b200.PbwtGpuError, match="decodes to"):
        eng.max_within(yz[:-3], M, N)  # truncated stream: the reference's isCheck assertion (pbwtCore.c:301-302)
    qyz = orc.build(O.pack_panel(np.zeros((9000, N), np.uint8)), 9000, N, with_d=False)[0]
    with pytest.raises(pbwt_b200.PbwtGpuError, match="unsupported shape"):
        eng.match_sweep(yz, M, N, qyz, 9000)  # more than 8,192 queries per call
    with pytest.raises(pbwt_b200.PbwtGpuError, match="unknown option"):
        eng.set_option("no_such_option", 1)
    # the context is still usable afterwards
    got = eng.max_within(yz, M, N)
    rec, n = orc.max_within(yz, M, N)
    assert len(got["matches"]) == n and (got["matches"] == rec).all()
