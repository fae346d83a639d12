"""CPU tests: pin the oracle port (oracle/pbwt_oracle.c) against the reference's golden
vectors (tests/golden/, produced by tests/golden/make_golden.py from the unmodified reference)
and, where the compiled reference is present (oracle/_ref/), against the reference itself."""
import os

import numpy as np
import pytest

import oracle_lib as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def rec4(rec):
    return np.stack([rec["ai"], rec["bi"], rec["start"], rec["end"]], 1)


@pytest.fixture(scope="module")
def orc():
    return O.port()


def impls():
    out = [("port", O.port())]
    if O.reference() is not None:
        out.append(("reference", O.reference()))
    return out


@pytest.mark.parametrize("name", ["merge1", "merge2", "omni", "mosaic300x400"])
def test_panel_goldens(name):
    g = load(name)
    M, N = int(g["M"]), int(g["N"])
    for label, imp in impls():
        yz, a_end, d_end = imp.build(g["xbits"], M, N, with_d=True)
        assert yz.tobytes() == g["yz"].tobytes(), label
        assert (a_end == g["a_end"]).all() and (d_end == g["d_end"]).all(), label
        yz2, a2, _ = imp.build(g["xbits"], M, N, with_d=False)
        assert yz2.tobytes() == g["yz"].tobytes() and (a2 == g["a_end"]).all(), label
        rec, n = imp.max_within(g["yz"], M, N)
        assert n == len(g["matches"]) and (rec4(rec) == g["matches"]).all(), label
        a3, d3 = imp.sweep_ad(g["yz"], M, N)
        assert (a3 == g["a_end"]).all() and (d3 == g["d_end"]).all(), label
        assert (imp.sweep_a(g["yz"], M, N) == g["a_end"]).all(), label
        if "a_trace" in g:
            a_tr, d_tr, y_tr, c_tr = imp.trace(g["xbits"], M, N)
            assert (a_tr == g["a_trace"]).all() and (d_tr == g["d_trace"]).all(), label
            assert (y_tr == g["y_trace"]).all() and (c_tr == g["c_trace"]).all(), label


def test_long_within_goldens():
    """-longWithin L (matchLongWithin2, pbwtMatch.c:85-113): records printed by the reference CLI."""
    g = load("long_within")
    for key in [k for k in g.files if not k.endswith("_md5")]:
        name, L = key.rsplit("_L", 1)
        p = load(name)
        M, N = int(p["M"]), int(p["N"])
        for label, imp in impls():
            rec, n = imp.long_within(p["yz"], M, N, int(L))
            assert n == len(g[key]) and (rec4(rec) == g[key]).all(), (label, key)


def test_long_within_port_equals_reference_random():
    ref = O.reference()
    if ref is None:
        pytest.skip("compiled reference not present")
    port = O.port()
    rng = np.random.default_rng(5)
    for it in range(40):
        M, N = int(rng.integers(2, 200)), int(rng.integers(0, 60))
        hap = (rng.random((M, N)) < rng.choice([0.02, 0.2, 0.5])).astype(np.uint8)
        if M > 8:
            hap[: M // 3] = hap[0]
        a0 = rng.permutation(M).astype(np.int32) if it % 2 else None
        x = O.pack_panel(hap) if N else np.zeros((0, O.words_per_col(M)), np.uint32)
        yz = port.build(x, M, N, a0, with_d=False)[0]
        for T in (1, 3, 10):
            r1, n1 = port.long_within(yz, M, N, T, a0)
            r2, n2 = ref.long_within(yz, M, N, T, a0)
            assert n1 == n2 and (r1 == r2).all(), (M, N, T)


@pytest.mark.parametrize("name", ["merge1", "merge2", "omni", "mosaic300x400"])
def test_reverse_goldens(name):
    """pbwtBuildReverse (pbwtCore.c:151-191): the reverse stream and orders written by the reference CLI."""
    g, r = load(name), load("reverse")
    M, N = int(g["M"]), int(g["N"])
    for label, imp in impls():
        a_fend, zz, a_rend = imp.build_reverse(g["yz"], M, N)
        assert (a_fend == g["a_end"]).all() and (a_rend == r[name + "_a_rend"]).all(), label
        assert zz.tobytes() == r[name + "_zz"].tobytes(), label


def test_merge1_known_answers():
    """the 34-byte stream, aFend and the first/last MATCH lines."""
    g = load("merge1")
    assert g["yz"].tobytes().hex() == ("8104810281038103810281048107810701810602810581028104810481028102" "8104")
    assert g["a_end"].tolist() == [3, 4, 0, 7, 1, 6, 2, 5]
    assert g["d_end"].tolist() == [11, 5, 8, 3, 9, 7, 10, 6, 11]
    assert g["matches"][0].tolist() == [3, 4, 0, 3] and g["matches"][-1].tolist() == [5, 2, 6, 10]
    assert str(g["pbwt_md5"]) == "81643097d3a122d8e3277214a75f9773"


def test_omni_known_answers():
    g = load("omni")
    assert str(g["pbwt_md5"]) == "d2753c044392ec4f474c7061cff773e9"
    assert str(g["maxwithin_stdout_md5"]) == "b1412a0e3057b3e92ade2cee7c4a4824"
    assert int(g["maxwithin_stdout_lines"]) == 21883 and g["yz"].size == 25156


def test_pack3_vectors():
    g = load("pack3_vectors")
    for label, imp in impls():
        for key in g.files:
            n = int(key[1:])
            y = np.concatenate([[1], np.zeros(n, np.uint8), [1, 1]]).astype(np.uint8)
            enc = imp.pack3(y)
            assert enc.tobytes() == g[key].tobytes(), (label, key)
            dec, used, n0 = imp.unpack3(enc, y.size)
            assert (dec == y).all() and used == enc.size and n0 == n, (label, key)


def test_query_sweep_golden():
    g = load("query_sweep")
    M, Q, N = int(g["M"]), int(g["Q"]), int(g["N"])
    for label, imp in impls():
        pyz = imp.build(g["pxbits"], M, N, with_d=False)[0]
        qyz = imp.build(g["qxbits"], Q, N, with_d=False)[0]
        assert pyz.tobytes() == g["pyz"].tobytes() and qyz.tobytes() == g["qyz"].tobytes()
        rec, nom = imp.match_sweep(pyz, M, N, qyz, Q)
        assert (rec4(rec) == g["matches"]).all(), label
        assert (np.stack([nom["query"], nom["allele"], nom["site"]], 1) == g["nomatch"]).all(), label


@pytest.mark.skipif(O.reference() is None, reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("M,N,dens,seed", [(2, 3, 0.5, 1), (3, 1, 0.5, 2), (31, 40, 0.3, 3), (33, 17, 0.05, 4),
                                           (64, 64, 0.5, 5), (129, 70, 0.02, 6), (1000, 60, 0.1, 7), (70003, 3, 0.4, 8)])
def test_port_vs_reference_random(M, N, dens, seed):
    rng = np.random.default_rng(seed)
    hap = (rng.random((M, N)) < dens).astype(np.uint8)
    if N > 2:
        hap[:, 1] = 0  # monomorphic columns
        hap[:, 2] = 1
    x = O.pack_panel(hap)
    a0 = rng.permutation(M).astype(np.int32)
    ref, orc = O.reference(), O.port()
    for a_start in (None, a0):
        r, o = ref.build(x, M, N, a_start, True), orc.build(x, M, N, a_start, True)
        assert all((u == v).all() for u, v in zip(r, o))
        yz = r[0]
        if M <= 2000:
            (mr, nr), (mo, no) = ref.max_within(yz, M, N, a_start), orc.max_within(yz, M, N, a_start)
            assert nr == no and (mr == mo).all()
        assert all((u == v).all() for u, v in zip(ref.sweep_ad(yz, M, N, a_start), orc.sweep_ad(yz, M, N, a_start)))
        assert (ref.sweep_a(yz, M, N, a_start) == orc.sweep_a(yz, M, N, a_start)).all()


@pytest.mark.skipif(O.reference() is None, reason="compiled reference (oracle/_ref) not present")
def test_port_vs_reference_query_sweep_edge_cases():
    rng = np.random.default_rng(11)
    ref, orc = O.reference(), O.port()
    for M, Q, N, dp, dq in [(60, 7, 80, 0.03, 0.3), (5, 9, 30, 0.5, 0.5), (200, 1, 50, 0.2, 0.2)]:
        hp = (rng.random((M, N)) < dp).astype(np.uint8)
        hq = (rng.random((Q, N)) < dq).astype(np.uint8)
        hq[0] = hp[M // 2]  # a query identical to a panel haplotype
        pyz = ref.build(O.pack_panel(hp), M, N, with_d=False)[0]
        qyz = ref.build(O.pack_panel(hq), Q, N, with_d=False)[0]
        (rr, rn), (oo, on) = ref.match_sweep(pyz, M, N, qyz, Q), orc.match_sweep(pyz, M, N, qyz, Q)
        assert len(rr) == len(oo) and (rr == oo).all() and len(rn) == len(on) and (rn == on).all()


def test_generator_is_deterministic_and_sane():
    x1 = O.generate_panel(500, 300, seed=42, L=50, mu=3.0)
    x2 = O.generate_panel(500, 300, seed=42, L=50, mu=3.0)
    assert (x1 == x2).all()
    hap = O.unpack_panel(x1, 500)
    assert hap[0].sum() == 0 and 0 < hap.mean() < 0.5
