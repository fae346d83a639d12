"""Second, independent oracle for the query sweep: the reference's -matchIndexed (matchSequencesIndexed,
pbwtMatch.c:254-340 -- Algorithm 5 with stored indices, one query at a time) reports the same set of maximal matches as
-matchDynamic (matchSequencesSweep, pbwtMatch.c:363-443), in a different order.  The CPU case pins the oracle port's
record multiset to it; the GPU case does the same for the device's records (window-parallel path included)."""
import collections
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "pbwt_ref")


def write_pbwt(path, yz, a_end, M, N):
    with open(path, "wb") as fh:  # .pbwt v3, pbwtIO.c:33-57
        fh.write(b"PBW3" + np.array([M, N], "<i4").tobytes() + np.arange(M, dtype="<i4").tobytes() + np.asarray(a_end, "<i4").tobytes() +
                 np.array([len(yz)], "<i8").tobytes() + b"    " + np.asarray(yz, np.uint8).tobytes())


def indexed_matches(tmp_path, pyz, pa, M, qyz, qa, Q, N):
    """multiset of (query, panel haplotype, start, end) printed by `pbwt_ref -read panel -matchIndexed queries`"""
    write_pbwt(tmp_path / "p.pbwt", pyz, pa, M, N)
    write_pbwt(tmp_path / "q.pbwt", qyz, qa, Q, N)
    out = subprocess.run([REF_CLI, "-read", str(tmp_path / "p.pbwt"), "-matchIndexed", str(tmp_path / "q.pbwt")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-400:]
    c = collections.Counter()
    for line in out.stdout.splitlines():
        f = line.split("\t")
        if f[0] == "MATCH":
            c[(int(f[1]), int(f[2]), int(f[3]), int(f[4]))] += 1
    return c


def sweep_multiset(rec):
    nz = rec[rec["start"] < rec["end"]]  # the CLI printer drops zero-length matches (pbwtMatch.c:46-49)
    return collections.Counter(zip(nz["ai"].tolist(), nz["bi"].tolist(), nz["start"].tolist(), nz["end"].tolist()))


def make_case(M, Q, N, seed):
    x = O.generate_panel(M + Q, N, seed=seed, L=60, mu=4.0)
    hap = O.unpack_panel(x, M + Q)
    hap[0, :] = 0
    hap[1, :] = 1  # both alleles present in every panel column: no "no match" sites (matchIndexed does not handle them)
    orc = O.port()
    pyz, pa, _ = orc.build(O.pack_panel(hap[:M]), M, N, with_d=False)
    qyz, qa, _ = orc.build(O.pack_panel(hap[M:]), Q, N, with_d=False)
    return pyz, pa, qyz, qa


@pytest.mark.skipif(not os.path.exists(REF_CLI), reason="oracle/_ref/pbwt_ref not built (needs the reference's sources)")
@pytest.mark.parametrize("M,Q,N,seed", [(300, 40, 400, 1), (1000, 25, 250, 2)])
def test_port_records_equal_match_indexed_multiset(tmp_path, M, Q, N, seed):
    pyz, pa, qyz, qa = make_case(M, Q, N, seed)
    rec, nom = O.port().match_sweep(pyz, M, N, qyz, Q)
    assert len(nom) == 0
    assert sweep_multiset(rec) == indexed_matches(tmp_path, pyz, pa, M, qyz, qa, Q, N)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(REF_CLI), reason="oracle/_ref/pbwt_ref not built (needs the reference's sources)")
@pytest.mark.parametrize("M,Q,N,seed,windows", [(300, 40, 400, 1, 0), (1000, 25, 250, 2, 6), (3000, 60, 700, 3, 11)])
def test_gpu_records_equal_match_indexed_multiset(tmp_path, M, Q, N, seed, windows):
    import pbwt_b200 as P

    pyz, pa, qyz, qa = make_case(M, Q, N, seed)
    eng = P.Engine(0)
    try:
        if windows:
            eng.set_option("force_global", 1)
            eng.set_option("windows", windows)
        rec, nom = eng.match_sweep(pyz, M, N, qyz, Q)
    finally:
        eng.close()
    assert len(nom) == 0
    assert sweep_multiset(rec) == indexed_matches(tmp_path, pyz, pa, M, qyz, qa, Q, N)
