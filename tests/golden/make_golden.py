#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container (needs /root/reference and `make -C oracle ref`):

    python tests/golden/make_golden.py

Everything written here is an OUTPUT of the reference (its CLI `oracle/_ref/pbwt_ref` or its
library `oracle/_ref/libpbwtref.so`) on inputs taken from the reference's own test directory
or from the seeded synthetic generator; the fixtures travel to the GPU box, the reference
does not.  Inputs are stored as bit-packed panels (our own format), not as copies of the
reference's test files.
"""
import hashlib
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402

REF = os.environ.get("PBWT_REF", "/root/reference")
CLI = os.path.join(ROOT, "oracle", "_ref", "pbwt_ref")


def read_vcfq(path):
    """chrom pos ref alt gt gt ... -> hap uint8 [M, N] (parseVcfqLine, pbwtIO.c:511-554)."""
    cols = []
    for line in open(path):
        f = line.rstrip("\n").split("\t")
        cols.append([int(ch) for g in f[4:] for ch in g if ch in "01"])
    return np.array(cols, np.uint8).T.copy()


def read_vcf_gt(path):
    """Plain-text VCF, biallelic, no missing (enough for test/OMNI.vcf; pbwtHtslib.c:52-172)."""
    cols = []
    for line in open(path):
        if line.startswith("#"):
            continue
        f = line.rstrip("\n").split("\t")
        cols.append([int(ch) for g in f[9:] for ch in g.split(":")[0] if ch in "01"])
    return np.array(cols, np.uint8).T.copy()


def write_vcfq(hap, path):
    M, N = hap.shape
    with open(path, "w") as fh:
        for k in range(N):
            fh.write("20\t%d\tA\tC\t" % (1000 + k) + "\t".join(str(int(v)) for v in hap[:, k]) + "\n")


def cli_pbwt_and_matches(hap):
    """Run the reference CLI: -readVcfq -> -write, then -read -> -maxWithin (stdout)."""
    with tempfile.TemporaryDirectory() as td:
        vq, pb = os.path.join(td, "x.vcfq"), os.path.join(td, "x.pbwt")
        write_vcfq(hap, vq)
        subprocess.run([CLI, "-readVcfq", vq, "-write", pb], check=True, capture_output=True)
        raw = open(pb, "rb").read()
        out = subprocess.run([CLI, "-read", pb, "-maxWithin"], check=True, capture_output=True).stdout
    return raw, out


def parse_pbwt(raw):
    """PBW3 layout, pbwtIO.c:33-57."""
    assert raw[:4] == b"PBW3"
    M, N = np.frombuffer(raw, "<i4", 2, 4)
    off = 12
    a_start = np.frombuffer(raw, "<i4", M, off); off += 4 * M
    a_end = np.frombuffer(raw, "<i4", M, off); off += 4 * M
    nz = int(np.frombuffer(raw, "<i8", 1, off)[0]); off += 8 + 4
    yz = np.frombuffer(raw, np.uint8, nz, off)
    return int(M), int(N), a_start.copy(), a_end.copy(), yz.copy()


def stdout_records(out):
    rows = [ln.split("\t")[1:5] for ln in out.decode().splitlines() if ln.startswith("MATCH")]
    return np.array(rows, np.int32).reshape(-1, 4)


def panel_fixture(name, hap, with_trace):
    ref = O.reference()
    M, N = hap.shape
    x = O.pack_panel(hap)
    raw, out = cli_pbwt_and_matches(hap)
    M2, N2, a_start, a_end, yz = parse_pbwt(raw)
    assert (M2, N2) == (M, N)
    lib_yz, lib_a, lib_d = ref.build(x, M, N, with_d=True)
    assert (lib_yz == yz).all() and (lib_a == a_end).all()
    rec, n = ref.max_within(yz, M, N)
    rec4 = np.stack([rec["ai"], rec["bi"], rec["start"], rec["end"]], 1)
    printed = rec4[rec4[:, 2] != rec4[:, 3]]  # reportMatch drops zero-length, pbwtMatch.c:48
    assert (printed == stdout_records(out)).all()
    d = dict(M=M, N=N, xbits=x, yz=yz, a_end=a_end, d_end=lib_d, matches=rec4,
             pbwt_md5=hashlib.md5(raw).hexdigest(), maxwithin_stdout_md5=hashlib.md5(out).hexdigest(),
             maxwithin_stdout_lines=len(out.splitlines()))
    if with_trace:
        a_tr, d_tr, y_tr, c_tr = ref.trace(x, M, N)
        d.update(a_trace=a_tr, d_trace=d_tr, y_trace=y_tr, c_trace=c_tr)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print(name, "M", M, "N", N, "yz", yz.size, "records", n, "stdout lines", d["maxwithin_stdout_lines"],
          d["pbwt_md5"], d["maxwithin_stdout_md5"])


def pack3_fixture():
    """Appendix B threshold vectors: column 1, 0 x n, 1, 1 through the reference's pack3."""
    ref = O.reference()
    ns = [1, 63, 64, 65, 127, 128, 2047, 2048, 2049, 4095, 63487, 63488, 63489, 65536, 70000, 126976, 127040, 200000]
    out = {}
    for n in ns:
        y = np.concatenate([[1], np.zeros(n, np.uint8), [1, 1]]).astype(np.uint8)
        out["n%d" % n] = ref.pack3(y)
    np.savez_compressed(os.path.join(HERE, "pack3_vectors.npz"), **out)
    print("pack3", {k: v.tobytes().hex() for k, v in out.items()})


def write_pbwt(path, M, N, a_start, a_end, yz):
    """PBW3 layout, pbwtIO.c:33-57."""
    with open(path, "wb") as fh:
        fh.write(b"PBW3")
        fh.write(np.array([M, N], "<i4").tobytes())
        fh.write(np.asarray(a_start, "<i4").tobytes())
        fh.write(np.asarray(a_end, "<i4").tobytes())
        fh.write(np.array([len(yz)], "<i8").tobytes())
        fh.write(b"    ")
        fh.write(np.asarray(yz, np.uint8).tobytes())


def query_fixture():
    """matchSequencesSweep on a seeded mosaic split 550 panel / 50 query haplotypes x 500 sites."""
    ref = O.reference()
    M, Q, N = 550, 50, 500
    x = O.generate_panel(M + Q, N, seed=0x5EED0005, L=100, mu=4.0)
    hap = O.unpack_panel(x, M + Q)
    px, qx = O.pack_panel(hap[:M]), O.pack_panel(hap[M:])
    pyz, pa, _ = ref.build(px, M, N, with_d=False)
    qyz, qa, _ = ref.build(qx, Q, N, with_d=False)
    rec, nom = ref.match_sweep(pyz, M, N, qyz, Q)
    rec4 = np.stack([rec["ai"], rec["bi"], rec["start"], rec["end"]], 1)
    nom3 = np.stack([nom["query"], nom["allele"], nom["site"]], 1)
    # the CLI route: pbwt -read panel.pbwt -matchDynamic query.pbwt (pbwtMain.c:412) -> stdout + log
    with tempfile.TemporaryDirectory() as td:
        pp, qp, lg = os.path.join(td, "p.pbwt"), os.path.join(td, "q.pbwt"), os.path.join(td, "log")
        write_pbwt(pp, M, N, np.arange(M), pa, pyz)
        write_pbwt(qp, Q, N, np.arange(Q), qa, qyz)
        out = subprocess.run([CLI, "-log", lg, "-read", pp, "-matchDynamic", qp], check=True, capture_output=True).stdout
        log = [ln for ln in open(lg).read().splitlines() if ln.startswith("no match") or ln.startswith("Average number")]
    printed = rec4[rec4[:, 2] != rec4[:, 3]]
    assert (printed == stdout_records(out)).all()
    np.savez_compressed(os.path.join(HERE, "query_sweep.npz"), M=M, Q=Q, N=N, pxbits=px, qxbits=qx,
                        pyz=pyz, qyz=qyz, pa_end=pa, qa_end=qa, matches=rec4, nomatch=nom3,
                        matchdynamic_stdout_md5=hashlib.md5(out).hexdigest(), matchdynamic_log="\n".join(log))
    print("query_sweep records", len(rec4), "nomatch", len(nom3), hashlib.md5(out).hexdigest(), log[-1])


def long_within_fixture():
    """pbwt -read X.pbwt -longWithin L (matchLongWithin2, pbwtMatch.c:85-113) through the reference CLI, on the
    panels of the other fixtures: records (all have length >= L > 0, so the printer drops none)."""
    out = {}
    for name, Ls in [("merge1", (1, 2, 4)), ("merge2", (1, 3)), ("omni", (20, 30)), ("mosaic300x400", (40, 100))]:
        g = np.load(os.path.join(HERE, name + ".npz"))
        M, N = int(g["M"]), int(g["N"])
        with tempfile.TemporaryDirectory() as td:
            pb = os.path.join(td, "x.pbwt")
            write_pbwt(pb, M, N, np.arange(M), g["a_end"], g["yz"])
            for L in Ls:
                txt = subprocess.run([CLI, "-read", pb, "-longWithin", str(L)], check=True, capture_output=True).stdout
                rec = stdout_records(txt)
                lib, n = O.reference().long_within(g["yz"], M, N, L)
                assert n == len(rec) and (np.stack([lib["ai"], lib["bi"], lib["start"], lib["end"]], 1) == rec).all()
                out[f"{name}_L{L}"] = rec
                out[f"{name}_L{L}_md5"] = hashlib.md5(txt).hexdigest()
                print("longWithin", name, "L", L, "records", len(rec), out[f"{name}_L{L}_md5"])
    np.savez_compressed(os.path.join(HERE, "long_within.npz"), **out)


def reverse_fixture():
    """pbwt -read X.pbwt -buildReverse -writeReverse R.pbwt through the reference CLI (pbwtBuildReverse, pbwtCore.c:151-191):
    the reverse file's md5 and its parsed contents (aRstart = aFend, aRend, zz)."""
    out = {}
    for name in ("merge1", "merge2", "omni", "mosaic300x400"):
        g = np.load(os.path.join(HERE, name + ".npz"))
        M, N = int(g["M"]), int(g["N"])
        with tempfile.TemporaryDirectory() as td:
            pb, rv = os.path.join(td, "x.pbwt"), os.path.join(td, "r.pbwt")
            write_pbwt(pb, M, N, np.arange(M), g["a_end"], g["yz"])
            subprocess.run([CLI, "-read", pb, "-buildReverse", "-writeReverse", rv], check=True, capture_output=True)
            raw = open(rv, "rb").read()
        M2, N2, a_rstart, a_rend, zz = parse_pbwt(raw)
        lib_fend, lib_zz, lib_rend = O.reference().build_reverse(g["yz"], M, N)
        assert (a_rstart == g["a_end"]).all() and (lib_fend == a_rstart).all() and (lib_rend == a_rend).all()
        assert lib_zz.tobytes() == zz.tobytes()
        out[name + "_zz"], out[name + "_a_rend"], out[name + "_md5"] = zz, a_rend, hashlib.md5(raw).hexdigest()
        print("reverse", name, "zz", zz.size, out[name + "_md5"])
    np.savez_compressed(os.path.join(HERE, "reverse.npz"), **out)


def main():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref", "port"], check=True, capture_output=True)
    panel_fixture("merge1", read_vcfq(os.path.join(REF, "test", "merge.1.tab")), with_trace=True)
    panel_fixture("merge2", read_vcfq(os.path.join(REF, "test", "merge.2.tab")), with_trace=True)
    panel_fixture("omni", read_vcf_gt(os.path.join(REF, "test", "OMNI.vcf")), with_trace=False)
    x = O.generate_panel(300, 400, seed=0x5EED0001, L=80, mu=3.0)
    panel_fixture("mosaic300x400", O.unpack_panel(x, 300), with_trace=False)
    pack3_fixture()
    query_fixture()
    long_within_fixture()
    reverse_fixture()


if __name__ == "__main__":
    main()
