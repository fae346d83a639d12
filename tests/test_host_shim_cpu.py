"""CPU tests of the HOST side of the drop-in (pbwt_b200/host/pbwt_gpu_shim.c): the reference's own pbwtMain.c /
pbwtIO.c / callers linked with the shim exactly as in pbwt_b200/host/Makefile, but against tests/host_mock/
mock_pbwtgpu.c -- a stand-in for libpbwtgpu.so answered by the CPU oracle port -- so that what is under test without
a GPU is the shim's own logic: record -> callback order, the interleaving of "no match" log lines, the summary lines,
-stats histograms, Array / aFend ownership, and the readers' lazy cursor with its batches.  Expected outputs are the
reference's (golden md5s recorded from the unmodified reference CLI, tests/golden/).  The same commands run against
the real device library in tests/test_gpu_cli.py."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from test_gpu_cli import GOLD, ROOT, write_pbwt

MOCK_DIR = os.path.join(ROOT, "tests", "host_mock")
HOST_OBJ = os.path.join(ROOT, "pbwt_b200", "host", "_build", "pbwt_gpu_shim.o")
CLI = os.environ.get("PBWT_HOSTCHECK", os.path.join(MOCK_DIR, "_build", "pbwt_hostcheck"))  # override: e.g. an ASan/UBSan build of the same link
SELFCHECK = os.path.join(MOCK_DIR, "_build", "shim_selfcheck")


@pytest.fixture(scope="module", autouse=True)
def hostcheck():
    if not os.path.exists(HOST_OBJ):
        pytest.skip("host shim objects not built (they need the reference sources at build time)")
    O.port()  # makes sure oracle/libpbwt_oracle.so exists
    r = subprocess.run(["make", "-C", MOCK_DIR], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def golden_pbwt(tmp_path, name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    M, N = int(g["M"]), int(g["N"])
    pb = tmp_path / "x.pbwt"
    write_pbwt(pb, M, N, np.arange(M), g["a_end"], g["yz"])
    return g, pb


def run(*args, env=None):
    return subprocess.run([CLI, *map(str, args)], check=True, capture_output=True, env=env).stdout


@pytest.mark.parametrize("name", ["merge1", "omni", "mosaic300x400"])
def test_maxwithin_stdout(tmp_path, name):
    g, pb = golden_pbwt(tmp_path, name)
    out = run("-read", pb, "-maxWithin")
    assert len(out.splitlines()) == int(g["maxwithin_stdout_lines"])
    assert hashlib.md5(out).hexdigest() == str(g["maxwithin_stdout_md5"])


@pytest.mark.parametrize("name,L", [("merge1", 2), ("omni", 20), ("mosaic300x400", 40)])
def test_longwithin_stdout(tmp_path, name, L):
    _, pb = golden_pbwt(tmp_path, name)
    lw = np.load(os.path.join(GOLD, "long_within.npz"))
    out = run("-read", pb, "-longWithin", L)
    assert hashlib.md5(out).hexdigest() == str(lw[f"{name}_L{L}_md5"])


@pytest.mark.parametrize("name", ["merge1", "omni", "mosaic300x400"])
def test_build_reverse_file(tmp_path, name):
    _, pb = golden_pbwt(tmp_path, name)
    r = np.load(os.path.join(GOLD, "reverse.npz"))
    rv = tmp_path / "r.pbwt"
    run("-read", pb, "-buildReverse", "-writeReverse", rv)
    assert hashlib.md5(open(rv, "rb").read()).hexdigest() == str(r[name + "_md5"])


@pytest.mark.parametrize("name,md5,log", [
    ("merge1", "8d1029f275f8c490f3c50923db651419", ["Average 2.9 matches per sample", "Average length 4.9"]),
    ("omni", "3efa4cec8b933c7a984c8b96cae7e035", ["Average 4.0 matches per sample", "Average length 20.7"]),
    ("mosaic300x400", "a6e0ff75052b0c53c1e193a479eb6965", ["Average 12.4 matches per sample", "Average length 69.8"])])
def test_stats_maxwithin(tmp_path, name, md5, log):
    _, pb = golden_pbwt(tmp_path, name)
    lg = tmp_path / "log"
    out = run("-log", lg, "-stats", "-read", pb, "-maxWithin")
    assert hashlib.md5(out).hexdigest() == md5
    assert [ln for ln in open(lg).read().splitlines() if ln.startswith("Average")] == log


def test_matchdynamic_stdout_and_log(tmp_path):
    g = np.load(os.path.join(GOLD, "query_sweep.npz"))
    M, Q, N = int(g["M"]), int(g["Q"]), int(g["N"])
    pp, qp, lg = tmp_path / "p.pbwt", tmp_path / "q.pbwt", tmp_path / "log"
    write_pbwt(pp, M, N, np.arange(M), g["pa_end"], g["pyz"])
    write_pbwt(qp, Q, N, np.arange(Q), g["qa_end"], g["qyz"])
    out = run("-log", lg, "-read", pp, "-matchDynamic", qp)
    assert hashlib.md5(out).hexdigest() == str(g["matchdynamic_stdout_md5"])
    log = [ln for ln in open(lg).read().splitlines() if ln.startswith("no match") or ln.startswith("Average number")]
    assert "\n".join(log) == str(g["matchdynamic_log"])


@pytest.mark.parametrize("name,batch", [("merge1", 1), ("merge1", 3), ("omni", 7), ("omni", 0), ("mosaic300x400", 128)])
def test_reader_batched_ingest(tmp_path, name, batch):
    """The readers' per-site cursor step through the shim's lazy cursor: several batches, order carried between them."""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    M, N = int(g["M"]), int(g["N"])
    hap = O.unpack_panel(g["xbits"], M)
    vq, pb = tmp_path / "x.vcfq", tmp_path / "x.pbwt"
    with open(vq, "w") as fh:  # parseVcfqLine, pbwtIO.c:511-554
        for k in range(N):
            fh.write("20\t%d\tA\tC\t" % (1000 + k) + "\t".join(str(int(v)) for v in hap[:, k]) + "\n")
    env = dict(os.environ)
    if batch:
        env["PBWTGPU_READER_BATCH"] = str(batch)
    run("-readVcfq", vq, "-write", pb, env=env)
    assert hashlib.md5(open(pb, "rb").read()).hexdigest() == str(g["pbwt_md5"])
    out = run("-readVcfq", vq, "-maxWithin", env=env)
    assert hashlib.md5(out).hexdigest() == str(g["maxwithin_stdout_md5"])


def test_batched_build_writes_reference_file(tmp_path):
    g = np.load(os.path.join(GOLD, "omni.npz"))
    M, N = int(g["M"]), int(g["N"])
    hap = O.unpack_panel(g["xbits"], M)
    xin, pb = tmp_path / "x.bytes", tmp_path / "x.pbwt"
    np.ascontiguousarray(hap.T).tofile(xin)
    subprocess.run([SELFCHECK, str(xin), str(M), str(N), str(pb)], check=True, capture_output=True)
    assert hashlib.md5(open(pb, "rb").read()).hexdigest() == str(g["pbwt_md5"])


@pytest.mark.parametrize("batch", [0, 5])
def test_lazy_cursor_adopts_a_non_identity_start_order(tmp_path, batch):
    """A reader's cursor that does not start from the identity (p->aFstart set) goes through the batched device build as
    well -- there is no hand-over to the reference's CPU function: u->a stays frozen at the start order, the shim
    recovers the parsed column through it and carries the true order between batches."""
    M, N = 37, 23
    rng = np.random.default_rng(5)
    hap = (rng.random((M, N)) < 0.4).astype(np.uint8)
    start = rng.permutation(M).astype(np.int32)
    xin, sin, pb = tmp_path / "x.bytes", tmp_path / "start.i32", tmp_path / "x.pbwt"
    np.ascontiguousarray(hap.T).tofile(xin)
    start.tofile(sin)
    env = dict(os.environ)
    if batch:
        env["PBWTGPU_READER_BATCH"] = str(batch)
    subprocess.run([SELFCHECK, str(xin), str(M), str(N), str(pb), str(sin)], check=True, capture_output=True, env=env)
    imp = O.reference() or O.port()
    yz, a_end, _ = imp.build(O.pack_panel(hap), M, N, a_start=start, with_d=False)
    want = tmp_path / "want.pbwt"
    write_pbwt(want, M, N, start, a_end, yz)
    assert open(pb, "rb").read() == open(want, "rb").read()


# ---- differential runs against the unmodified reference CLI (oracle/_ref/pbwt_ref), where it is present ----

REF_CLI = os.path.join(ROOT, "oracle", "_ref", "pbwt_ref")
needs_ref_cli = pytest.mark.skipif(not os.path.exists(REF_CLI), reason="reference CLI (oracle/_ref/pbwt_ref) not present")


def write_vcfq(path, hap):
    with open(path, "w") as fh:
        for k in range(hap.shape[1]):
            fh.write("20\t%d\tA\tC\t" % (1000 + k) + "\t".join(str(int(v)) for v in hap[:, k]) + "\n")


def both(tmp_path, args, env=None):
    """Run the same command line through the reference and through the shim build; returns (stdout, log) of each with
    the timing lines (utils.c timeUpdate) and the shim's own notices removed."""
    res = []
    for tag, exe in (("ref", REF_CLI), ("shim", CLI)):
        lg = tmp_path / (tag + ".log")
        a = [str(x).replace("@", tag) for x in args]
        out = subprocess.run([exe, "-log", str(lg), *a], check=True, capture_output=True, env=env).stdout
        log = [ln for ln in open(lg).read().splitlines() if "\tsystem\t" not in ln and not ln.startswith("pbwt_gpu:")]
        res.append((out, log))
    return res


@needs_ref_cli
@pytest.mark.parametrize("opts", [["-maxWithin"], ["-longWithin", "3"], ["-longWithin", "11"], ["-stats", "-maxWithin"],
                                  ["-stats", "-longWithin", "4"], ["-check", "-maxWithin"], ["-check", "-longWithin", "5"]])
@pytest.mark.parametrize("M,N,dens,seed", [(40, 30, 0.3, 5), (9, 64, 0.5, 6), (130, 25, 0.05, 7)])
def test_match_commands_equal_reference(tmp_path, opts, M, N, dens, seed):
    rng = np.random.default_rng(seed)
    hap = (rng.random((M, N)) < dens).astype(np.uint8)
    hap[: M // 4] = hap[0]
    vq = tmp_path / "x.vcfq"
    write_vcfq(vq, hap)
    pre = [o for o in opts if o in ("-stats", "-check")]
    cmd = [o for o in opts if o not in ("-stats", "-check")]
    (ro, rl), (so, sl) = both(tmp_path, [*pre, "-readVcfq", vq, *cmd], env={**os.environ, "PBWTGPU_READER_BATCH": "7"})
    assert ro == so
    assert rl == sl


@needs_ref_cli
@pytest.mark.parametrize("M,Q,N,seed", [(30, 5, 40, 1), (8, 12, 25, 2)])
def test_matchdynamic_equals_reference(tmp_path, M, Q, N, seed):
    rng = np.random.default_rng(seed)
    hp = (rng.random((M, N)) < 0.2).astype(np.uint8)
    hq = (rng.random((Q, N)) < 0.4).astype(np.uint8)
    hq[0] = hp[1]
    write_vcfq(tmp_path / "p.vcfq", hp)
    write_vcfq(tmp_path / "q.vcfq", hq)
    for tag, exe in (("ref", REF_CLI), ("shim", CLI)):  # the two files each build writes must be identical too
        subprocess.run([exe, "-readVcfq", str(tmp_path / "p.vcfq"), "-write", str(tmp_path / f"p.{tag}.pbwt")], check=True, capture_output=True)
        subprocess.run([exe, "-readVcfq", str(tmp_path / "q.vcfq"), "-write", str(tmp_path / f"q.{tag}.pbwt")], check=True, capture_output=True)
    assert open(tmp_path / "p.ref.pbwt", "rb").read() == open(tmp_path / "p.shim.pbwt", "rb").read()
    assert open(tmp_path / "q.ref.pbwt", "rb").read() == open(tmp_path / "q.shim.pbwt", "rb").read()
    (ro, rl), (so, sl) = both(tmp_path, ["-read", tmp_path / "p.ref.pbwt", "-matchDynamic", tmp_path / "q.ref.pbwt"])
    assert ro == so and rl == sl


@needs_ref_cli
def test_reverse_and_macs_reader_equal_reference(tmp_path):
    rng = np.random.default_rng(9)
    M, N = 24, 50
    hap = (rng.random((M, N)) < 0.35).astype(np.uint8)
    macs = tmp_path / "x.macs"  # pbwtReadMacs, pbwtIO.c:460-492
    with open(macs, "w") as fh:
        fh.write(f"COMMAND:\tmacs {M} 100000\nSEED:\t1\n")
        for k in range(N):
            fh.write(f"SITE:\t{k}\t{(k + 1) / (N + 2):.6f}\t0.1\t" + "".join(str(int(v)) for v in hap[:, k]) + "\n")
        fh.write(f"TOTAL_SAMPLES:\t{M}\nTOTAL_SITES:\t{N}\nBEGIN_SELECTED_SITES\n" + "\t".join(str(k) for k in range(N)) + "\nEND_SELECTED_SITES\n")
    for batch in ("4", "1000"):
        (ro, rl), (so, sl) = both(tmp_path, ["-readMacs", macs, "-write", tmp_path / "@.pbwt", "-buildReverse", "-writeReverse", tmp_path / "@.rev", "-maxWithin"],
                                  env={**os.environ, "PBWTGPU_READER_BATCH": batch})
        assert ro == so and rl == sl
        assert open(tmp_path / "ref.pbwt", "rb").read() == open(tmp_path / "shim.pbwt", "rb").read()
        assert open(tmp_path / "ref.rev", "rb").read() == open(tmp_path / "shim.rev", "rb").read()


def founder_haps(rng, M, N, founders=6, flip=0.01):
    F = (rng.random((founders, N)) < 0.4).astype(np.uint8)
    hap = np.empty((M, N), np.uint8)
    for h in range(M):
        src = rng.integers(founders)
        for k in range(N):
            if rng.random() < 0.05:
                src = rng.integers(founders)
            hap[h, k] = F[src, k]
    return hap ^ (rng.random((M, N)) < flip).astype(np.uint8)


@needs_ref_cli
@pytest.mark.parametrize("seed,step", [(3, 3), (4, 1), (5, 7)])
def test_reference_impute_consumer_equals_reference(tmp_path, seed, step):
    """-referenceImpute (pbwtImpute.c:1106-1290): the reference's imputation drives matchMaximalWithin / matchSequencesSweep
    through its own reportMatch callback (:1142-1146, appending per-query arrays in callback order) and pbwtBuildReverse
    (:1283).  With the shim's versions underneath, the imputed .pbwt, sites, haplotypes and log must be the reference's."""
    rng = np.random.default_rng(seed)
    M, N = 60, 80
    hap = founder_haps(rng, M, N)
    write_vcfq(tmp_path / "ref.vcfq", hap[:40])
    with open(tmp_path / "q.vcfq", "w") as fh:  # the query panel is typed at every step-th site only
        for k in range(0, N, step):
            fh.write("20\t%d\tA\tC\t" % (1000 + k) + "\t".join(str(int(v)) for v in hap[40:, k]) + "\n")
    outs = {}
    for tag, exe in (("ref", REF_CLI), ("shim", CLI)):
        d = tmp_path / tag
        d.mkdir()
        subprocess.run([exe, "-readVcfq", str(tmp_path / "ref.vcfq"), "-writeAll", "panel"], check=True, capture_output=True, cwd=d)
        out = subprocess.run([exe, "-log", "log", "-readVcfq", str(tmp_path / "q.vcfq"), "-referenceImpute", "panel", "-write", "imp.pbwt",
                              "-writeSites", "imp.sites", "-haps", "imp.haps"], check=True, capture_output=True, cwd=d).stdout
        log = [ln for ln in open(d / "log").read().splitlines() if "\tsystem\t" not in ln and not ln.startswith("pbwt_gpu:")]
        outs[tag] = (out, log, *(open(d / f, "rb").read() for f in ("panel.pbwt", "imp.pbwt", "imp.sites", "imp.haps")))
    assert outs["ref"] == outs["shim"]
    assert any(ln.startswith("impute against reference") for ln in outs["ref"][1]) and len(outs["ref"][3]) > 0


@needs_ref_cli
def test_paint_consumer_equals_reference(tmp_path):
    """-paint (pbwtPaint.c:66-137): chromosome painting accumulates the records of matchMaximalWithin (:87) per
    haplotype pair; its four output matrices must be the reference's."""
    rng = np.random.default_rng(8)
    hap = founder_haps(rng, 40, 90)
    write_vcfq(tmp_path / "x.vcfq", hap)
    files = ["pt.chunkcounts.out", "pt.chunklengths.out", "pt.regionchunkcounts.out", "pt.regionsquaredchunkcounts.out"]
    outs = {}
    for tag, exe in (("ref", REF_CLI), ("shim", CLI)):
        d = tmp_path / tag
        d.mkdir()
        subprocess.run([exe, "-readVcfq", str(tmp_path / "x.vcfq"), "-paint", "pt", "10", "2"], check=True, capture_output=True, cwd=d)
        outs[tag] = [open(d / f, "rb").read() for f in files]
    assert outs["ref"] == outs["shim"] and all(len(b) > 0 for b in outs["ref"])


@needs_ref_cli
@pytest.mark.parametrize("batch", ["4", "25", "1000"])
def test_checkpoint_hook_equals_reference(tmp_path, batch):
    """-checkpoint n (pbwtCheckPoint, pbwtIO.c:158-168, called from the readers every n sites): the hook copies the
    cursor's order into p->aFend and writes the panel so far; through the shim's lazy cursor that forces a flush of the
    pending batch in the middle of reading.  Checkpoint files and the final file must be the reference's."""
    rng = np.random.default_rng(12)
    hap = founder_haps(rng, 30, 64)
    write_vcfq(tmp_path / "x.vcfq", hap)
    files = ["check_A.pbwt", "check_A.sites", "check_B.pbwt", "check_B.sites", "out.pbwt"]
    outs = {}
    for tag, exe in (("ref", REF_CLI), ("shim", CLI)):
        d = tmp_path / tag
        d.mkdir()
        subprocess.run([exe, "-checkpoint", "25", "-readVcfq", str(tmp_path / "x.vcfq"), "-write", "out.pbwt"], check=True,
                       capture_output=True, cwd=d, env={**os.environ, "PBWTGPU_READER_BATCH": batch})
        outs[tag] = [open(d / f, "rb").read() for f in files]
    assert outs["ref"] == outs["shim"]


@needs_ref_cli
@pytest.mark.parametrize("M,Q,N,chunk,seed", [(30, 13, 40, 4, 1), (8, 12, 25, 5, 2), (50, 33, 60, 8, 3), (20, 9, 30, 1, 4), (40, 17, 35, 16, 5)])
def test_matchdynamic_more_queries_than_one_device_call(tmp_path, M, Q, N, chunk, seed):
    """More queries than one pbwtgpu_match_sweep call takes (8,192 on the device; here the mock refuses more than
    `chunk`): the shim cuts the queries into chunks, sweeps each against the panel and merges the record streams site by
    site in the order of the whole query PBWT.  stdout (MATCH lines in callback order) and the log ("no match" lines,
    summary) must be the reference's."""
    rng = np.random.default_rng(seed)
    hap = founder_haps(rng, M + Q, N, founders=4, flip=0.03)
    hp, hq = hap[:M].copy(), hap[M:].copy()
    hp[:, 5] = 0  # a monomorphic panel site with some queries carrying the other allele: "no match" events
    hq[::2, 5] = 1
    hq[1] = hq[0]  # tied queries
    write_vcfq(tmp_path / "p.vcfq", hp)
    write_vcfq(tmp_path / "q.vcfq", hq)
    subprocess.run([REF_CLI, "-readVcfq", str(tmp_path / "p.vcfq"), "-write", str(tmp_path / "p.pbwt")], check=True, capture_output=True)
    subprocess.run([REF_CLI, "-readVcfq", str(tmp_path / "q.vcfq"), "-write", str(tmp_path / "q.pbwt")], check=True, capture_output=True)
    env = {**os.environ, "PBWTGPU_QUERY_CHUNK": str(chunk), "PBWTGPU_MOCK_QMAX": str(chunk)}
    (ro, rl), (so, sl) = both(tmp_path, ["-read", tmp_path / "p.pbwt", "-matchDynamic", tmp_path / "q.pbwt"], env=env)
    assert ro == so
    assert rl == sl
    assert any(ln.startswith("no match") for ln in rl) and len(ro) > 0


@needs_ref_cli
def test_reference_impute_with_chunked_queries(tmp_path):
    """-referenceImpute with more target haplotypes than one device call takes: the imputation's per-query arrays are
    filled in callback order (pbwtImpute.c:1149-1165), so the merged order must be exact."""
    rng = np.random.default_rng(21)
    hap = founder_haps(rng, 70, 80)
    write_vcfq(tmp_path / "ref.vcfq", hap[:40])
    with open(tmp_path / "q.vcfq", "w") as fh:
        for k in range(0, 80, 3):
            fh.write("20\t%d\tA\tC\t" % (1000 + k) + "\t".join(str(int(v)) for v in hap[40:, k]) + "\n")
    env = {**os.environ, "PBWTGPU_QUERY_CHUNK": "7", "PBWTGPU_MOCK_QMAX": "7"}
    outs = {}
    for tag, exe in (("ref", REF_CLI), ("shim", CLI)):
        d = tmp_path / tag
        d.mkdir()
        subprocess.run([exe, "-readVcfq", str(tmp_path / "ref.vcfq"), "-writeAll", "panel"], check=True, capture_output=True, cwd=d)
        subprocess.run([exe, "-log", "log", "-readVcfq", str(tmp_path / "q.vcfq"), "-referenceImpute", "panel", "-write", "imp.pbwt", "-haps", "imp.haps"],
                       check=True, capture_output=True, cwd=d, env=env)
        log = [ln for ln in open(d / "log").read().splitlines() if "\tsystem\t" not in ln and not ln.startswith("pbwt_gpu:")]
        outs[tag] = (log, open(d / "imp.pbwt", "rb").read(), open(d / "imp.haps", "rb").read())
    assert outs["ref"] == outs["shim"]


@needs_ref_cli
def test_check_option_verifies_device_records(tmp_path):
    """-check: the matches still come from the device library and every printed one is verified against the decoded
    haplotypes (the reference's checkMatchMaximal, pbwtMatch.c:33-58); -buildReverse also dumps rev.haps (pbwtCore.c:181-187).
    Same stdout / log / files as the reference, and a corrupted device record must be fatal."""
    rng = np.random.default_rng(31)
    hap = founder_haps(rng, 36, 50, founders=4, flip=0.03)
    write_vcfq(tmp_path / "p.vcfq", hap[:28])
    write_vcfq(tmp_path / "q.vcfq", hap[28:])
    subprocess.run([REF_CLI, "-readVcfq", str(tmp_path / "q.vcfq"), "-write", str(tmp_path / "q.pbwt")], check=True, capture_output=True)
    outs = {}
    for tag, exe in (("ref", REF_CLI), ("shim", CLI)):
        d = tmp_path / tag
        d.mkdir()
        out = subprocess.run([exe, "-log", "log", "-check", "-readVcfq", str(tmp_path / "p.vcfq"), "-maxWithin", "-longWithin", "6",
                              "-matchDynamic", str(tmp_path / "q.pbwt"), "-buildReverse"], check=True, capture_output=True, cwd=d).stdout
        log = [ln for ln in open(d / "log").read().splitlines() if "\tsystem\t" not in ln]
        outs[tag] = (out, log, open(d / "rev.haps", "rb").read())
    assert outs["ref"] == outs["shim"] and len(outs["ref"][2]) > 0
    bad = subprocess.run([CLI, "-check", "-readVcfq", str(tmp_path / "p.vcfq"), "-maxWithin"], capture_output=True,
                         env={**os.environ, "PBWTGPU_MOCK_CORRUPT": "1"})
    assert bad.returncode != 0 and b"FATAL ERROR: match not" in bad.stderr
    good = subprocess.run([CLI, "-readVcfq", str(tmp_path / "p.vcfq"), "-maxWithin"], capture_output=True,
                          env={**os.environ, "PBWTGPU_MOCK_CORRUPT": "1"})
    assert good.returncode == 0  # without -check nobody looks


@needs_ref_cli
@pytest.mark.parametrize("batch", ["5", "1000"])
def test_other_readers_equal_reference(tmp_path, batch):
    """The remaining readers of pbwtIO.c that step a cursor per parsed site -- -readHap / -readGen (pbwtReadLineFile,
    :558-586), -readHapLegend (pbwtReadLineTwoFile, :590-623) and -readPhase (:777-826) -- through the shim's lazy cursor."""
    rng = np.random.default_rng(17)
    M, N = 20, 33
    hap = (rng.random((M, N)) < 0.35).astype(np.uint8)
    with open(tmp_path / "x.hap", "w") as fh:
        for k in range(N):
            fh.write(f"snp{k} rs{k} {1000 + k} A C " + " ".join(str(int(v)) for v in hap[:, k]) + "\n")
    with open(tmp_path / "x.gen", "w") as fh:
        for k in range(N):
            tr = [["1 0 0", "0 1 0", "0 0 1"][int(hap[2 * i, k]) + int(hap[2 * i + 1, k])] for i in range(M // 2)]
            fh.write(f"snp{k} rs{k} {1000 + k} A C " + " ".join(tr) + "\n")
    with open(tmp_path / "x.hl.hap", "w") as fh, open(tmp_path / "x.hl.legend", "w") as lh:
        for k in range(N):
            fh.write(" ".join(str(int(v)) for v in hap[:, k]) + "\n")
            lh.write(f"rs{k} {1000 + k} A C\n")
    with open(tmp_path / "x.phase", "w") as fh:
        fh.write(f"{M}\n{N}\nP " + " ".join(str(1000 + k) for k in range(N)) + "\n")
        for h in range(M):
            fh.write("".join(str(int(v)) for v in hap[h]) + "\n")
    env = {**os.environ, "PBWTGPU_READER_BATCH": batch}
    for reader in (["-readHap", tmp_path / "x.hap", "20"], ["-readGen", tmp_path / "x.gen", "20"],
                   ["-readHapLegend", tmp_path / "x.hl.hap", tmp_path / "x.hl.legend", "20"], ["-readPhase", tmp_path / "x.phase"]):
        (ro, rl), (so, sl) = both(tmp_path, [*reader, "-write", tmp_path / "@.pbwt", "-maxWithin"], env=env)
        assert ro == so and rl == sl and len(ro) > 0, reader[0]
        assert open(tmp_path / "ref.pbwt", "rb").read() == open(tmp_path / "shim.pbwt", "rb").read(), reader[0]


@needs_ref_cli
@pytest.mark.parametrize("batch", ["6", "1000"])
def test_subsample_rebuild_equals_reference(tmp_path, batch):
    """-subsample start n (pbwtSubSample, pbwtSample.c:57-93) rebuilds the panel of the selected haplotypes with the same
    per-site gather + pbwtCursorWriteForwards step as the readers; pbwtSample.c is compiled with the same two renames,
    so the rebuild is batched onto the device too.  File and matches of the sub-panel must be the reference's."""
    rng = np.random.default_rng(41)
    hap = founder_haps(rng, 40, 70)
    write_vcfq(tmp_path / "x.vcfq", hap)
    env = {**os.environ, "PBWTGPU_READER_BATCH": batch}
    (ro, rl), (so, sl) = both(tmp_path, ["-readVcfq", tmp_path / "x.vcfq", "-subsample", "5", "20", "-write", tmp_path / "@.pbwt", "-maxWithin",
                                         "-subsample", "2", "7", "-write", tmp_path / "@.2.pbwt"], env=env)
    assert ro == so and rl == sl and len(ro) > 0
    for f in ("pbwt", "2.pbwt"):
        assert open(tmp_path / f"ref.{f}", "rb").read() == open(tmp_path / f"shim.{f}", "rb").read()


REF_TEST = "/root/reference/test"


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_TEST, "merge.1.tab")), reason="the reference's test fixtures are not present")
def test_reference_test_pl_cases(tmp_path):
    """The cases of the reference's own test/test.pl that need no htslib, replayed through the drop-in build:
    test_pbwt (merge.1, merge.2: -readVcfq -> -write / -writeSites, then -read -> -haps), test_merge and
    test_merge_sites (-merge of the two files), against the reference's expected outputs test/merge.*.out / .sites."""
    def sh(cmd):
        return subprocess.run(["/bin/bash", "-o", "pipefail", "-c", cmd], check=True, capture_output=True, env={**os.environ, "PBWTGPU_READER_BATCH": "3"})
    for name in ("merge.1", "merge.2"):
        sh(f"{CLI} -readVcfq {REF_TEST}/{name}.tab -write {tmp_path}/{name}.pbwt -writeSites {tmp_path}/{name}.sites")
        sh(f"{CLI} -read {tmp_path}/{name}.pbwt -haps {tmp_path}/{name}.out")
        assert open(tmp_path / f"{name}.out").read() == open(f"{REF_TEST}/{name}.out").read()
    files = f"{tmp_path}/merge.1.pbwt {tmp_path}/merge.2.pbwt"
    sh(f"{CLI} -merge {files} -write - | {CLI} -read - -haps {tmp_path}/merge.12.out")
    assert open(tmp_path / "merge.12.out").read() == open(f"{REF_TEST}/merge.12.out").read()
    sh(f"{CLI} -merge {files} -write - -writeSites {tmp_path}/merge.12.sites > /dev/null")
    assert open(tmp_path / "merge.12.sites").read() == open(f"{REF_TEST}/merge.12.sites").read()


# ---- -readVcfGT without htslib (pbwt_b200/host/vcf_text.c) ----

def write_vcf(path, hap, chrom="20", pos0=1000, phased=True):
    """Diploid text VCF of a [M, N] haplotype matrix (sample i = haplotypes 2i, 2i+1)."""
    M, N = hap.shape
    sep = "|" if phased else "/"
    with open(path, "w") as fh:
        fh.write("##fileformat=VCFv4.2\n##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n")
        fh.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(f"S{i}" for i in range(M // 2)) + "\n")
        for k in range(N):
            gts = "\t".join(f"{int(hap[2 * i, k])}{sep}{int(hap[2 * i + 1, k])}" for i in range(M // 2))
            fh.write(f"{chrom}\t{pos0 + k}\t.\tA\tC\t.\tPASS\t.\tGT\t{gts}\n")


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_TEST, "OMNI.vcf")), reason="the reference's test fixtures are not present")
def test_baseline_config1_command_line(tmp_path):
    """BASELINE.json configs[0], verbatim: `pbwt -readVcfGT test/OMNI.vcf -maxWithin` (plus -write).  The reference needs
    htslib for it; the drop-in's text reader (vcf_text.c, following pbwtHtslib.c:52-172) feeds the batched device build.
    Known answers of SURVEY section 8c: .pbwt md5 d2753c04..., -maxWithin stdout 21,883 lines md5 b1412a0e...."""
    pb, lg = tmp_path / "omni.pbwt", tmp_path / "log"
    out = run("-log", lg, "-readVcfGT", os.path.join(REF_TEST, "OMNI.vcf"), "-write", pb, "-maxWithin",
              env={**os.environ, "PBWTGPU_READER_BATCH": "7"})
    assert hashlib.md5(open(pb, "rb").read()).hexdigest() == "d2753c044392ec4f474c7061cff773e9"
    assert len(out.splitlines()) == 21883 and hashlib.md5(out).hexdigest() == "b1412a0e3057b3e92ade2cee7c4a4824"
    assert "with 1376 sample names and 38 sites on chromosome 20: M, N are 2752, 38" in open(lg).read()


@needs_ref_cli
@pytest.mark.parametrize("phased", [True, False])
def test_read_vcf_gt_equals_vcfq_route(tmp_path, phased):
    """Biallelic, no missing: -readVcfGT must give the file the reference builds from the same genotypes via -readVcfq."""
    rng = np.random.default_rng(51)
    hap = founder_haps(rng, 30, 45)
    write_vcf(tmp_path / "x.vcf", hap, phased=phased)
    write_vcfq(tmp_path / "x.vcfq", hap)
    subprocess.run([REF_CLI, "-readVcfq", str(tmp_path / "x.vcfq"), "-write", str(tmp_path / "ref.pbwt")], check=True, capture_output=True)
    run("-readVcfGT", tmp_path / "x.vcf", "-write", tmp_path / "shim.pbwt", "-writeSites", tmp_path / "shim.sites", env={**os.environ, "PBWTGPU_READER_BATCH": "8"})
    assert open(tmp_path / "ref.pbwt", "rb").read() == open(tmp_path / "shim.pbwt", "rb").read()
    assert open(tmp_path / "shim.sites").read().splitlines()[0] == "20\t1000\tA\tC"


def test_read_vcf_gt_multiallelic_missing_haploid(tmp_path):
    """The record handling of pbwtHtslib.c:94-168: multi-allelic records become one site per ALT allele (y = allele == i),
    missing alleles read as 0 and are counted, all-haploid records as homozygous diploid, ALT '.' stays a site, records
    without GT are skipped, reading stops at the second chromosome, REF/ALT are upper-cased."""
    vcf = tmp_path / "m.vcf"
    with open(vcf, "w") as fh:
        fh.write("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tA\tB\tC\n")
        fh.write("X\t10\t.\tt\t.\t.\t.\t.\tGT\t.\t.\t.\n")                          # haploid, all missing, no ALT
        fh.write("X\t20\t.\tT\tc\t9\tPASS\tDP=3\tGT:DP\t0/1:5\t1/1:7\t1|0:2\n")
        fh.write("X\t30\t.\tT\tTTTA,TA\t.\t.\t.\tDP:GT\t3:0/2\t4:0/1\t5:1/2\n")       # two sites
        fh.write("X\t40\t.\tA\tAC\t.\t.\t.\tGT\t0|1\t0|1\t./.\n")                     # missing diploid
        fh.write("X\t50\t.\tG\tC\t.\t.\t.\tDP\t1\t2\t3\n")                            # no GT: skipped
        fh.write("X\t60\t.\tG\tC\t.\t.\t.\tGT\t0\t1\t1\n")                            # haploid line
        fh.write("Y\t70\t.\tG\tC\t.\t.\t.\tGT\t0|1\t0|1\t0|1\n")                      # another chromosome: stop
    lg = tmp_path / "log"
    run("-log", lg, "-readVcfGT", vcf, "-haps", tmp_path / "h.out", "-writeSites", tmp_path / "s.out")
    assert open(tmp_path / "h.out").read().split() == ["000000", "011110", "000110", "010001", "010100", "001111"]
    assert open(tmp_path / "s.out").read().splitlines() == ["X\t10\tT\t.", "X\t20\tT\tC", "X\t30\tT\tTTTA", "X\t30\tT\tTA", "X\t40\tA\tAC", "X\t60\tG\tC"]
    log = open(lg).read()
    assert "with 3 sample names and 6 sites on chromosome X: M, N are 6, 6" in log and "8 missing values at 2 sites" in log


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_TEST, "refImpute.out.vcf")), reason="the reference's test fixtures are not present")
def test_reference_test_pl_reference_impute(tmp_path):
    """test_pbwt_reference_impute of the reference's test/test.pl: -readVcfGT OMNI.vcf -writeAll, then
    -readVcfGT refImpute.in.vcf -referenceImpute OMNI.  The reference writes a VCF (htslib); here the imputed haplotypes
    and sites are written with -haps / -writeSites and compared with the GT columns and the fixed fields of the expected
    test/refImpute.out.vcf."""
    subprocess.run([CLI, "-readVcfGT", os.path.join(REF_TEST, "OMNI.vcf"), "-writeAll", "OMNI"], check=True, capture_output=True, cwd=tmp_path)
    subprocess.run([CLI, "-readVcfGT", os.path.join(REF_TEST, "refImpute.in.vcf"), "-referenceImpute", "OMNI", "-haps", "out.haps",
                    "-writeSites", "out.sites"], check=True, capture_output=True, cwd=tmp_path)
    want_haps, want_sites = [], []
    for ln in open(os.path.join(REF_TEST, "refImpute.out.vcf")):
        if not ln.startswith("#"):
            f = ln.rstrip("\n").split("\t")
            want_sites.append("\t".join((f[0], f[1], f[3], f[4])))
            want_haps.append("".join(g.replace("|", "").replace("/", "") for g in f[9:]))
    assert open(tmp_path / "out.haps").read().split() == want_haps and len(want_haps) == 2964
    assert open(tmp_path / "out.sites").read().splitlines() == want_sites


@needs_ref_cli
def test_read_vcf_gt_compressed_and_stdin(tmp_path):
    """The same VCF plain, gzip-compressed, bgzip-like (several concatenated gzip members, a line split across two of
    them) and on stdin must all give the reference's file; BCF is refused with a message."""
    import gzip
    rng = np.random.default_rng(52)
    hap = founder_haps(rng, 24, 40)
    write_vcf(tmp_path / "x.vcf", hap)
    write_vcfq(tmp_path / "x.vcfq", hap)
    subprocess.run([REF_CLI, "-readVcfq", str(tmp_path / "x.vcfq"), "-write", str(tmp_path / "ref.pbwt")], check=True, capture_output=True)
    want = open(tmp_path / "ref.pbwt", "rb").read()
    text = open(tmp_path / "x.vcf", "rb").read()
    open(tmp_path / "x.vcf.gz", "wb").write(gzip.compress(text))
    cut = [0, len(text) // 3 + 7, 2 * len(text) // 3 + 5, len(text)]
    open(tmp_path / "x.bgz.vcf.gz", "wb").write(b"".join(gzip.compress(text[a:b]) for a, b in zip(cut, cut[1:])))
    for name in ("x.vcf", "x.vcf.gz", "x.bgz.vcf.gz"):
        run("-readVcfGT", tmp_path / name, "-write", tmp_path / "out.pbwt", env={**os.environ, "PBWTGPU_READER_BATCH": "9"})
        assert open(tmp_path / "out.pbwt", "rb").read() == want, name
    subprocess.run([CLI, "-readVcfGT", "-", "-write", str(tmp_path / "stdin.pbwt")], input=text, check=True, capture_output=True)
    assert open(tmp_path / "stdin.pbwt", "rb").read() == want
    open(tmp_path / "x.bcf", "wb").write(gzip.compress(b"BCF\x02\x02" + b"\0" * 64))
    bad = subprocess.run([CLI, "-readVcfGT", str(tmp_path / "x.bcf")], capture_output=True)
    assert bad.returncode != 0 and b"BCF needs htslib" in bad.stderr
