"""GPU test of the drop-in CLI: the reference's own pbwtMain.c / pbwtIO.c linked with the host shim and
libpbwtgpu.so (pbwt_b200/host/_build/pbwt_gpu, built where the reference is present and shipped prebuilt).
stdout of -maxWithin / -matchDynamic and the bytes of -write must equal the reference's (golden md5s)."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
CLI = os.path.join(ROOT, "pbwt_b200", "host", "_build", "pbwt_gpu")
SELFTEST = os.path.join(ROOT, "pbwt_b200", "host", "_build", "shim_selftest")

needs_cli = pytest.mark.skipif(not os.path.exists(CLI), reason="pbwt_gpu not built (needs the reference sources at build time)")


def write_pbwt(path, M, N, a_start, a_end, yz):
    with open(path, "wb") as fh:  # PBW3 layout, pbwtIO.c:33-57
        fh.write(b"PBW3" + np.array([M, N], "<i4").tobytes() + np.asarray(a_start, "<i4").tobytes() +
                 np.asarray(a_end, "<i4").tobytes() + np.array([len(yz)], "<i8").tobytes() + b"    " +
                 np.asarray(yz, np.uint8).tobytes())


@needs_cli
@pytest.mark.parametrize("name", ["merge1", "omni", "mosaic300x400"])
def test_cli_maxwithin_stdout(tmp_path, name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    M, N = int(g["M"]), int(g["N"])
    pb = tmp_path / "x.pbwt"
    write_pbwt(pb, M, N, np.arange(M), g["a_end"], g["yz"])
    assert hashlib.md5(open(pb, "rb").read()).hexdigest() == str(g["pbwt_md5"])
    out = subprocess.run([CLI, "-read", str(pb), "-maxWithin"], check=True, capture_output=True).stdout
    assert len(out.splitlines()) == int(g["maxwithin_stdout_lines"])
    assert hashlib.md5(out).hexdigest() == str(g["maxwithin_stdout_md5"])


@needs_cli
@pytest.mark.parametrize("name,L", [("merge1", 2), ("omni", 20), ("mosaic300x400", 40)])
def test_cli_longwithin_stdout(tmp_path, name, L):
    """pbwt_gpu -read X.pbwt -longWithin L prints the reference's MATCH lines (golden md5 of its stdout)."""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    lw = np.load(os.path.join(GOLD, "long_within.npz"))
    M, N = int(g["M"]), int(g["N"])
    pb = tmp_path / "x.pbwt"
    write_pbwt(pb, M, N, np.arange(M), g["a_end"], g["yz"])
    out = subprocess.run([CLI, "-read", str(pb), "-longWithin", str(L)], check=True, capture_output=True).stdout
    assert len(out.splitlines()) == len(lw[f"{name}_L{L}"])
    assert hashlib.md5(out).hexdigest() == str(lw[f"{name}_L{L}_md5"])


@needs_cli
@pytest.mark.parametrize("name", ["merge1", "omni", "mosaic300x400"])
def test_cli_build_reverse_file(tmp_path, name):
    """pbwt_gpu -read X.pbwt -buildReverse -writeReverse R.pbwt writes the reference's reverse file byte for byte."""
    g, r = np.load(os.path.join(GOLD, name + ".npz")), np.load(os.path.join(GOLD, "reverse.npz"))
    M, N = int(g["M"]), int(g["N"])
    pb, rv = tmp_path / "x.pbwt", tmp_path / "r.pbwt"
    write_pbwt(pb, M, N, np.arange(M), g["a_end"], g["yz"])
    subprocess.run([CLI, "-read", str(pb), "-buildReverse", "-writeReverse", str(rv)], check=True, capture_output=True)
    assert hashlib.md5(open(rv, "rb").read()).hexdigest() == str(r[name + "_md5"])


@needs_cli
@pytest.mark.parametrize("name,md5,log", [
    ("merge1", "8d1029f275f8c490f3c50923db651419", ["Average 2.9 matches per sample", "Average length 4.9"]),
    ("omni", "3efa4cec8b933c7a984c8b96cae7e035", ["Average 4.0 matches per sample", "Average length 20.7"]),
    ("mosaic300x400", "a6e0ff75052b0c53c1e193a479eb6965", ["Average 12.4 matches per sample", "Average length 69.8"])])
def test_cli_stats_maxwithin(tmp_path, name, md5, log):
    """pbwt_gpu -stats -read X.pbwt -maxWithin: the length histogram and the two summary lines of the reference
    (stdout md5s and log lines recorded from the reference CLI, `pbwt -log L -stats -read X.pbwt -maxWithin`)."""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    M, N = int(g["M"]), int(g["N"])
    pb, lg = tmp_path / "x.pbwt", tmp_path / "log"
    write_pbwt(pb, M, N, np.arange(M), g["a_end"], g["yz"])
    out = subprocess.run([CLI, "-log", str(lg), "-stats", "-read", str(pb), "-maxWithin"], check=True, capture_output=True).stdout
    assert hashlib.md5(out).hexdigest() == md5
    assert [ln for ln in open(lg).read().splitlines() if ln.startswith("Average")] == log


@needs_cli
def test_cli_matchdynamic_stdout_and_log(tmp_path):
    g = np.load(os.path.join(GOLD, "query_sweep.npz"))
    M, Q, N = int(g["M"]), int(g["Q"]), int(g["N"])
    pp, qp, lg = tmp_path / "p.pbwt", tmp_path / "q.pbwt", tmp_path / "log"
    write_pbwt(pp, M, N, np.arange(M), g["pa_end"], g["pyz"])
    write_pbwt(qp, Q, N, np.arange(Q), g["qa_end"], g["qyz"])
    out = subprocess.run([CLI, "-log", str(lg), "-read", str(pp), "-matchDynamic", str(qp)], check=True, capture_output=True).stdout
    assert hashlib.md5(out).hexdigest() == str(g["matchdynamic_stdout_md5"])
    log = [ln for ln in open(lg).read().splitlines() if ln.startswith("no match") or ln.startswith("Average number")]
    assert "\n".join(log) == str(g["matchdynamic_log"])


@needs_cli
@pytest.mark.parametrize("name,batch", [("merge1", 3), ("omni", 7), ("omni", 0), ("mosaic300x400", 128)])
def test_cli_reader_batched_ingest(tmp_path, name, batch):
    """pbwt_gpu -readVcfq x -write y: the reference's own reader and writer, with the reader's per-site cursor
    step batched onto the GPU by the shim (several batches: the order is carried between them): the .pbwt
    file must be the reference's, byte for byte."""
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
    subprocess.run([CLI, "-readVcfq", str(vq), "-write", str(pb)], check=True, capture_output=True, env=env)
    assert hashlib.md5(open(pb, "rb").read()).hexdigest() == str(g["pbwt_md5"])
    # and straight into a match command, without a file in between
    out = subprocess.run([CLI, "-readVcfq", str(vq), "-maxWithin"], check=True, capture_output=True, env=env).stdout
    assert hashlib.md5(out).hexdigest() == str(g["maxwithin_stdout_md5"])


@needs_cli
def test_shim_batched_build_writes_reference_file(tmp_path):
    """pbwtGpuBuildFromSites + the reference's own pbwtWrite must give the reference's .pbwt bytes."""
    g = np.load(os.path.join(GOLD, "omni.npz"))
    M, N = int(g["M"]), int(g["N"])
    hap = O.unpack_panel(g["xbits"], M)  # [M, N]
    xin, pb = tmp_path / "x.bytes", tmp_path / "x.pbwt"
    np.ascontiguousarray(hap.T).tofile(xin)  # site-major bytes
    subprocess.run([SELFTEST, str(xin), str(M), str(N), str(pb)], check=True, capture_output=True)
    assert hashlib.md5(open(pb, "rb").read()).hexdigest() == str(g["pbwt_md5"])
