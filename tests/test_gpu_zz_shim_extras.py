"""GPU tests of the host-shim features added after the drop-in CLI tests (tests/test_gpu_cli.py), against the real device
library: query sets larger than one device call (forced here by PBWTGPU_QUERY_CHUNK), -check as a verification of the
device's records, -stats -longWithin.  Expected outputs are the reference's goldens; the same commands are checked on the
CPU against the reference CLI with a device stand-in in tests/test_host_shim_cpu.py."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

from test_gpu_cli import CLI, GOLD, needs_cli, write_pbwt

pytestmark = pytest.mark.gpu


@needs_cli
@pytest.mark.parametrize("chunk", ["3", "7"])
def test_cli_matchdynamic_chunked_queries(tmp_path, chunk):
    g = np.load(os.path.join(GOLD, "query_sweep.npz"))
    M, Q, N = int(g["M"]), int(g["Q"]), int(g["N"])
    assert Q > int(chunk)
    pp, qp, lg = tmp_path / "p.pbwt", tmp_path / "q.pbwt", tmp_path / "log"
    write_pbwt(pp, M, N, np.arange(M), g["pa_end"], g["pyz"])
    write_pbwt(qp, Q, N, np.arange(Q), g["qa_end"], g["qyz"])
    out = subprocess.run([CLI, "-log", str(lg), "-read", str(pp), "-matchDynamic", str(qp)], check=True, capture_output=True,
                         env={**os.environ, "PBWTGPU_QUERY_CHUNK": chunk}).stdout
    assert hashlib.md5(out).hexdigest() == str(g["matchdynamic_stdout_md5"])
    log = [ln for ln in open(lg).read().splitlines() if ln.startswith("no match") or ln.startswith("Average number")]
    assert "\n".join(log) == str(g["matchdynamic_log"])


@needs_cli
@pytest.mark.parametrize("name", ["merge1", "omni"])
def test_cli_check_verifies_device_records(tmp_path, name):
    """-check: every MATCH line printed from the device's records is verified against the decoded haplotypes
    (checkMatchMaximal, pbwtMatch.c:33-58); a violation would be fatal."""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    M, N = int(g["M"]), int(g["N"])
    pb = tmp_path / "x.pbwt"
    write_pbwt(pb, M, N, np.arange(M), g["a_end"], g["yz"])
    out = subprocess.run([CLI, "-check", "-read", str(pb), "-maxWithin"], check=True, capture_output=True).stdout
    assert hashlib.md5(out).hexdigest() == str(g["maxwithin_stdout_md5"])


@needs_cli
def test_cli_stats_longwithin(tmp_path):
    """-stats -longWithin L: the reference prints the MATCH lines and the summary of an empty histogram."""
    g, lw = np.load(os.path.join(GOLD, "omni.npz")), np.load(os.path.join(GOLD, "long_within.npz"))
    M, N = int(g["M"]), int(g["N"])
    pb, lg = tmp_path / "x.pbwt", tmp_path / "log"
    write_pbwt(pb, M, N, np.arange(M), g["a_end"], g["yz"])
    out = subprocess.run([CLI, "-log", str(lg), "-stats", "-read", str(pb), "-longWithin", "20"], check=True, capture_output=True).stdout
    assert hashlib.md5(out).hexdigest() == str(lw["omni_L20_md5"])
    assert "Average 0.0 matches per sample" in open(lg).read()


@needs_cli
def test_cli_read_vcf_gt_config1(tmp_path):
    """BASELINE config 1's command line, `pbwt -readVcfGT OMNI.vcf -maxWithin`, through the text VCF reader of the drop-in
    (pbwt_b200/host/vcf_text.c) and the batched device build.  The VCF is regenerated from the golden OMNI genotypes
    (the reference's fixture does not travel to the GPU box); known answers: .pbwt md5 d2753c04..., stdout md5 b1412a0e...."""
    import oracle_lib as O
    g = np.load(os.path.join(GOLD, "omni.npz"))
    M, N = int(g["M"]), int(g["N"])
    hap = O.unpack_panel(g["xbits"], M)
    vcf, pb = tmp_path / "omni.vcf", tmp_path / "omni.pbwt"
    with open(vcf, "w") as fh:
        fh.write("##fileformat=VCFv4.2\n##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n")
        fh.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(f"S{i}" for i in range(M // 2)) + "\n")
        for k in range(N):
            fh.write(f"20\t{1000 + k}\t.\tA\tC\t.\tPASS\t.\tGT\t" + "\t".join(f"{int(hap[2 * i, k])}|{int(hap[2 * i + 1, k])}" for i in range(M // 2)) + "\n")
    out = subprocess.run([CLI, "-readVcfGT", str(vcf), "-write", str(pb), "-maxWithin"], check=True, capture_output=True,
                         env={**os.environ, "PBWTGPU_READER_BATCH": "16"}).stdout
    assert hashlib.md5(open(pb, "rb").read()).hexdigest() == str(g["pbwt_md5"])
    assert hashlib.md5(out).hexdigest() == str(g["maxwithin_stdout_md5"])
