"""CPU tests: the C-ABI library is built, loads, and exports every symbol include/pbwtgpu.h declares
(no compute calls without a GPU); without a device the constructors fail loudly -- no CPU fallback."""
import ctypes
import os
import re

import pytest

import pbwt_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "pbwtgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pbwtgpu_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    assert os.path.exists(pbwt_b200.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(pbwt_b200.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/pbwtgpu.h but not exported"


def test_words_per_col_and_strerror():
    lib = pbwt_b200.load_library()
    for M in (1, 127, 128, 129, 10000, 100000):
        assert lib.pbwtgpu_words_per_col(M) == pbwt_b200.words_per_col(M) == 4 * ((M + 127) // 128)
    assert lib.pbwtgpu_strerror(0) == b"ok"
    assert b"no CPU fallback" in lib.pbwtgpu_strerror(-5)


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pbwt_b200.PbwtGpuError):
        pbwt_b200.Engine(0)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pbwt_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle_lib" not in txt and "libpbwt_oracle" not in txt and "libpbwtref" not in txt, f
                assert "mock_pbwtgpu" not in txt and "host_mock" not in txt, f  # the CPU stand-in of tests/host_mock stays in tests/


def test_drop_in_cli_fails_loudly_without_device(tmp_path):
    """The real drop-in binary (linked with libpbwtgpu.so, not the test stand-in) must die() when there is no B200."""
    import subprocess

    import numpy as np
    import torch

    cli = os.path.join(ROOT, "pbwt_b200", "host", "_build", "pbwt_gpu")
    if torch.cuda.is_available() or not os.path.exists(cli):
        pytest.skip("a GPU is present, or pbwt_gpu is not built")
    g = np.load(os.path.join(ROOT, "tests", "golden", "merge1.npz"))
    M, N = int(g["M"]), int(g["N"])
    pb = tmp_path / "x.pbwt"
    with open(pb, "wb") as fh:
        fh.write(b"PBW3" + np.array([M, N], "<i4").tobytes() + np.arange(M, dtype="<i4").tobytes() + np.asarray(g["a_end"], "<i4").tobytes() +
                 np.array([len(g["yz"])], "<i8").tobytes() + b"    " + np.asarray(g["yz"], np.uint8).tobytes())
    r = subprocess.run([cli, "-read", str(pb), "-maxWithin"], capture_output=True)
    assert r.returncode != 0 and b"no CPU fallback" in r.stderr and r.stdout == b""
