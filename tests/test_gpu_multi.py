"""More than one GPU (skipped when fewer than two are visible): the NCCL site-sharded build + maxWithin must equal the
single-GPU build of the same panel -- which the parity suite pins to the reference -- pack3 bytes, aFend, d and every
record in callback order.  One process per GPU under torchrun, rendezvous on 127.0.0.1 (scripts/multi_gpu_check.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _world_sizes():
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    return [w for w in (2, 3, 4, 8) if w <= n]


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_site_sharded_build_over_nccl_equals_one_gpu(world):
    if world not in _world_sizes():
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29530 + world), os.path.join(ROOT, "scripts", "multi_gpu_check.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("MULTI_GPU_CHECK")]
    assert out.returncode == 0 and lines, out.stderr[-2000:]
    assert lines[-1].split()[1] == "OK" and f"world {world}" in lines[-1], lines[-1]
