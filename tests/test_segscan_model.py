"""CPU model of the warp-wide segmented max-scan in phase A of the sweep kernel (pbwt_b200/csrc/sweep_smem.cu, the
divergence update of pbwtCursorForwardsAD, pbwtCore.c:485-508): for every lane, the maximum of d from the nearest run
head at or below the lane (lane 0 when there is none) up to the lane.  Two formulations of the per-stage guard of the
Hillis-Steele scan are compared with the definition over all head masks of interest: the shipped one (window mask per
stage) and the opt-in PBWTGPU_SCAN_BY_DISTANCE one (a single distance to the nearest head), which round 2 will measure."""
import numpy as np
import pytest


def shfl_up(v, dl):
    out = v.copy()
    out[dl:] = v[:-dl]
    return out  # lanes below dl get their own value back, as __shfl_up_sync does


def scan_window_masks(H, d):
    v = d.copy()
    lanes = np.arange(32)
    for dl in (1, 2, 4, 8, 16):
        t = shfl_up(v, dl)
        ok = np.array([l >= dl and ((H >> (l - dl + 1)) & ((1 << dl) - 1)) == 0 for l in lanes])
        v = np.where(ok, np.maximum(v, t), v)
    return v


def scan_by_distance(H, d):
    v = d.copy()
    lanes = np.arange(32)
    below = [H & ((2 << l) - 1) for l in range(32)]  # heads at or below the lane
    dist = np.array([l - (b.bit_length() - 1) for l, b in zip(range(32), below)])  # lane + 1 when there is none
    for dl in (1, 2, 4, 8, 16):
        t = shfl_up(v, dl)
        v = np.where(dl <= dist, np.maximum(v, t), v)
    return v


def definition(H, d):
    out = np.empty(32, np.int64)
    for l in range(32):
        b = H & ((2 << l) - 1)
        start = b.bit_length() - 1 if b else 0
        out[l] = d[start:l + 1].max()
    return out


@pytest.mark.parametrize("seed", range(4))
def test_segmented_max_scan_variants_agree(seed):
    rng = np.random.default_rng(seed)
    masks = [0, 1, 0x80000000, 0xFFFFFFFF, 0x55555555, 0x00010000, 0x80000001]
    masks += [int(x) for x in rng.integers(0, 2**32, 300, dtype=np.uint64)]
    masks += [int(rng.integers(0, 2**32, dtype=np.uint64) & rng.integers(0, 2**32, dtype=np.uint64) & rng.integers(0, 2**32, dtype=np.uint64)) for _ in range(300)]
    for H in masks:
        d = rng.integers(0, 1000, 32).astype(np.int64)
        want = definition(H, d)
        assert (scan_window_masks(H, d) == want).all(), hex(H)
        assert (scan_by_distance(H, d) == want).all(), hex(H)
