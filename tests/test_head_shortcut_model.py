"""CPU model of two things the sweep kernel does to matchMaximalWithin's candidates (pbwt_b200/csrc/sweep_smem.cu, round 2).

(1) The run-head shortcut.  The reference scans upward from position i while d[m+1] <= d[i] and gives up ("the match
extends") as soon as it meets i's own allele (pbwtMatch.c:124-126).  For the HEAD of a run of equal alleles the positions
above it form the previous run, all of the other allele, so the scan extends iff it gets through that whole run, i.e.
iff no d inside the previous run exceeds d[i] -- and that maximum is the very quantity the divergence update of a run
head needs (pbwtCore.c:496-505), which the kernel has in a register.  Such heads are dropped from the candidates without
a scan.  The model checks the equivalence against the reference's loop on random states, including the case the kernel
leaves out (a head whose previous run begins at position 0: d[0] = k + 1 stops every scan).

(2) The candidate test lane by lane.  Lane e of a warp tests the run boundaries of word e with the reference's first step
(a position whose neighbour in a direction it scans carries its allele is no candidate); the model checks that the set
it keeps, with the shortcut applied, still contains every position the reference reports.

(3) Not in the kernel (direction (a)): the mirror image for run TAILS, stated here so that whoever builds
it starts from a checked formulation.  The downward scan from the last position i of a run walks through the next run
(head h = i + 1, all of the other allele) and reaches the run after it -- i's allele again, head h2 -- iff nothing in
(h, h2] exceeds d[h]; when the next run is the last one the scan ends at the d[M] sentinel and never extends.  That is a
SUFFIX maximum over the next run, which the forward sweep does not have in hand at position i."""
import numpy as np
import pytest


def reference_reports(y, d, k):
    """pbwtMatch.c:122-137 for one state: (reporters, for each position whether its upward scan extended)."""
    M = len(y)
    rep, up_ext = set(), {}
    for i in range(M):
        m, n = i - 1, i + 1
        extends = False
        if d[i] <= d[i + 1]:
            while d[m + 1] <= d[i]:
                if y[m] == y[i]:
                    extends = True
                    break
                m -= 1
            up_ext[i] = extends
        if not extends and d[i] >= d[i + 1]:
            while d[n] <= d[i + 1]:
                if y[n] == y[i]:
                    extends = True
                    break
                n += 1
        if not extends:
            rep.add(i)
    return rep, up_ext


def random_state(rng, M, k, run_scale):
    """y with runs of geometric length, d with the sentinels d[0] = d[M] = k + 1 and a mix of small and large values."""
    y = np.empty(M, np.int64)
    pos, a = 0, int(rng.integers(0, 2))
    while pos < M:
        ln = 1 + int(rng.geometric(1.0 / run_scale))
        y[pos:pos + ln] = a
        pos += ln
        a ^= 1
    d = np.where(rng.random(M + 1) < 0.3, rng.integers(0, k + 1, M + 1), rng.integers(0, max(k // 8, 1) + 1, M + 1)).astype(np.int64)
    d[0] = d[M] = k + 1
    return y, d


@pytest.mark.parametrize("seed", range(6))
def test_run_head_shortcut_equals_the_upward_scan(seed):
    rng = np.random.default_rng(seed)
    for _ in range(60):
        M, k = int(rng.integers(2, 400)), int(rng.integers(1, 1000))
        y, d = random_state(rng, M, k, float(rng.choice([1.5, 4.0, 30.0])))
        _, up_ext = reference_reports(y, d, k)
        heads = [i for i in range(M) if i == 0 or y[i] != y[i - 1]]
        for j, i in enumerate(heads):
            if i == 0 or d[i] > d[i + 1]:
                continue  # no upward scan from here (position 0 holds the k + 1 sentinel)
            prev_head = heads[j - 1]
            prev_max = d[prev_head:i].max()  # max d over the previous run, what the divergence update carries to this head
            assert (prev_max <= d[i]) == up_ext[i], (M, k, i)


def reference_down_extends(y, d, k):
    """for every position whose downward scan the reference performs without an upward extension first: did it extend"""
    M = len(y)
    out = {}
    for i in range(M):
        m, extends = i - 1, False
        if d[i] <= d[i + 1]:
            while d[m + 1] <= d[i]:
                if y[m] == y[i]:
                    extends = True
                    break
                m -= 1
        if not extends and d[i] >= d[i + 1]:
            n, down = i + 1, False
            while d[n] <= d[i + 1]:
                if y[n] == y[i]:
                    down = True
                    break
                n += 1
            out[i] = down
    return out


@pytest.mark.parametrize("seed", range(6))
def test_run_tail_mirror_image_equals_the_downward_scan(seed):
    rng = np.random.default_rng(200 + seed)
    for _ in range(60):
        M, k = int(rng.integers(2, 400)), int(rng.integers(1, 1000))
        y, d = random_state(rng, M, k, float(rng.choice([1.5, 4.0, 30.0])))
        down = reference_down_extends(y, d, k)
        heads = [i for i in range(M) if i == 0 or y[i] != y[i - 1]] + [M]
        for j in range(len(heads) - 1):
            i = heads[j + 1] - 1  # the tail of run j
            if i not in down:
                continue
            if j + 2 >= len(heads):  # no run after this one
                continue
            h = heads[j + 1]
            if j + 2 == len(heads) - 1:  # the next run is the last: the scan ends at d[M] = k + 1
                assert down[i] is False, (M, k, i)
                continue
            h2 = heads[j + 2]
            assert (d[h + 1:h2 + 1].max() <= d[h]) == down[i], (M, k, i)


def kernel_candidates(y, d, k):
    """The kernel's candidate set: run heads and tails that pass the first-step test, minus the heads the shortcut resolves."""
    M = len(y)
    heads = [i for i in range(M) if i == 0 or y[i] != y[i - 1]]
    head_set = set(heads)
    prev_head = {h: heads[j - 1] for j, h in enumerate(heads) if j > 0}
    cand = set()
    for i in range(M):
        is_head = i in head_set
        is_tail = i + 1 == M or (i + 1) in head_set
        if not (is_head or is_tail):
            continue
        same_prev = not is_head
        if is_head and i > 0 and d[prev_head[i]:i].max() <= d[i]:
            same_prev = True  # the shortcut: the upward scan is known to extend
        same_next = not is_tail
        if (d[i] <= d[i + 1] and same_prev) or (d[i] >= d[i + 1] and same_next):
            continue
        cand.add(i)
    return cand


@pytest.mark.parametrize("seed", range(6))
def test_candidates_still_contain_every_reporter(seed):
    rng = np.random.default_rng(100 + seed)
    for _ in range(60):
        M, k = int(rng.integers(2, 400)), int(rng.integers(1, 1000))
        y, d = random_state(rng, M, k, float(rng.choice([1.5, 4.0, 30.0])))
        rep, _ = reference_reports(y, d, k)
        cand = kernel_candidates(y, d, k)
        assert rep <= cand, (M, k, sorted(rep - cand)[:5])
