"""CPU model of the window-parallel query sweep (pbwt_b200/csrc/query_chain.cu, SweepArgs::queryMode): per-window
CLAMPED runs of matchSequencesSweep's state machine (pbwtMatch.c:376-426) from the fresh state, stitched by the
boundary fix-up, reproduce the sequential sweep's state (f, e) of every query at every window boundary -- including
sites whose allele is absent from the panel ("no match" events) and every branch of the fix-up."""
import numpy as np
import pytest

import oracle_lib as O

NONE = 1 << 30


def site_step(f, e, x, k, y, d, M, clamp=0):
    """One site for one query, in the closed form the device uses (nearest carriers above / below, tier = range max of
    d, up wins ties); divergence tiers below `clamp` count as `clamp`.  -> (f, e) entering site k + 1."""
    if y[f] != x:
        hi, moved = f + 1, False
        while hi < M and d[hi] <= e:
            if y[hi] == x:
                f, moved = hi, True
                break
            hi += 1
        if not moved:
            u = f - 1
            while u >= 0 and y[u] != x:
                u -= 1
            v = hi
            while v < M and y[v] != x:
                v += 1
            if u < 0 and v >= M:
                e = k + 1  # pbwtMatch.c:405-410
            else:
                t_up = max(max(d[u + 1:f + 1]), clamp) if u >= 0 else NONE
                t_down = max(max(d[hi:v + 1]), clamp) if v < M else NONE
                if t_up <= t_down:
                    t = u
                    while t > 0 and d[t] <= t_up:
                        t -= 1
                    f = t
                    while y[f] != x:
                        f += 1
                    e = t_up
                else:
                    f, e = v, t_down
    z_before, c = int((y[:f] == 0).sum()), int((y == 0).sum())
    nf = c + f - z_before if x else z_before
    return (0 if nf == M else nf), e


def fix_up(state, clamped, K0, a0, d0, a1, d1, M, branches):
    """true state entering the next window from the true state entering this one and this window's clamped end state"""
    f0, e0 = state
    fp, ep = clamped
    if ep > K0:
        branches["inside"] += 1
        return fp, ep
    inv0 = np.empty(M, np.int64)
    inv0[a0] = np.arange(M)
    hp = fp + 1
    while hp < M and d1[hp] <= K0:
        hp += 1
    P = [int(inv0[a1[i]]) for i in range(fp, hp)]  # positions at K0 of the block's members: increasing (stable sort)
    assert all(P[i] < P[i + 1] for i in range(len(P) - 1))
    h0 = f0 + 1
    while h0 < M and d0[h0] <= e0:
        h0 += 1
    i0 = 0
    while i0 < len(P) and P[i0] < f0:
        i0 += 1
    if i0 < len(P) and P[i0] < h0:
        branches["member of the old block"] += 1
        return fp + i0, e0
    j1 = P[i0 - 1] if i0 > 0 else -1
    j2 = P[i0] if i0 < len(P) else M
    Dm = max(d0[j1 + 1:f0 + 1]) if j1 >= 0 else NONE
    Dp = max(d0[h0:j2 + 1]) if j2 < M else NONE
    if Dm <= Dp:
        t = j1
        while t > 0 and d0[t] <= Dm:
            t -= 1
        i1 = 0
        while P[i1] < t:
            i1 += 1
        branches["up"] += 1
        return fp + i1, int(Dm)
    branches["down"] += 1
    return fp + i0, int(Dp)


BRANCHES = {"inside": 0, "member of the old block": 0, "up": 0, "down": 0}


@pytest.mark.parametrize("M,Q,N,W,seed", [(60, 9, 120, 5, 1), (200, 12, 260, 7, 2), (33, 30, 90, 9, 3), (150, 10, 200, 4, 4)])
def test_clamped_windows_plus_chain_equal_the_sequential_sweep(M, Q, N, W, seed):
    x = O.generate_panel(M + Q, N, seed=100 + seed, L=30, mu=6.0)
    hap = O.unpack_panel(x, M + Q)
    rng = np.random.default_rng(seed)
    for k in rng.choice(N, size=max(1, N // 25), replace=False):  # monomorphic panel columns: "no match" events
        hap[:M, k] = rng.integers(0, 2)
    a_tr, d_tr, y_tr, _ = O.port().trace(O.pack_panel(hap[:M]), M, N)
    qh = hap[M:]
    f, e = [0] * Q, [0] * Q
    truth = [list(zip(f, e))]
    for k in range(N):
        for q in range(Q):
            f[q], e[q] = site_step(f[q], e[q], qh[q, k], k, y_tr[k], d_tr[k], M)
        truth.append(list(zip(f, e)))
    bounds = [round(i * N / W) for i in range(W + 1)]
    branches = BRANCHES
    state = [(0, 0)] * Q
    for w in range(W - 1):
        K0, K1 = bounds[w], bounds[w + 1]
        nxt = []
        for q in range(Q):
            ff, ee = 0, K0
            for k in range(K0, K1):
                ff, ee = site_step(ff, ee, qh[q, k], k, y_tr[k], d_tr[k], M, clamp=K0)
            nxt.append(fix_up(state[q], (ff, ee), K0, a_tr[K0], d_tr[K0], a_tr[K1], d_tr[K1], M, branches))
        state = nxt
        assert state == truth[K1], f"state entering window {w + 1} differs"


def test_every_branch_of_the_fix_up_was_taken():
    assert all(BRANCHES.values()), BRANCHES
