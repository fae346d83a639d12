"""CPU test of the closed form that replaces matchSequencesSweep's reseed loop on wide panels (DESIGN.md section 7,
pbwt_b200/csrc/sweep_global.cu): the reference (pbwtMatch.c:389-411) widens the interval [f, hi) tier by tier of
divergence, upper side first on ties, until a tier holds the query's allele x.  Closed form: with u / v the nearest
carriers of x above f / at or below hi, the up tier that reaches u is max d(u, f], the down tier that reaches v is
max d[hi, v]; up wins ties; the new f is the topmost carrier of the winning up tier, or v itself."""
import numpy as np
import pytest


def reseed_reference(d, y, M, k, f, hi, x):
    """Line by line after pbwtMatch.c:389-411; returns (f, e, no_match)."""
    i_minus, i_plus = f, hi
    d_plus = d[i_plus] if i_plus < M else k
    d_minus = d[i_minus]
    while True:
        if d_minus <= d_plus:
            i = -1
            while d[i_minus] <= d_minus:  # d[0] = k+1 prevents underflow
                i_minus -= 1
                if y[i_minus] == x:
                    i = i_minus
            if i >= 0:
                return i, d_minus, False
            d_minus = d[i_minus]
        else:
            while i_plus < M and d[i_plus] <= d_plus:
                if y[i_plus] == x:
                    return i_plus, d_plus, False
                i_plus += 1
            d_plus = k if i_plus == M else d[i_plus]
            if i_minus == 0 and i_plus == M:
                return f, k + 1, True


def reseed_closed_form(d, y, M, k, f, hi, x):
    up = [p for p in range(f - 1, -1, -1) if y[p] == x]
    down = [p for p in range(hi, M) if y[p] == x]
    if not up and not down:
        return f, k + 1, True
    t_up = max(d[up[0] + 1:f + 1]) if up else None
    t_down = max(d[hi:down[0] + 1]) if down else None
    if up and (not down or t_up <= t_down):
        q = up[0]
        while q > 0 and d[q] <= t_up:  # start of the tier: nearest q <= u with d[q] > t_up (d[0] is the sentinel)
            q -= 1
        top = min(p for p in range(q, up[0] + 1) if y[p] == x)
        return top, t_up, False
    return down[0], t_down, False


@pytest.mark.parametrize("M,dens,seed", [(6, 0.5, 1), (40, 0.5, 2), (40, 0.1, 3), (200, 0.02, 4), (200, 0.5, 5), (60, 0.0, 6)])
def test_reseed_closed_form_equals_tier_loop(M, dens, seed):
    rng = np.random.default_rng(seed)
    for _ in range(400):
        k = int(rng.integers(1, 50))
        d = rng.integers(0, k + 1, size=M + 1)
        d[0] = k + 1
        d[M] = k + 1
        y = (rng.random(M) < dens).astype(int)
        f = int(rng.integers(0, M))
        x = 1 - y[f]
        hi = f + 1  # [f, hi): a run without allele x (what the alternates scan leaves behind)
        while hi < M and y[hi] != x and rng.random() < 0.7:
            hi += 1
        assert reseed_reference(d, y, M, k, f, hi, x) == reseed_closed_form(d, y, M, k, f, hi, x)
