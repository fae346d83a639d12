import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import oracle_lib as O
import pbwt_b200 as P
eng = P.Engine(0); orc = O.port()
for (M, N, L, W, dens) in [(33, 40, 2, 1, 0.5), (33, 40, 2, 3, 0.5), (500, 64, 4, 1, 0.3)]:
    rng = np.random.default_rng(M + L)
    hap = (rng.random((M, N)) < dens).astype(np.uint8)
    if M > 40:
        hap[3:23] = hap[3]; hap[23:33, : N // 2] = hap[3, : N // 2]
    x = O.pack_panel(hap)
    yz = orc.build(x, M, N, None, with_d=False)[0]
    want, n = orc.long_within(yz, M, N, L)
    eng.set_option("windows", W)
    got = eng.long_within(yz, M, N, L)["matches"]
    print(M, N, L, W, "want", n, "got", len(got))
    m = min(len(got), n)
    bad = np.nonzero(got[:m] != want[:m])[0]
    print("mismatches", len(bad))
    if len(bad):
        i = bad[0]
        print("first bad at", i); print("want", want[max(0, i - 3):i + 6]); print("got ", got[max(0, i - 3):i + 6])
