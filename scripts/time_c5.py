"""Development aid: BASELINE config 5 (1,000 queries vs a 100,000-haplotype panel) timed by kernel class."""
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import pbwt_b200 as P
from fullsize_checks import split_panel

M, Q = 100000, 1000
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
eng = P.Engine(0); eng.profile(True)
both = eng.panel_synthetic(M + Q, N, seed=0x5EED0005, L=1000, mu=8.0)
x = both.to_host(); both.free()
px, qx = split_panel(x, M, Q); del x
panel, queries = eng.panel_from_host(px, M), eng.panel_from_host(qx, Q)
rp = eng.build_device(panel, P.WANT_YZ); pyz = rp.fetch()["yz"]; rp.free()
rq = eng.build_device(queries, P.WANT_YZ); qyz = rq.fetch()["yz"]; rq.free()
panel.free(); queries.free()
eng.last_timings()
for w in [int(a) for a in sys.argv[2:]] or [0]:
    eng.set_option("windows", w)
    for rep in range(2):
        t0 = time.time(); rec, nom = eng.match_sweep(pyz, M, N, qyz, Q); dt = time.time() - t0
        tm = eng.last_timings()
        dev = sum(v[0] for v in tm.values())
        print('windows', w, 'wall %.3f s' % dt, 'device %.1f ms' % dev, 'panel hs/s (device) %.3e' % (M * N / (dev * 1e-3)), len(rec), len(nom),
              {k: (round(v[0], 2), v[1]) for k, v in tm.items() if v[0] > 0}, flush=True)
