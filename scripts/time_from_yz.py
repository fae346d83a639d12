import sys, time
sys.path.insert(0, '.')
import numpy as np
import pbwt_b200 as P
M, N = 10000, int(sys.argv[1]) if len(sys.argv) > 1 else 100000
eng = P.Engine(0); eng.profile(True)
p = eng.panel_synthetic(M, N, seed=0x5EED0002, L=2000, mu=2.0)
r = eng.build_device(p, P.WANT_YZ | P.WANT_MAXWITHIN); out = r.fetch(); r.free(); eng.last_timings()
for rep in range(3):
    t0 = time.time(); res = eng.sweep_yz_device(out["yz"], M, N, P.WANT_MAXWITHIN); eng.synchronize(); dt = time.time() - t0
    tm = eng.last_timings()
    print('sweep_yz_device wall %.4f s' % dt, 'nmatch', res.nmatch, {k: (round(v[0], 2), v[1]) for k, v in tm.items() if v[1]}, flush=True)
    got = res.fetch(); res.free()
print('same as build path:', (got["a_end"] == out["a_end"]).all(), (got["d_end"] == out["d_end"]).all(), (got["matches"] == out["matches"]).all())
