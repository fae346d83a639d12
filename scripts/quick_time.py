"""Ad-hoc timing of the device-resident build on a synthetic panel (development aid, not the bench)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import pbwt_b200 as P

M = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
eng = P.Engine(0)
eng.profile(True)
t0 = time.time(); p = eng.panel_synthetic(M, N, seed=0x5EED0002, L=2000, mu=2.0); eng.synchronize()
print('synth', time.time() - t0, eng.last_timings()['synth'])
for name, flags in [('a+yz', P.WANT_YZ), ('ad+yz', P.WANT_YZ | P.WANT_D), ('ad+yz+match', P.WANT_YZ | P.WANT_MAXWITHIN)]:
    for rep in range(2):
        t0 = time.time(); r = eng.build_device(p, flags); eng.synchronize(); dt = time.time() - t0
        tm = eng.last_timings()
        print(name, 'wall %.4f' % dt, 'hs/s %.3e' % (M * N / dt), 'nyz', r.nyz, 'nmatch', r.nmatch,
              {k: v for k, v in tm.items() if v[1]})
        r.free()
