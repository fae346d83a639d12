"""Development aid: wide-panel builds timed by kernel class (C3: 100k x 200k a only; C4-like: 1M x 20k with d)."""
import sys, time
sys.path.insert(0, '.')
import pbwt_b200 as P
eng = P.Engine(0); eng.profile(True)
cases = [(100000, 200000, 0x5EED0003, P.WANT_YZ, 'C3 a+yz'), (100000, 50000, 0x5EED0003, P.WANT_YZ | P.WANT_D, '100k x 50k ad+yz'),
         (1000000, 20000, 0x5EED0004, P.WANT_YZ | P.WANT_D, 'C4-like 1M x 20k ad+yz')]
for M, N, seed, flags, name in cases:
    p = eng.panel_synthetic(M, N, seed=seed, L=1000, mu=8.0); eng.synchronize(); eng.last_timings()
    for rep in range(3):
        t0 = time.time(); r = eng.build_device(p, flags); eng.synchronize(); dt = time.time() - t0
        tm = eng.last_timings(); r.free()
    print(name, 'wall %.4f' % dt, 'hs/s %.3e' % (M * N / dt), {k: round(v[0], 2) for k, v in tm.items() if v[0] > 0}, flush=True)
    p.free()
