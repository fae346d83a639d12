import sys, time
sys.path.insert(0, '.')
import pbwt_b200 as P
eng = P.Engine(0)
eng.profile(True)
for M, N, L, mu, flags, name in [(100000, 20000, 1000, 8.0, P.WANT_YZ, 'C3-like a+yz'), (100000, 20000, 1000, 8.0, P.WANT_YZ | P.WANT_D, 'ad+yz'),
                                 (1000000, 3000, 1000, 8.0, P.WANT_YZ | P.WANT_D, 'C4-like ad+yz')]:
    t0 = time.time(); p = eng.panel_synthetic(M, N, seed=0x5EED0003, L=L, mu=mu); eng.synchronize(); tg = time.time() - t0
    eng.last_timings()
    for rep in range(2):
        t0 = time.time(); r = eng.build_device(p, flags); eng.synchronize(); dt = time.time() - t0
        tm = eng.last_timings()
        print(name, M, N, 'gen %.2fs' % tg, 'wall %.4f' % dt, 'hs/s %.3e' % (M * N / dt), 'nyz', r.nyz, {k: (round(v[0], 2), v[1]) for k, v in tm.items() if v[1]}, flush=True)
        r.free()
    p.free()
