"""Development aid: kernel-class timings of the full BASELINE config 3 (100k x 200k build + pack3)."""
import sys, time
sys.path.insert(0, '.')
import pbwt_b200 as P
eng = P.Engine(0); eng.profile(True)
M, N = 100000, 200000
p = eng.panel_synthetic(M, N, seed=0x5EED0003, L=1000, mu=8.0); eng.synchronize(); eng.last_timings()
for rep in range(2):
    t0 = time.time(); r = eng.build_device(p, P.WANT_YZ); eng.synchronize(); dt = time.time() - t0
    print('C3', 'wall %.4f' % dt, 'hs/s %.3e' % (M * N / dt), {k: (round(v[0], 2), v[1]) for k, v in eng.last_timings().items() if v[1]}, flush=True)
    r.free()
