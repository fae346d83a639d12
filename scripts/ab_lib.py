"""Development aid: time config 2 with an alternative build of the library (A/B runs of compile-time constants)."""
import sys, time, os
sys.path.insert(0, '.')
import pbwt_b200.api as A
if len(sys.argv) > 1: A.LIB_PATH = os.path.abspath(sys.argv[1])
import pbwt_b200 as P
eng = P.Engine(0); eng.profile(True)
p = eng.panel_synthetic(10000, 100000, seed=0x5EED0002, L=2000, mu=2.0); eng.synchronize(); eng.last_timings()
for rep in range(4):
    t0 = time.time(); r = eng.build_device(p, P.WANT_YZ | P.WANT_MAXWITHIN); eng.synchronize(); dt = time.time() - t0
    tm = eng.last_timings(); r.free()
print(sys.argv[1:], 'wall %.4f' % dt, 'sweep %.3f' % tm['sweep'][0])
