"""Development aid: pass 2 of config 2 with the match candidates discarded (debug_no_candidates), for builds that leave
out one piece of the match bookkeeping each."""
import sys
sys.path.insert(0, '.')
import pbwt_b200 as P
M, N = 10000, 100000
eng = P.Engine(0); eng.profile(True)
p = eng.panel_synthetic(M, N, seed=0x5EED0002, L=2000, mu=2.0)
eng.set_option("debug_no_candidates", 1)
for rep in range(3):
    r = eng.build_device(p, P.WANT_YZ | P.WANT_MAXWITHIN); eng.synchronize(); tm = eng.last_timings(); r.free()
print('nocand sweep %.3f' % tm['sweep'][0], flush=True)
