import sys, time
sys.path.insert(0, '.')
import pbwt_b200 as P
M, N = 10000, 100000
eng = P.Engine(0); eng.profile(True)
p = eng.panel_synthetic(M, N, seed=0x5EED0002, L=2000, mu=2.0)
for name, flags, nocand in [('match', P.WANT_YZ | P.WANT_MAXWITHIN, 0), ('match-nocand', P.WANT_YZ | P.WANT_MAXWITHIN, 1), ('ad+yz', P.WANT_YZ | P.WANT_D, 0)]:
    eng.set_option("debug_no_candidates", nocand)
    for rep in range(3):
        r = eng.build_device(p, flags); eng.synchronize(); tm = eng.last_timings(); n = r.nmatch; r.free()
    print(name, 'nmatch', n, {k: round(v[0], 2) for k, v in tm.items() if v[1]}, flush=True)
