"""Development aid: config 2's step by what the sweep is asked to do (no match / candidates detected but not scanned /
scans without records (-stats) / records)."""
import sys, time
sys.path.insert(0, '.')
import pbwt_b200 as P
M, N = 10000, 100000
eng = P.Engine(0); eng.profile(True)
p = eng.panel_synthetic(M, N, seed=0x5EED0002, L=2000, mu=2.0)
MW = P.WANT_MAXWITHIN
for name, flags, nocand in [('ad+yz', P.WANT_YZ | P.WANT_D, 0), ('match-nocand', P.WANT_YZ | MW, 1), ('stats', P.WANT_YZ | MW | P.WANT_STATS, 0),
                            ('match', P.WANT_YZ | MW, 0)]:
    eng.set_option("debug_no_candidates", nocand)
    for rep in range(3):
        r = eng.build_device(p, flags); eng.synchronize(); tm = eng.last_timings(); n = r.nmatch; r.free()
    print(name, 'nmatch', n, {k: round(v[0], 2) for k, v in tm.items() if v[1]}, flush=True)
eng.set_option("debug_no_candidates", 0)
