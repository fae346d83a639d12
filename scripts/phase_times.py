"""Development aid: per-phase cycles of the wide-panel kernel (option window_times), pass 2 only."""
import sys
sys.path.insert(0, '.')
import pbwt_b200 as P

eng = P.Engine(0)
for M, N, flags, name in [(100000, 20000, P.WANT_YZ, 'a'), (100000, 20000, P.WANT_YZ | P.WANT_D, 'ad'),
                          (100000, 20000, P.WANT_YZ | P.WANT_MAXWITHIN, 'ad+match'), (1000000, 3000, P.WANT_YZ | P.WANT_D, 'ad 1M'),
                          (30000, 20000, P.WANT_YZ | P.WANT_D, 'ad 30k')]:
    p = eng.panel_synthetic(M, N, seed=0x5EED0003, L=1000, mu=8.0)
    for nc in (0, 1):
        eng.set_option("no_cluster", nc)
        eng.set_option("window_times", 1)
        print(name, M, N, 'no_cluster', nc, flush=True)
        r = eng.build_device(p, flags); eng.synchronize(); r.free()
        eng.set_option("window_times", 0)
    eng.set_option("no_cluster", 0)
    p.free()
