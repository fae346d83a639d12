"""Development aid: pass 2 of config 2 against the sizes of the edge windows (options edge_*_pct)."""
import sys
sys.path.insert(0, '.')
import pbwt_b200 as P
M, N = 10000, 100000
eng = P.Engine(0); eng.profile(True)
p = eng.panel_synthetic(M, N, seed=0x5EED0002, L=2000, mu=2.0)
for f0, f1, fl in [(80, 92, 74), (75, 90, 70), (70, 88, 66), (66, 86, 62), (75, 90, 62), (80, 92, 66), (70, 90, 58), (62, 84, 58), (75, 95, 54)]:
    eng.set_option("edge_first_pct", f0); eng.set_option("edge_second_pct", f1); eng.set_option("edge_last_pct", fl)
    for rep in range(3):
        r = eng.build_device(p, P.WANT_YZ | P.WANT_MAXWITHIN); eng.synchronize(); tm = eng.last_timings(); r.free()
    print((f0, f1, fl), {k: round(v[0], 3) for k, v in tm.items() if k == 'sweep'}, flush=True)
