import sys
sys.path.insert(0, '.')
import pbwt_b200 as P
M, N = int(sys.argv[1]), int(sys.argv[2])
eng = P.Engine(0)
p = eng.panel_synthetic(M, N, seed=0x5EED0002, L=2000, mu=2.0)
r = eng.build_device(p, P.WANT_YZ | P.WANT_MAXWITHIN); r.free()
eng.set_option("window_times", 1)
r = eng.build_device(p, P.WANT_YZ | P.WANT_MAXWITHIN); r.free()
r = eng.build_device(p, P.WANT_YZ | P.WANT_D); r.free()
