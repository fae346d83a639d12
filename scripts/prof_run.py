import sys
sys.path.insert(0, '.')
import pbwt_b200 as P
M, N = int(sys.argv[1]), int(sys.argv[2])
flags = int(sys.argv[3])
eng = P.Engine(0)
p = eng.panel_synthetic(M, N, seed=0x5EED0002, L=2000, mu=2.0)
r = eng.build_device(p, flags); eng.synchronize()
print(r.nyz, r.nmatch)
