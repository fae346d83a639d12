"""Development aid for ncu: BASELINE config 3 (100k x 200k, a + pack3) built twice; capture the second pass 2."""
import sys
sys.path.insert(0, '.')
import pbwt_b200 as P
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
eng = P.Engine(0)
p = eng.panel_synthetic(100000, N, seed=0x5EED0003, L=1000, mu=8.0)
for _ in range(2):
    r = eng.build_device(p, P.WANT_YZ); eng.synchronize()
    print(r.nyz)
    r.free()
