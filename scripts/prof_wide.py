"""Development aid for ncu: one wide-panel build (100k x 20k, a + pack3)."""
import sys
sys.path.insert(0, '.')
import pbwt_b200 as P
eng = P.Engine(0)
p = eng.panel_synthetic(100000, 20000, seed=0x5EED0003, L=1000, mu=8.0)
r = eng.build_device(p, P.WANT_YZ); eng.synchronize()
print(r.nyz)
