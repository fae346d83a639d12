import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, oracle_lib as O, pbwt_b200 as P
eng = P.Engine(0)
M, N = 1500, 96
x = O.generate_panel(M, N, seed=7, L=40, mu=4.0)
eng.set_option("windows", 3)
got = eng.build(x, M, N, with_d=True, with_matches=True)
yz = got["yz"]
r = eng.long_within(yz, M, N, 10)
a, zz, b = eng.build_reverse(yz, M, N)
eng.set_option("force_global", 1)
got2 = eng.build(x, M, N, with_d=True, with_matches=True)
print("ok", len(got["matches"]), len(got2["matches"]), len(r["matches"]), len(zz))
