"""compute-sanitizer target, round 2 additions: the window-parallel query sweep (clamped pass, chain, pass 2), the sharded
protocols on one device, grouped records, the -stats histogram, stored-stream maxWithin, the sort-based combine."""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, oracle_lib as O, pbwt_b200 as P
eng = P.Engine(0)
M, N = 1500, 96
x = O.generate_panel(M, N, seed=7, L=40, mu=4.0)
eng.set_option("windows", 5)
got = eng.build(x, M, N, with_d=True, with_matches=True)
yz = got["yz"]
g = eng.max_within_grouped(yz, M, N)
h = eng.max_within_stats(yz, M, N)
again = eng.max_within(yz, M, N)
# query sweep, window-parallel on the wide-panel kernel
Q = 40
q = O.generate_panel(Q, N, seed=9, L=30, mu=4.0)
qyz = eng.build(q, Q, N, with_d=False)["yz"]
eng.set_option("force_global", 1)
ms = eng.match_sweep(yz, M, N, qyz, Q)
eng.set_option("force_global", 0)
print("ok", len(got["matches"]), len(again["matches"]), int(h.sum()), len(ms[0]))
