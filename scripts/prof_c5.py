"""Development aid for ncu: the window-parallel query sweep on a config-5 shaped panel (1,000 queries vs 100,000 haplotypes)."""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import pbwt_b200 as P
from fullsize_checks import split_panel
M, Q = 100000, 1000
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
eng = P.Engine(0)
both = eng.panel_synthetic(M + Q, N, seed=0x5EED0005, L=1000, mu=8.0)
x = both.to_host(); both.free()
px, qx = split_panel(x, M, Q); del x
panel, queries = eng.panel_from_host(px, M), eng.panel_from_host(qx, Q)
rp = eng.build_device(panel, P.WANT_YZ); pyz = rp.fetch()["yz"]; rp.free()
rq = eng.build_device(queries, P.WANT_YZ); qyz = rq.fetch()["yz"]; rq.free()
rec, nom = eng.match_sweep(pyz, M, N, qyz, Q)
print(len(rec), len(nom))
