"""Ad-hoc timing of the wide-panel kernels (development aid, not the bench): thread-block clusters vs the
cooperative-grid path, build / build+maxWithin / query sweep."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import pbwt_b200 as P

eng = P.Engine(0)
eng.profile(True)
cases = [(100000, 20000, 1000, 8.0, P.WANT_YZ, 'C3-like a+yz'), (100000, 20000, 1000, 8.0, P.WANT_YZ | P.WANT_D, 'ad+yz'),
         (100000, 20000, 1000, 8.0, P.WANT_YZ | P.WANT_MAXWITHIN, 'ad+yz+maxWithin'),
         (1000000, 3000, 1000, 8.0, P.WANT_YZ | P.WANT_D, 'C4-like ad+yz')]
for M, N, L, mu, flags, name in cases:
    t0 = time.time(); p = eng.panel_synthetic(M, N, seed=0x5EED0003, L=L, mu=mu); eng.synchronize(); tg = time.time() - t0
    eng.last_timings()
    for nc in (0, 1):
        eng.set_option("no_cluster", nc)
        for rep in range(2):
            t0 = time.time(); r = eng.build_device(p, flags); eng.synchronize(); dt = time.time() - t0
            tm = eng.last_timings()
            print(name, 'no_cluster', nc, M, N, 'wall %.4f' % dt, 'hs/s %.3e' % (M * N / dt), 'nyz', r.nyz, 'nmatch', r.nmatch,
                  {k: (round(v[0], 2), v[1]) for k, v in tm.items() if v[1]}, flush=True)
            r.free()
    eng.set_option("no_cluster", 0)
    p.free()

# C5-like: 1000 queries against a 100,000-haplotype panel
M, Q, N = 100000, 1000, 20000
both = eng.panel_synthetic(M + Q, N, seed=0x5EED0005, L=1000, mu=8.0)
x = both.to_host(); both.free()
bits = np.unpackbits(np.ascontiguousarray(x).view(np.uint8), axis=1, bitorder="little")
def repack(b):
    n, m = b.shape
    wpc = P.words_per_col(m)
    pad = np.zeros((n, wpc * 32), np.uint8); pad[:, :m] = b
    return np.ascontiguousarray(np.packbits(pad, axis=1, bitorder="little")).view("<u4").reshape(n, wpc)
px, qx = repack(bits[:, :M]), repack(bits[:, M:M + Q])
panel, queries = eng.panel_from_host(px, M), eng.panel_from_host(qx, Q)
rp = eng.build_device(panel, P.WANT_YZ); pyz = rp.fetch()["yz"]; rp.free()
rq = eng.build_device(queries, P.WANT_YZ); qyz = rq.fetch()["yz"]; rq.free()
eng.last_timings()
for nc in (0, 1):
    eng.set_option("no_cluster", nc)
    for rep in range(2):
        t0 = time.time(); rec, nom = eng.match_sweep(pyz, M, N, qyz, Q); dt = time.time() - t0
        tm = eng.last_timings()
        print('C5-like query sweep no_cluster', nc, 'wall %.3f' % dt, 'panel hs/s %.3e' % (M * N / dt), len(rec), len(nom),
              {k: (round(v[0], 2), v[1]) for k, v in tm.items() if v[1]}, flush=True)
eng.set_option("no_cluster", 0)
